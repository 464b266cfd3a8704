#!/bin/bash
set -x
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 12 gpurun_out/pytest_gpu.log; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['kernel_ms'],d['e2e']['value']);print(json.dumps(d['extra'],indent=1))"; tail -5 gpurun_out/bench.err
