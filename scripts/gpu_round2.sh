#!/bin/bash
set -x
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_contact -s 2 -c 1 -o gpurun_out/prof_contact python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_full2.log 2>&1
tail -n 8 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['kernel_ms'],d['e2e']['value']);print(json.dumps(d['extra'],indent=1))"
