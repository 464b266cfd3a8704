#!/bin/bash
# cooperative-lane sweep: parity tests, then batch x lanes timing grid, then the default bench line
set -x
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
: > gpurun_out/lanes_sweep.txt
for B in 1024 4096 16384 65536; do for K in 0 1 2 4; do
  timeout 300 python bench.py --steps 30 --warmup 5 --batch $B --lanes $K --no-extra 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('B',d['config']['global_batch'],'K',d['lanes_per_world'],'value %.3e'%d['value'],d['kernel_ms'],'e2e %.3e'%d['e2e']['value'])" >> gpurun_out/lanes_sweep.txt
done; done
cat gpurun_out/lanes_sweep.txt
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['kernel_ms'],d['lanes_per_world'],d['e2e']['value']);print(json.dumps(d['extra'],indent=1))"; tail -5 gpurun_out/bench.err
