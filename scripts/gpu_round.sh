#!/bin/bash
# One GPU session: tests, smoke, bench, ncu launch list + one full capture.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --batch 16384 > gpurun_out/bench_b16k.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --batch 65536 > gpurun_out/bench_b64k.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --precision fp64 > gpurun_out/bench_fp64.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 6 -c 2 -o gpurun_out/prof_step python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
cat gpurun_out/bench.json | cut -c1-1500
