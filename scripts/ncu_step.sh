#!/bin/bash
# one full ncu capture of the headline kernels (forward + backward) at the bench configuration
set -x
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
ncu --set full --clock-control none --import-source on -k regex:k_step -s 12 -c 2 -o gpurun_out/r01_step_coop -f \
    python bench.py --steps 4 --warmup 5 --no-extra > gpurun_out/ncu_step.log 2>&1
ncu -i gpurun_out/r01_step_coop.ncu-rep --page raw --csv > gpurun_out/r01_step_coop_raw.csv 2>/dev/null
ls -la gpurun_out/
