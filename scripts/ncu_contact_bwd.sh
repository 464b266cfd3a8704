#!/bin/bash
set -x
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
ncu --set full --clock-control none --import-source on -k regex:k_step_bwd_contact -s 2 -c 1 -o gpurun_out/r01_contact_bwd_full -f \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_contact_bwd.log 2>&1
ncu -i gpurun_out/r01_contact_bwd_full.ncu-rep --page raw --csv > gpurun_out/r01_contact_bwd_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -3
