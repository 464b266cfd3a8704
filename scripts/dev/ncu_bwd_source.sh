#!/bin/bash
# Dev: one full ncu capture of k_cstep_bwd (Atlas + ground, B from $B) with per-instruction counters exported as CSV
set -e
mkdir -p gpurun_out
MODEL=atlas_ground B=${B:-8192} ncu --set full --clock-control none --import-source on -k regex:k_cstep_bwd -s 1 -c 1 -o /tmp/bwd_full -f \
    python scripts/dev/one_contact.py > gpurun_out/ncu_bwd.log 2>&1
ncu -i /tmp/bwd_full.ncu-rep --page source --print-source sass --csv > gpurun_out/r02_bwd_sass.csv 2>/dev/null
ncu -i /tmp/bwd_full.ncu-rep --page raw --csv > gpurun_out/r02_bwd_raw.csv 2>/dev/null
ls -la gpurun_out/r02_bwd_sass.csv gpurun_out/r02_bwd_raw.csv
