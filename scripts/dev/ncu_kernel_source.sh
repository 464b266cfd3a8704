#!/bin/bash
# Dev: one full ncu capture of ONE contact kernel (regex $1, launch index $2) with per-instruction counters exported as CSV
set -e
mkdir -p gpurun_out
K=${1:-k_csolve}; SKIP=${2:-1}; TAG=${3:-ksrc}
MODEL=atlas_ground B=${B:-8192} ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o /tmp/${TAG}_full -f \
    python scripts/dev/one_contact.py > gpurun_out/ncu_${TAG}.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page source --print-source sass --csv > gpurun_out/r02_${TAG}_sass.csv 2>/dev/null
ls -la gpurun_out/r02_${TAG}_sass.csv
