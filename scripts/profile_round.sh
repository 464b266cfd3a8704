#!/bin/bash
# Round profile: (1) launch list of the default bench command, (2) one full ncu capture of the step kernels,
# (3) summaries copied into profiles/ by scripts/summarize_profile.py (run afterwards in the build container).
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 12 -c 2 -o gpurun_out/${TAG}_step_full -f \
    python bench.py --steps 4 --warmup 5 --no-extra > gpurun_out/${TAG}_ncu_step.log 2>&1
ncu -i gpurun_out/${TAG}_step_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_step_full.ncu-rep --page details --csv > gpurun_out/${TAG}_step_full_details.csv 2>/dev/null
python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
ls -la gpurun_out | tail -12
