#!/bin/bash
# Round-2 profile (run on the GPU box through gpurun; summaries: scripts/summarize_r02.py in the build container):
#  (1) launch list of the default bench command, (2) full ncu capture of the two step kernels, (3) launch list and (4) full capture of the five
#  kernels of one Atlas + ground contact step at B = 8192, (5) the bench line itself.  Reports stay in /tmp (they exceed the gpurun_out quota);
#  only the CSV exports travel back.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 12 -c 2 -o /tmp/r02_step_full -f \
    python bench.py --steps 4 --warmup 5 --no-extra > gpurun_out/r02_ncu_step.log 2>&1
ncu -i /tmp/r02_step_full.ncu-rep --page raw --csv > gpurun_out/r02_step_full_raw.csv 2>/dev/null
MODEL=atlas_ground B=8192 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_contact_launches.csv \
    python scripts/dev/one_contact.py > gpurun_out/r02_contact_launches.log 2>&1
MODEL=atlas_ground B=8192 ncu --set full --clock-control none --import-source on -k regex:k_c -s 5 -c 5 -o /tmp/r02_contact_full -f \
    python scripts/dev/one_contact.py > gpurun_out/r02_ncu_contact.log 2>&1
ncu -i /tmp/r02_contact_full.ncu-rep --page raw --csv > gpurun_out/r02_contact_full_raw.csv 2>/dev/null
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
ls -la gpurun_out | tail -12
