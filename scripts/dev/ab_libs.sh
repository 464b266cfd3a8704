#!/bin/bash
# A/B of prebuilt libnb2 variants (dev only): copies each over the in-tree library and runs the bench
cd "$(dirname "$0")/../.."
cp nimblephysics_b200/csrc/libnb2.so /tmp/libnb2_orig.so
for v in build_variants/*.so; do
  cp $v nimblephysics_b200/csrc/libnb2.so
  for B in 4096 65536; do
    python bench.py --no-extra --batch $B 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v',d['config']['global_batch'],'%.3e'%d['value'],d['kernel_ms'],d['lanes_per_world'],'e2e %.3e'%d['e2e']['value'])"
  done
done
cp /tmp/libnb2_orig.so nimblephysics_b200/csrc/libnb2.so
