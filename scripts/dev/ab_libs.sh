#!/bin/bash
# A/B of prebuilt libnb2 variants (dev only): copies each over the in-tree library, runs the parity tests and the bench
cd "$(dirname "$0")/../.."
cp nimblephysics_b200/csrc/libnb2.so /tmp/libnb2_orig.so
for rep in 1 2; do
for v in build_variants/*.so; do
  cp $v nimblephysics_b200/csrc/libnb2.so
  if [ $rep = 1 ]; then timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1; fi
  python bench.py --no-extra 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v','%.3e'%d['value'],d['kernel_ms'],'e2e %.3e'%d['e2e']['value'])"
done; done
cp /tmp/libnb2_orig.so nimblephysics_b200/csrc/libnb2.so
