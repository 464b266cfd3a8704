#!/bin/bash
# Dev: sweep the worlds-per-block caps of the lockstep contact kernels (NB2_WPB="build,solve,apply,bwd")
for w in "0,0,0,0" "0,8,0,0" "0,6,0,0" "0,4,0,0" "0,3,0,0" "0,0,0,4" "0,0,0,3" "0,0,0,2" "4,0,0,0" "2,0,0,0"; do
  echo "== NB2_WPB=$w"
  NB2_WPB=$w MODELS=${MODELS:-atlas_ground} BS=${BS:-8192} python scripts/dev/bench_contact.py 2>&1 | grep "rollout T=8" | sed 's/| last step.*//'
done
