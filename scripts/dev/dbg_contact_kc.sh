#!/bin/bash
cd "$(dirname "$0")/../.."
for w in 32 8 4; do
NB2_CONTACT_BWD_WPW=$w python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());e=d['extra'];print('BWD_WPW $w', {k:(round(v.get('world_steps_per_s',0)),round(v.get('fwd_bwd_world_steps_per_s',0))) for k,v in e.items()})"
done
