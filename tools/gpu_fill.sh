#!/bin/bash
# hot-id rule sweep at the ML-20M shape: MFK_HOT_FILL (ratings per step the lightest hot id must still bring) and MFK_HOT_FLOOR
for cfg in "32 256" "24 256" "16 256" "12 128" "8 128"; do
  set -- $cfg
  MFK_HOT_FILL=$1 MFK_HOT_FLOOR=$2 python bench.py --steps 3 --warmup 3 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['config']['plan']; r=d['roofline']
print('fill $1 floor $2', 'kernel_ms', round(r['kernel_ms'],2), 'ms/step', round(d['ms_per_step'],2), [round(k['ms'],2) for k in r['per_kernel']], p['n_hot_items'], p['n_hot_ratings'], p['n_hot_users'], p['n_hot_user_ratings'], p['n_hot_workers'], p['n_hot_user_workers'], p['hot_parallel'])
"
done | tee gpurun_out/fill_sweep.log
