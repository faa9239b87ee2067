#!/bin/bash
V=$PWD/matrix_factorization_b200/csrc/variants
for v in "" spin4 spin16; do
  if [ -z "$v" ]; then unset MFK_LIB_PATH; else export MFK_LIB_PATH=$V/libmfk_$v.so; fi
  for wl in ml-20m netflix; do
  python bench.py --workload $wl --steps 3 --warmup 3 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('variant [$v] $wl', 'kernel_ms', round(r['kernel_ms'],2), [round(k['ms'],2) for k in r['per_kernel']])
"
  done
done
