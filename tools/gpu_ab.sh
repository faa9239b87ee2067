#!/bin/bash
# A/B of compile-time variants on the isolated chain and on phase 1:  tools/gpu_ab.sh "DEF1" "DEF2" ...
for defs in "$@"; do
  MFK_NVCC_DEFS="$defs" python -m matrix_factorization_b200.build --force > /dev/null 2>&1
  echo "== $defs: $(python tools/prof_chain.py 2>&1 | tail -1 | cut -c1-70) | phase1 $(python tools/prof_hot.py --phases 1 --epochs 3 2>&1 | grep 'epoch ms' | tail -1 | cut -c1-28)"
done
