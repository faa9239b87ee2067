#!/bin/bash
# A/B of library variants (tools/build_variant.sh) on the isolated chain and the hot-item phase
out=gpurun_out/ab.log
: > $out
V=matrix_factorization_b200/csrc/variants
for v in "" $@; do
  if [ -z "$v" ]; then unset MFK_LIB_PATH; name=default; else export MFK_LIB_PATH=$PWD/$V/libmfk_$v.so; name=$v; fi
  a=$(timeout 120 python tools/prof_chain.py --epochs 3 2>&1 | tail -1 | awk '{print $NF}')
  b=$(timeout 120 python tools/prof_chain.py --factors 256 --users 200000 --epochs 3 2>&1 | tail -1 | awk '{print $NF}')
  c=$(timeout 200 python tools/prof_hot.py --phases 1 --epochs 3 2>&1 | grep "epoch ms" | tail -1 | cut -c1-26)
  echo "$name chain128 $a chain256 $b phase1 $c" | tee -a $out
done
