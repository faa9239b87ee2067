#!/bin/bash
# timing experiments on the isolated chain (one hot item, one CTA): MFK_BT_EXP bits switch parts of the batch engine off
out=gpurun_out/exp_chain.log
: > $out
for e in ${@:-0 1 2 4 8 16 32}; do
  echo "EXP $e: $(MFK_BT_EXP=$e timeout 120 python tools/prof_chain.py --epochs 3 2>&1 | tail -1)" | tee -a $out
done
