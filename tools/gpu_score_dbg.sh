#!/bin/bash
# where the scoring kernel's time goes: MFK_TC_DEBUG bits (1 no epilogue scan, 2 no MMAs, 4 no drain rounds, 8 nothing parked)
for d in 0 3 1 8; do
  echo "debug $d: $(MFK_TC_DEBUG=$d python tools/score_bench.py --workload ml-20m --all-users 2>&1 | tail -1 | cut -c88-130)"
done | tee gpurun_out/sc_debug.log
for d in 3 1 8; do
  echo "ES=1 debug $d: $(MFK_SCORE_ES=1 MFK_TC_DEBUG=$d python tools/score_bench.py --workload ml-20m --all-users 2>&1 | tail -1 | cut -c88-130)"
done | tee -a gpurun_out/sc_debug.log
