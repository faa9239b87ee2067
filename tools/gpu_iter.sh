#!/bin/bash
# One GPU iteration on the hot engine: parity tests, per-phase epoch times (normal build), phase cycle counters (profile build).
#   tools/gpu_iter.sh <tag> [pytest -k expression]
tag=${1:-x}
kexpr=${2:-"hot or batch or flat or replay"}
python -m pytest tests/test_hot_engine_gpu.py tests/test_kmf_gpu.py -x -q -k "$kexpr" > gpurun_out/it_${tag}_tests.log 2>&1
tail -3 gpurun_out/it_${tag}_tests.log
for ph in 1 2 4 7; do python tools/prof_hot.py --phases $ph --epochs 3 2>&1 | grep "epoch ms" | tail -1 | cut -c1-40 | sed "s/^/phases=$ph /"; done | tee gpurun_out/it_${tag}_phases.log
MFK_RING_PROFILE=1 python -m matrix_factorization_b200.build --force > /dev/null 2>&1
python tools/ring_stats.py --workload ml-20m 2>&1 | grep -E "hot worker 0|hot item phase|hot user phase|flat worker " | tee gpurun_out/it_${tag}_stats.log
