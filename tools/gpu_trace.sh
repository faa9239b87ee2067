#!/bin/bash
# traced profile build of the batch engine: per-phase cycle counters and one iteration's per-warp time stamps
tag=${1:-x}
MFK_RING_PROFILE=1 python -m matrix_factorization_b200.build --force > /dev/null 2>&1
python tools/ring_stats.py --workload ml-20m 2>&1 | grep -E "hot worker 0|trace warp" | tee gpurun_out/tr_${tag}.log
