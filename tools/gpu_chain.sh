#!/bin/bash
# quick loop on the batch engine: parity tests, the isolated chain (one hot item, one CTA), then a traced profile build
tag=${1:-x}
python -m pytest tests/test_hot_engine_gpu.py -x -q > gpurun_out/ch_${tag}_tests.log 2>&1; tail -2 gpurun_out/ch_${tag}_tests.log
python tools/prof_chain.py 2>&1 | tail -1
python tools/prof_hot.py --phases 1 --epochs 3 2>&1 | grep "epoch ms" | tail -1 | cut -c1-30
python tools/prof_hot.py --phases 2 --epochs 3 2>&1 | grep "epoch ms" | tail -1 | cut -c1-30
MFK_RING_PROFILE=1 python -m matrix_factorization_b200.build --force > /dev/null 2>&1
python tools/ring_stats.py --workload ml-20m 2>&1 | grep -E "hot worker 0|trace warp" | tee gpurun_out/ch_${tag}_trace.log
