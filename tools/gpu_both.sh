#!/bin/bash
tag=${1:-x}
timeout 900 python -m pytest tests/test_score_gpu.py tests/test_hot_engine_gpu.py tests/test_kmf_gpu.py -x -q > gpurun_out/bo_${tag}_tests.log 2>&1; tail -3 gpurun_out/bo_${tag}_tests.log
timeout 300 python tools/score_bench.py --workload ml-20m --all-users 2>&1 | tail -1 | cut -c88-140
timeout 300 python tools/score_bench.py --workload netflix --all-users 2>&1 | tail -1 | cut -c88-140
python tools/prof_chain.py 2>&1 | tail -1
python tools/prof_chain.py --factors 256 --users 200000 2>&1 | tail -1
python tools/prof_hot.py --phases 1 --epochs 3 2>&1 | grep "epoch ms" | tail -1 | cut -c1-30
python tools/prof_hot.py --phases 2 --epochs 3 2>&1 | grep "epoch ms" | tail -1 | cut -c1-30
python tools/prof_hot.py --phases 4 --epochs 3 2>&1 | grep "epoch ms" | tail -1 | cut -c1-30
