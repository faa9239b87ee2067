#!/bin/bash
# batch-engine iteration loop: parity tests, isolated chain, phase timings
tag=${1:-x}
python -m pytest tests/test_hot_engine_gpu.py tests/test_kmf_gpu.py -x -q > gpurun_out/it_${tag}_tests.log 2>&1; tail -3 gpurun_out/it_${tag}_tests.log
python tools/prof_chain.py 2>&1 | tail -1
python tools/prof_chain.py --factors 256 --users 200000 2>&1 | tail -1
python tools/prof_hot.py --phases 1 --epochs 3 2>&1 | grep "epoch ms" | tail -1 | cut -c1-30
python tools/prof_hot.py --phases 2 --epochs 3 2>&1 | grep "epoch ms" | tail -1 | cut -c1-30
python tools/prof_hot.py --phases 4 --epochs 3 2>&1 | grep "epoch ms" | tail -1 | cut -c1-30
