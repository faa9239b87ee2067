#!/bin/bash
# scoring iteration loop: parity tests, then recommend for all users of both bench shapes
tag=${1:-x}
timeout 600 python -m pytest tests/test_score_gpu.py tests/test_api_gpu.py -x -q > gpurun_out/sc_${tag}_tests.log 2>&1; tail -3 gpurun_out/sc_${tag}_tests.log
timeout 300 python tools/score_bench.py --workload ml-20m --all-users 2>&1 | tail -1 | tee gpurun_out/sc_${tag}_ml20m.json | cut -c88-140
MFK_SCORE_ES=1 timeout 300 python tools/score_bench.py --workload ml-20m --all-users 2>&1 | tail -1 | cut -c88-140
MFK_SCORE_TS=0 timeout 300 python tools/score_bench.py --workload ml-20m --all-users 2>&1 | tail -1 | cut -c88-140
timeout 300 python tools/score_bench.py --workload netflix --all-users 2>&1 | tail -1 | tee gpurun_out/sc_${tag}_netflix.json | cut -c88-140
