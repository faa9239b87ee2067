#!/bin/bash
# scoring iteration loop: parity tests, then recommend for all users of both bench shapes
tag=${1:-x}
python -m pytest tests/test_score_gpu.py tests/test_api_gpu.py tests/test_hot_engine_gpu.py tests/test_kmf_gpu.py -x -q > gpurun_out/sc_${tag}_tests.log 2>&1; tail -3 gpurun_out/sc_${tag}_tests.log
python tools/score_bench.py --workload ml-20m --all-users 2>&1 | tail -1 | tee gpurun_out/sc_${tag}_ml20m.json | cut -c1-200
python tools/score_bench.py --workload netflix --all-users 2>&1 | tail -1 | tee gpurun_out/sc_${tag}_netflix.json | cut -c1-200
