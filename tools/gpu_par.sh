#!/bin/bash
tag=${1:-x}
timeout 900 python -m pytest tests/test_hot_engine_gpu.py tests/test_kmf_gpu.py tests/test_host_api_gpu.py tests/test_baseline_gpu.py tests/test_dist_gpu.py -x -q > gpurun_out/par_${tag}_tests.log 2>&1; tail -3 gpurun_out/par_${tag}_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/par_${tag}_ml20m.json 2> gpurun_out/par_${tag}_ml20m.err; tail -c 300 gpurun_out/par_${tag}_ml20m.err
MFK_HOT_PARALLEL=0 python bench.py --steps 5 --warmup 3 > gpurun_out/par_${tag}_ml20m_seq.json 2> /dev/null
python - <<PY
import json
for n in ["ml20m", "ml20m_seq"]:
    try:
        d = json.load(open(f"gpurun_out/par_${tag}_{n}.json"))
        print(n, round(d["value"] / 1e9, 3), round(d["ms_per_step"], 2), round(d["roofline"]["kernel_ms"], 2), round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 3), d.get("parity"), [ (k["kernel"][:22], round(k["ms"],2)) for k in d["roofline"]["per_kernel"]], {k: d["config"]["plan"][k] for k in ("n_hot_ratings","n_hot_user_ratings","n_hot_workers","n_hot_user_workers")})
    except Exception as e:
        print(n, "failed", e)
PY
