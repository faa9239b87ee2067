#!/bin/bash
timeout 900 python -m pytest tests/test_kmf_gpu.py tests/test_hot_engine_gpu.py tests/test_host_api_gpu.py -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/flat_f1.json 2> gpurun_out/flat_f1.err; tail -c 300 gpurun_out/flat_f1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/flat_f1.json"))
print(round(d["value"] / 1e9, 3), round(d["ms_per_step"], 2), round(d["roofline"]["kernel_ms"], 2), round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 3), [round(k["ms"], 2) for k in d["roofline"]["per_kernel"]], d["parity"]["rel_err_P"], d["parity"]["rel_err_Q"], d["parity"]["ok"])
PY
python bench.py --workload netflix --steps 3 --warmup 3 > gpurun_out/flat_f1_netflix.json 2> /dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/flat_f1_netflix.json"))
print(round(d["value"] / 1e9, 3), round(d["ms_per_step"], 2), round(d["roofline"]["kernel_ms"], 2), round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 3), [round(k["ms"], 2) for k in d["roofline"]["per_kernel"]], d["parity"]["rel_err_P"], d["parity"]["rel_err_Q"], d["parity"]["ok"])
PY
