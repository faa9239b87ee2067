#!/bin/bash
# Final evidence of the round, part $1 (a: tests + bench lines, b: ncu launch list + full capture of the scoring kernel)
R=r02b
if [ "$1" = a ]; then
  python -m pytest tests -m gpu -q > gpurun_out/${R}_gputests.log 2>&1; tail -2 gpurun_out/${R}_gputests.log
  python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench_ml20m.json 2> gpurun_out/${R}_bench_ml20m.err; tail -c 200 gpurun_out/${R}_bench_ml20m.err
  python bench.py --steps 5 --warmup 3 --uniform > gpurun_out/${R}_bench_ml20m_uniform.json 2> /dev/null
  python bench.py --workload netflix --steps 3 --warmup 3 > gpurun_out/${R}_bench_netflix.json 2> gpurun_out/${R}_bench_netflix.err; tail -c 200 gpurun_out/${R}_bench_netflix.err
  python - <<'PY'
import json
for n in ["ml20m", "ml20m_uniform", "netflix"]:
    try:
        d = json.load(open(f"gpurun_out/r02b_bench_{n}.json"))
        print(n, round(d["value"] / 1e9, 3), round(d["ms_per_step"], 2), round(d["roofline"]["kernel_ms"], 2), round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 3), "rec", round(d["recommend"]["value"] / 1e6, 2), round(d["recommend"]["ms"], 1))
    except Exception as e:
        print(n, "failed", e)
PY
else
  python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> /dev/null; cut -c1-200 gpurun_out/${R}_bench_reference.json
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/${R}_launches_bench_ml20m.csv \
      python bench.py --steps 2 --warmup 1 > gpurun_out/${R}_ncu_bench.log 2>&1
  bash tools/gpu_score_ncu.sh ${R}
fi
