#!/bin/bash
# Round evidence: plain runs first, then the ncu launch list of the bench command and --set full captures of the three
# dominant kernels (SGD hot phase, SGD flat phase, scoring).  Outputs land in gpurun_out/ (copied to profiles/ by hand).
set -x
R=${1:-r02}
python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench_ml20m.json 2> gpurun_out/${R}_bench_ml20m.err || exit 1
tail -c 300 gpurun_out/${R}_bench_ml20m.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/${R}_launches_bench_ml20m.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/${R}_ncu_bench.log 2>&1
python tools/prof_hot.py --phases 7 --epochs 2 > /dev/null 2>&1
for kk in k_sgd_batch k_sgd_flat; do
  ncu --set full --clock-control none --import-source on -k regex:$kk -s 1 -c 1 -f -o gpurun_out/${R}_full_$kk python tools/prof_hot.py --phases 7 --epochs 2 > gpurun_out/${R}_full_$kk.log 2>&1
  ncu -i gpurun_out/${R}_full_$kk.ncu-rep --page details > gpurun_out/${R}_full_$kk.txt 2>/dev/null
  ncu -i gpurun_out/${R}_full_$kk.ncu-rep --page raw --csv > gpurun_out/${R}_full_${kk}_raw.csv 2>/dev/null
done
python tools/score_bench.py --workload ml-20m --users 37888 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_tc -s 1 -c 1 -f -o gpurun_out/${R}_full_k_score_tc python tools/score_bench.py --workload ml-20m --users 37888 > gpurun_out/${R}_full_k_score_tc.log 2>&1
ncu -i gpurun_out/${R}_full_k_score_tc.ncu-rep --page details > gpurun_out/${R}_full_k_score_tc.txt 2>/dev/null
ncu -i gpurun_out/${R}_full_k_score_tc.ncu-rep --page raw --csv > gpurun_out/${R}_full_k_score_tc_raw.csv 2>/dev/null
ls -la gpurun_out/${R}_*
