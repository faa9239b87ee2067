#!/bin/bash
# ncu --set full capture of the scoring kernel (recommend for all users of the ML-20M shape), with the source page
R=${1:-r02b}
python tools/score_bench.py --workload ml-20m --all-users > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_tc -s 1 -c 1 -f -o gpurun_out/${R}_full_k_score_tc python tools/score_bench.py --workload ml-20m --all-users > gpurun_out/${R}_full_k_score_tc.log 2>&1
ncu -i gpurun_out/${R}_full_k_score_tc.ncu-rep --page details > gpurun_out/${R}_full_k_score_tc.txt 2>/dev/null
ncu -i gpurun_out/${R}_full_k_score_tc.ncu-rep --page raw --csv > gpurun_out/${R}_full_k_score_tc_raw.csv 2>/dev/null
ncu -i gpurun_out/${R}_full_k_score_tc.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${R}_k_score_tc_lines.csv 2>/dev/null
tail -3 gpurun_out/${R}_full_k_score_tc.log
