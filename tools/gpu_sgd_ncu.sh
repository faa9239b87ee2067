#!/bin/bash
# ncu --set full captures of the two SGD kernels of an epoch at the ML-20M shape (serialised by ncu: the side-by-side hot launches
# appear one after the other)
R=${1:-r02b}
python tools/prof_hot.py --phases 7 --epochs 2 > /dev/null 2>&1
for kk in k_sgd_batch k_sgd_flat; do
  ncu --set full --clock-control none --import-source on -k regex:$kk -s 1 -c 1 -f -o gpurun_out/${R}_full_$kk python tools/prof_hot.py --phases 7 --epochs 2 > gpurun_out/${R}_full_$kk.log 2>&1
  ncu -i gpurun_out/${R}_full_$kk.ncu-rep --page details > gpurun_out/${R}_full_$kk.txt 2>/dev/null
  ncu -i gpurun_out/${R}_full_$kk.ncu-rep --page raw --csv > gpurun_out/${R}_full_${kk}_raw.csv 2>/dev/null
  tail -2 gpurun_out/${R}_full_$kk.log
done
