"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per source line: instructions executed and
stall samples.   python tools/ncu_lines.py file.csv [min_instr]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
min_inst = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
cur_file, cur_line = "", 0
agg = defaultdict(lambda: [0, 0])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) < 8:
        continue
    if r[0].strip().isdigit():
        cur_line = int(r[0])
        continue
    if r[0] == "" and r[2].startswith("0x"):
        try:
            agg[(cur_file, cur_line)][0] += int(r[7])
            agg[(cur_file, cur_line)][1] += int(r[6])
        except ValueError:
            pass
tot_i = sum(v[0] for v in agg.values())
tot_s = sum(v[1] for v in agg.values())
print("total warp instructions", tot_i, "samples", tot_s)
byfile = defaultdict(lambda: [0, 0])
for (f, l), v in agg.items():
    byfile[f][0] += v[0]
    byfile[f][1] += v[1]
print({f: v for f, v in byfile.items()})
for (f, l), v in sorted(agg.items()):
    if v[0] >= min_inst or v[1] >= 15:
        print(f"{f}:{l}  instr {v[0]}  samples {v[1]}")
