"""Summarise an `ncu --page source --csv` export: stall reasons in total and the most-sampled SASS instructions.
    python tools/ncu_src_top.py file.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
print({hdr[i][6:]: sum(int(r[i] or 0) for r in data) for i in stall})
top = sorted(range(len(data)), key=lambda k: -int(data[k][isamp]))[:n_top]
for k in sorted(top):
    r = data[k]
    st = {hdr[i][6:]: int(r[i]) for i in stall if int(r[i] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(k, r[isamp], r[iex], r[isrc].strip()[:64], st)
