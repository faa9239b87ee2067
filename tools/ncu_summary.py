"""Selected metrics of `ncu --page raw --csv` exports (one captured launch each), in the layout of profiles/*_ncu_full_summary.txt.
    python tools/ncu_summary.py title=file_raw.csv [title=file_raw.csv ...]"""
import csv
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
]
for arg in sys.argv[1:]:
    title, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"## {title} -- {vals[col['Kernel Name']]}")
    for k in sorted(KEEP):
        if k in col:
            print(f"  {k:<96}{units[col[k]]:<15}{vals[col[k]]}")
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
                h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio"):
            try:
                stalls.append((float(vals[i]), h.split("stalled_")[1].split("_per_issue")[0].replace(".ratio", "")))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    if stalls:
        print("  top stall reasons (warps stalled per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))
    print()
