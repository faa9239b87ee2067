"""Diagnostics: per-worker cycle counters of the ring-DSGD kernel on a bench workload.
    python tools/ring_stats.py --workload ml-20m [--workers W --warps K] -> gpurun_out/ring_stats_<workload>.npz"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from matrix_factorization_b200 import engine

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ml-20m")
    ap.add_argument("--uniform", action="store_true")
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--tag", default="")
    ap.add_argument("--no-hot", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = bench.gen_workload(args.workload, dev, uniform=args.uniform)
    F, U, I, N = wl["F"], wl["U"], wl["I"], wl["N"]
    g = torch.Generator(device=dev).manual_seed(5)
    P = torch.randn(U, F, device=dev, generator=g) * 0.1
    Q = torch.randn(I, F, device=dev, generator=g) * 0.1
    bu, bi = torch.zeros(U, device=dev), torch.zeros(I, device=dev)
    mu = float(wl["r"].double().mean().item())
    plan = engine.Plan(wl["u"], wl["i"], wl["r"], U, I, n_factors=F, n_workers=args.workers, warps_per_cta=args.warps,
                       hot_min_degree=engine.Plan.NO_HOT_SPLIT if args.no_hot else 0)
    info = plan.info()
    w, s = plan.assignment()
    H = info["n_hot_workers"]
    HH = H + info["n_hot_user_workers"]
    w, s = w[w >= HH] - HH, s[w >= HH] - HH  # cold workers only (the workers / steps of the hot phases are numbered first)
    counts = torch.bincount(w.long(), minlength=info["n_workers"]).cpu().numpy()
    steps_nonempty = torch.unique(w.long() * 65536 + s.long()).div(65536, rounding_mode="floor").bincount(minlength=info["n_workers"]).cpu().numpy()
    ms = []
    for e in range(args.epochs):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        engine.kmf_sgd_epoch(plan, "linear", P, Q, bu, bi, F, mu, wl["lr"], wl["reg"], 1.0 / F, 0.0, 5.0,
                             upd_user=not os.environ.get("NO_UPD_USER"), upd_item=not os.environ.get("NO_UPD_ITEM"))
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    st = plan.stats()
    out = os.path.join(ROOT, "gpurun_out", f"ring_stats_{args.workload}{args.tag}.npz")
    np.savez(out, stats=st, counts=counts, steps_nonempty=steps_nonempty, ms=np.array(ms), info=str(info))
    busy = st[:, 0] - st[:, 1]
    print("info", info)
    print("epoch ms", ms)
    order = np.argsort(-counts)[:8]
    print("top workers by ratings: id, ratings, nonempty steps, total Mcyc, blocked Mcyc, busy cyc/rating, quads, singles")
    for x in order:
        print(x, counts[x], steps_nonempty[x], st[x, 0] / 1e6, st[x, 1] / 1e6, busy[x] / max(1, counts[x]), st[x, 2], st[x, 3])
    prof = plan.last_profile
    names = ["handoff", "prefetch", "itemswitch", "cpwait", "quadmath", "quadupd", "single", "n_par4"]
    if info.get("flat"):
        x = int(np.argmax(st[:, 0]))
        print("flat worker", x, "chunks", st[x, 2], "ratings", st[x, 3], "Mcyc", st[x, 0] / 1e6, "ring wait Mcyc", st[x, 1] / 1e6,
              "consumers 0..3: dependency wait Mcyc", [round(float(v) / 1e6, 2) for v in prof[x][:4]], "producer/row wait Mcyc", [round(float(v) / 1e6, 2) for v in prof[x][4:]],
              "cycles per rating", st[x, 0] / max(1, st[x, 3]))
        print("flat workers: Mcyc min/median/max", st[:, 0].min() / 1e6, np.median(st[:, 0]) / 1e6, st[:, 0].max() / 1e6)
    if prof.any():
        for x in list(order[:2]) + [np.argsort(counts)[len(counts) // 2]]:
            print("phase Mcyc worker", x, {n: round(float(v) / 1e6, 2) for n, v in zip(names, prof[x])})
    med = np.argsort(counts)[len(counts) // 2]
    print("median worker:", med, counts[med], steps_nonempty[med], st[med, 0] / 1e6, st[med, 1] / 1e6, busy[med] / max(1, counts[med]), st[med, 2], st[med, 3])
    print("busy cycles/rating percentiles (all workers):", np.percentile(busy / np.maximum(1, counts), [5, 50, 95]))
    print("sum busy Mcyc", busy.sum() / 1e6, "max total Mcyc", st[:, 0].max() / 1e6)
    for label, hs, hp in (("hot item", plan.hot_stats, plan.hot_profile), ("hot user", plan.hot_user_stats, plan.hot_user_profile)):
        if not len(hs):
            continue
        x = int(np.argmax(hs[:, 0]))
        print(label, "phase:", len(hs), "workers, max Mcyc", hs[:, 0].max() / 1e6, "ratings max", hs[:, 3].max(), "sum", hs[:, 3].sum(),
              "slowest: blocked Mcyc", hs[x, 1] / 1e6, "batches", hs[x, 2], "ratings", hs[x, 3])
    if H:
        hs, hp = plan.hot_stats, plan.hot_profile
        top = np.argsort(-hs[:, 3])[:3]
        print("hot items:", H, "max cycles (M)", hs[:, 0].max() / 1e6, "ratings max", hs[:, 3].max())
        names = ["t", "solve", "sums", "sweep", "fetch", "gram", "inverse", "sync_A"]
        for x in top:
            print("hot worker", x, "ratings", hs[x, 3], "batches", hs[x, 2], "Mcyc", hs[x, 0] / 1e6, "blocked", hs[x, 1] / 1e6,
                  {n: round(float(v) / 1e6, 2) for n, v in zip(names, hp[x])},
                  "cyc/batch", {n: int(v / max(1, hs[x, 2])) for n, v in zip(names, hp[x])})
    # one iteration of hot worker 0, warp by warp (profile builds): clock stamps relative to the earliest one
    import ctypes
    from matrix_factorization_b200 import _lib
    L = _lib.lib()
    if hasattr(L, "mfk_debug_trace"):
        buf = np.zeros(256, dtype=np.int64)
        L.mfk_debug_trace.restype = ctypes.c_int
        if L.mfk_debug_trace(buf.ctypes.data_as(ctypes.c_void_p), 256) == 0 and buf.any():
            t = buf.reshape(16, 16)
            t0 = t[t > 0].min()
            for wp in range(16):
                print("trace warp", wp, [int(v - t0) if v > 0 else -1 for v in t[wp, :16]])


if __name__ == "__main__":
    main()
