"""Times the scoring + mask + top-k pass (recommend for many users) on a bench workload shape.
    python tools/score_bench.py --workload netflix --users 32768 --k 50"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(workload="netflix", users=32768, k=50, kernel="linear", reps=3, mask_per_user=200, seed=1):
    import torch
    from matrix_factorization_b200 import engine
    from matrix_factorization_b200.data import SHAPES
    import bench

    shape, F, *_ = bench.WORKLOADS[workload]
    U, I, N, _, _ = SHAPES[shape]
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(seed)
    ld = engine.round_up4(F)
    P = torch.zeros(U, ld, device=dev)
    P[:, :F] = torch.randn(U, F, device=dev, generator=g) * 0.1
    Q = torch.zeros(I, ld, device=dev)
    Q[:, :F] = torch.randn(I, F, device=dev, generator=g) * 0.1
    bu = torch.randn(U, device=dev, generator=g) * 0.1
    bi = torch.randn(I, device=dev, generator=g) * 0.1
    m = min(users, U)
    ulist = torch.randperm(U, device=dev, generator=g)[:m].int()
    # synthetic known-item lists: mask_per_user distinct sorted items per user
    mi = torch.rand(m, I, device=dev, generator=g).topk(mask_per_user, dim=1).indices.sort(dim=1).values.int().reshape(-1)
    mp = (torch.arange(m + 1, device=dev) * mask_per_user).long()
    out = None
    ts = []
    for r in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = engine.score_topk(kernel, ulist, P, Q, bu, bi, I, F, 3.5, 1.0 / F, 0.0, 5.0, k, True, mp, mi)
        b.record()
        torch.cuda.synchronize()
        if r > 0:
            ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    kp = (F + 31) // 32 * 32
    flop = 2.0 * m * I * kp
    return {"workload": workload, "users": m, "n_items": I, "n_factors": F, "k": k, "kernel": kernel, "ms": ms,
            "users_per_s": m / (ms * 1e-3), "tflops_fp32_equiv": flop / (ms * 1e-3) / 1e12,
            "tflops_tf32_issued": 3 * flop / (ms * 1e-3) / 1e12,
            "path": "simt" if os.environ.get("MFK_SCORE_SIMT") == "1" or k > 64 else "tcgen05 split-TF32"}


def run_workload(wl, k=50, kernel="linear", reps=3, seed=1, cpu_users=1000):
    """recommend for ALL users of a bench workload (bench.gen_workload dict): top-k with each user's TRAINING items as the
    known-item mask (CSR built on the GPU by mfk_csr_create) -- the second half of BASELINE.json's metric.  Next to it the
    reference-equivalent per-user loop on the host (predict all candidates, mask, sort, head: recommender_base.py:245-266)
    on `cpu_users` users, one core."""
    import torch
    from matrix_factorization_b200 import engine

    F, U, I = wl["F"], wl["U"], wl["I"]
    dev = wl["u"].device
    g = torch.Generator(device=dev).manual_seed(seed)
    ld = engine.round_up4(F)
    P = torch.zeros(U, ld, device=dev)
    P[:, :F] = torch.randn(U, F, device=dev, generator=g) * 0.1
    Q = torch.zeros(I, ld, device=dev)
    Q[:, :F] = torch.randn(I, F, device=dev, generator=g) * 0.1
    bu = torch.randn(U, device=dev, generator=g) * 0.1
    bi = torch.randn(I, device=dev, generator=g) * 0.1
    csr = engine.Csr(wl["u"], wl["i"], wl["r"], U, I)
    row_ptr, col, *_ = csr.export()
    csr.close()
    users = torch.arange(U, device=dev, dtype=torch.int32)
    ts = []
    out = None
    for r in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = engine.score_topk(kernel, users, P, Q, bu, bi, I, F, 3.5, 1.0 / F, 0.0, 5.0, k, True, row_ptr, col)
        b.record()
        torch.cuda.synchronize()
        if r > 0:
            ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    kp = (F + 31) // 32 * 32
    flop = 2.0 * U * I * kp
    # host loop, the reference's way: scores of all items for one user, known items out, full sort, head
    Ph, Qh = P[:cpu_users, :F].double().cpu().numpy(), Q[:, :F].double().cpu().numpy()
    buh, bih = bu[:cpu_users].double().cpu().numpy(), bi.double().cpu().numpy()
    rp, cl = row_ptr[: cpu_users + 1].cpu().numpy(), col[: int(row_ptr[cpu_users].item())].cpu().numpy()
    t0 = time.perf_counter()
    agree = 0
    for uu in range(cpu_users):
        sc = 3.5 + buh[uu] + bih + Qh @ Ph[uu]
        sc[cl[rp[uu]:rp[uu + 1]]] = -np.inf
        top = np.argsort(-sc, kind="stable")[:k]
        if uu < 50:
            agree += int(np.array_equal(top, out[1][uu].cpu().numpy()))
    cpu_dt = time.perf_counter() - t0
    return {"users": U, "n_items": I, "n_factors": F, "k": k, "kernel": kernel, "ms": ms, "users_per_s": U / (ms * 1e-3),
            "mask": f"each user's training items ({int(row_ptr[-1].item())} entries, CSR from mfk_csr_create)",
            "tflops_fp32_equiv": flop / (ms * 1e-3) / 1e12, "tflops_tf32_issued": 3 * flop / (ms * 1e-3) / 1e12,
            "path": "simt" if os.environ.get("MFK_SCORE_SIMT") == "1" or k > 64 else "tcgen05 split-TF32",
            "cpu_baseline": {"value": cpu_users / cpu_dt, "unit": "users/s", "cores": 1, "kind": "port",
                             "sample": f"{cpu_users} users, numpy fp64: all-item scores, mask, full sort, head ({cpu_dt:.2f} s)"},
            "lists_equal_cpu_first_50_users": agree}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="netflix")
    ap.add_argument("--users", type=int, default=32768)
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--kernel", default="linear")
    ap.add_argument("--all-users", action="store_true", help="all users of the workload, training items as the mask")
    a = ap.parse_args()
    if a.all_users:
        import torch
        import bench

        print(json.dumps(run_workload(bench.gen_workload(a.workload, torch.device("cuda", 0)), k=a.k, kernel=a.kernel)))
    else:
        print(json.dumps(run(a.workload, a.users, a.k, a.kernel)))
