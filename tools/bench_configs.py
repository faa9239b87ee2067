"""BASELINE.json configs C1-C3 through the public estimator API, with the CPU arm (fp64 oracle port of the reference loops,
one host core) beside each line.  One JSON line per (config, model); written to stdout and, with --out, appended to a file.

    python tools/bench_configs.py [--out profiles/r02_configs.jsonl] [--only c1,c2,c3]

C1  KernelMF linear F=100, 20 epochs, synthetic ML-100K shape          (lr 0.001, reg 0.005)
C2  BaselineModel sgd (lr 0.01, reg 0.005) and als (reg 0.5), 20 epochs, ML-1M shape
C3  KernelMF sigmoid (lr 0.01) / rbf (lr 0.5, gamma 0.01) F=100, reg 0.005, ML-1M shape: fit on the initial users, then
    update_users(lr 0.001, 20 epochs) for 1000 held-out users (train_update_test_split)
Working sets fit in L2 (<= 12 MB): the lines report time and updates/s; a roofline fraction is not meaningful here
 and is not claimed.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ORACLE_FAST", "1")


def _rmse(pred, y):
    return float(np.sqrt(np.mean((np.asarray(pred) - np.asarray(y)) ** 2)))


def _line(config, model, n_updates, fit_s, extra, cpu_rate, cpu_sample):
    return {"metric": "rating-updates/s through the estimator API (fit: host preprocessing + H2D + plan + epochs + RMSE + D2H)",
            "config": config, "model": model, "value": n_updates / fit_s, "unit": "rating-updates/s", "fit_seconds": fit_s,
            "updates": int(n_updates), "dtype": "f32", "data": "synthetic", **extra,
            "cpu_baseline": {"value": cpu_rate, "unit": "rating-updates/s", "cores": 1, "kind": "port", "sample": cpu_sample},
            "speedup_vs_cpu_port": (n_updates / fit_s) / cpu_rate if cpu_rate else None}


def main():
    import torch
    import matrix_factorization_b200 as mfb
    from matrix_factorization_b200.data import SHAPES, split_rows, synth_ratings
    from oracle import oracle as orc

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="c1,c2,c3")
    args = ap.parse_args()
    only = set(args.only.split(","))
    lines = []

    def emit(d):
        lines.append(d)
        print(json.dumps(d), flush=True)

    def frames(shape, seed):
        U, I, N, step, mpu = SHAPES[shape]
        df = synth_ratings(U, I, N, seed=seed, grid_step=step, min_per_user=mpu)
        return split_rows(df, 0.1, seed=0)

    def internal(model, df):
        u = df["user_id"].map(model.user_id_map).to_numpy()
        i = df["item_id"].map(model.item_id_map).to_numpy()
        return u, i, df["rating"].to_numpy(dtype=np.float64)

    warm = synth_ratings(300, 200, 8000, seed=1, min_per_user=5)
    mfb.KernelMF(n_factors=8, n_epochs=1, verbose=0).fit(warm[["user_id", "item_id"]], warm["rating"])  # context, module load
    orc.lib()

    if "c1" in only:
        tr, te = frames("ml-100k", 1001)
        np.random.seed(1)
        m = mfb.KernelMF(n_factors=100, n_epochs=20, lr=0.001, reg=0.005, verbose=0)
        t0 = time.perf_counter()
        m.fit(tr[["user_id", "item_id"]], tr["rating"])
        dt = time.perf_counter() - t0
        test_rmse = _rmse(m.predict(te[["user_id", "item_id"]]), te["rating"])
        u, i, r = internal(m, tr)
        rng = np.random.default_rng(0)
        P, Q = rng.normal(0, 0.1, (m.n_users, 100)), rng.normal(0, 0.1, (m.n_items, 100))
        t0 = time.perf_counter()
        *_, rm = orc.kmf_sgd("linear", u, i, r, m.global_mean, np.zeros(m.n_users), np.zeros(m.n_items), P, Q, 20, 0.001, 0.005)
        cdt = time.perf_counter() - t0
        emit(_line("C1 KernelMF linear n_factors=100 n_epochs=20, synthetic ML-100K shape", "KernelMF(linear)", len(tr) * 20, dt,
                   {"train_rmse_last": m.train_rmse[-1], "test_rmse": test_rmse, "cpu_port_train_rmse_last": float(rm[-1])},
                   len(tr) * 20 / cdt, f"the same fit, all 20 epochs ({cdt:.1f} s)"))

    if "c2" in only:
        tr, te = frames("ml-1m", 1002)
        for method, kw in (("sgd", dict(lr=0.01, reg=0.005)), ("als", dict(reg=0.5))):
            np.random.seed(2)
            m = mfb.BaselineModel(method=method, n_epochs=20, verbose=0, **kw)
            t0 = time.perf_counter()
            m.fit(tr[["user_id", "item_id"]], tr["rating"])
            dt = time.perf_counter() - t0
            test_rmse = _rmse(m.predict(te[["user_id", "item_id"]]), te["rating"])
            u, i, r = internal(m, tr)
            t0 = time.perf_counter()
            if method == "sgd":
                *_, rm = orc.bias_sgd(u, i, r, m.global_mean, np.zeros(m.n_users), np.zeros(m.n_items), 20, 0.01, 0.005)
            else:
                *_, rm = orc.bias_als(u, i, r, m.global_mean, m.n_users, m.n_items, 20, 0.5)
            cdt = time.perf_counter() - t0
            emit(_line("C2 BaselineModel n_epochs=20, synthetic ML-1M shape", f"BaselineModel({method})", len(tr) * 20, dt,
                       {"train_rmse_last": m.train_rmse[-1], "test_rmse": test_rmse, "cpu_port_train_rmse_last": float(np.asarray(rm)[-1])},
                       len(tr) * 20 / cdt, f"the same fit, all 20 epochs ({cdt:.2f} s)"))

    if "c3" in only:
        U, I, N, step, mpu = SHAPES["ml-1m"]
        df = synth_ratings(U, I, N, seed=1003, grid_step=step, min_per_user=mpu)
        np.random.seed(3)
        Xi, yi, Xu, yu, Xt, yt = mfb.train_update_test_split(df, frac_new_users=1000 / U)
        for kernel, kw in (("sigmoid", dict(lr=0.01)), ("rbf", dict(lr=0.5, gamma=0.01))):
            np.random.seed(4)
            m = mfb.KernelMF(n_factors=100, n_epochs=20, kernel=kernel, reg=0.005, verbose=0, **kw)
            t0 = time.perf_counter()
            m.fit(Xi, yi)
            dt = time.perf_counter() - t0
            t0 = time.perf_counter()
            m.update_users(Xu, yu, lr=0.001, n_epochs=20, verbose=0)
            dtu = time.perf_counter() - t0
            test_rmse = _rmse(m.predict(Xt), yt)
            # CPU arm: 3 epochs of the same fit (bounded), extrapolated per update
            tri = Xi.assign(rating=yi)
            u, i, r = internal(m, tri)
            rng = np.random.default_rng(0)
            nu = int(u.max()) + 1
            P, Q = rng.normal(0, 0.1, (nu, 100)), rng.normal(0, 0.1, (m.n_items, 100))
            t0 = time.perf_counter()
            orc.kmf_sgd(kernel, u, i, r, m.global_mean, np.zeros(nu), np.zeros(m.n_items), P, Q, 3, kw["lr"], 0.005,
                        gamma=kw.get("gamma", 0.01))
            cdt = time.perf_counter() - t0
            emit(_line("C3 KernelMF n_factors=100 n_epochs=20, synthetic ML-1M shape, fit + update_users for 1000 new users",
                       f"KernelMF({kernel})", len(Xi) * 20, dt,
                       {"train_rmse_last": m.train_rmse[-1], "test_rmse_after_update": test_rmse, "update_users_seconds": dtu,
                        "update_users_updates": int(len(Xu) * 20), "update_users_rate": len(Xu) * 20 / dtu},
                       len(Xi) * 3 / cdt, f"3 of the 20 epochs of the same fit ({cdt:.1f} s)"))
    if args.out:
        with open(os.path.join(ROOT, args.out), "a") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
