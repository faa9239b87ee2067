"""At-scale parity of one epoch vs the fp64 oracle replay under the Gram-matrix variants of the batch engine.
    python tools/parity_scale.py [--n 2000000] [--factors 128]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2_000_000)
ap.add_argument("--factors", type=int, default=128)
ap.add_argument("--variants", default="auto,passes3,mma1")
ap.add_argument("--users", type=int, default=40_000)
ap.add_argument("--items", type=int, default=8_000)
args = ap.parse_args()
for name in args.variants.split(","):
    env = {"auto": {}, "passes3": {"MFK_HOT_GRAM_PASSES": "3"}, "mma1": {"MFK_HOT_GRAM": "mma", "MFK_HOT_GRAM_PASSES": "1"},
           "nohot": {}}[name]
    for k in ("MFK_HOT_GRAM_PASSES", "MFK_HOT_GRAM"):
        os.environ.pop(k, None)
    os.environ.update(env)
    p = bench.parity_at_scale(n_sample=args.n, F=args.factors, U=args.users, I=args.items)
    print(name, json.dumps({k: p[k] for k in ("rel_err_P", "rel_err_Q", "rel_err_update_P", "rel_err_update_Q", "max_abs_err_bi", "hot_ratings", "oracle_replay_s", "ok")}))
