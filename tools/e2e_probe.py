"""Diagnostics: the host-buffer call of bench.py's e2e leg, several times, with MFK_HOST_TIMING phases on stderr."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MFK_HOST_TIMING"] = "1"
import torch
import bench
wl = bench.gen_workload("ml-20m", torch.device("cuda", 0))
for k in range(2):
    print(bench.e2e_host_call(wl, wl["n_epochs"])[5], flush=True)
