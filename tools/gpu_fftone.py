"""Run a few FFTSIM 256^3 realizations with the schedule given by the GSP_FFT_* environment (for ncu captures)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import aniso3
import torch
lib = gsp.Library()
dev = torch.device("cuda:0")
n = int(os.environ.get("ONE_N", "256")); dims = (n, n, n); N = n ** 3; R = int(os.environ.get("ONE_R", "4"))
st = aniso3(O.SPHERICAL, 1.0, (40.0, 20.0, 10.0), 30.0)
w = torch.rand((R, N), dtype=torch.float64, device=dev)
z = torch.empty((R, N), dtype=torch.float64, device=dev)
plan = gsp.FFTPlan(lib, st, dims, [0.0] * 3, [1.0] * 3)
if os.environ.get("ONE_PROFILE"):
    lib.profile_enable(True)
wp = None if os.environ.get("ONE_RNG") else w.data_ptr()
for _ in range(int(os.environ.get("ONE_REPS", "2"))):
    plan.sample_dev(R, wp, 0, 0, 1.0, 0.0, 0, None, z.data_ptr())
print("ms/real", lib.last_sample_ms() / R)
if os.environ.get("ONE_PROFILE"):
    for k, v in lib.profile_read().items():
        print(k, v)
plan.close()
