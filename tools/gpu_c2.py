"""C2 (FFTSIM 2-D 1024x1024 GaussianCovariance(range=50), 64 realizations) device-resident timing (development tool)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import iso
import torch
lib = gsp.Library()
dev = torch.device("cuda:0")
for dims, R in (((1024, 1024), 64), ((4096, 4096), 16), ((256, 256), 256)):
    nd = len(dims); N = int(np.prod(dims))
    plan = gsp.FFTPlan(lib, iso(O.GAUSSIAN, 1.0, 50.0, nd), dims, [0.0] * nd, [1.0] * nd)
    w = torch.rand((R, N), dtype=torch.float64, device=dev); z = torch.empty((R, N), dtype=torch.float64, device=dev)
    for _ in range(3):
        plan.sample_dev(R, w.data_ptr(), 0, 0, 1.0, 0.0, 0, None, z.data_ptr())
    ms = lib.last_sample_ms()
    zz = z[R - 1]
    print(dims, "R", R, "ms", round(ms, 3), "real/s", round(R / ms * 1e3), "alg GB/s", round(20 * N * R / ms / 1e6), "mean", float(zz.mean()), "var", float((zz * zz).sum() / (N - 1)), flush=True)
    plan.close()
