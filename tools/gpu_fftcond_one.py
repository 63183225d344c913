"""one conditional FFTSIM 256^3 chunk (for ncu captures of the conditioning kernels)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import aniso3
import torch
lib = gsp.Library(); dev = torch.device("cuda:0")
n = int(os.environ.get("ONE_N", "256")); dims = (n, n, n); N = n ** 3; nd = 1000; R = 32
rng = np.random.default_rng(1)
knodes0 = np.sort(rng.choice(N, nd, replace=False))
cent = np.stack([(knodes0 // (n ** a)) % n + 0.5 for a in range(3)], axis=1)
plan = gsp.FFTPlan(lib, aniso3(O.SPHERICAL, 1.0, (40.0, 20.0, 10.0), 30.0), dims, [0.0] * 3, [1.0] * 3)
plan.condition(0.0, cent, rng.standard_normal(nd), knodes0 + 1)
z = torch.empty((R, N), dtype=torch.float64, device=dev)
plan.sample_dev(R, None, 3, 0, 1.0, 0.0, 0, None, z.data_ptr())
print("ms", lib.last_sample_ms())
