"""conditional FFTSIM 256^3, 1,000 data: time of gsp_fft_plan_condition (two Krigings) and of a conditional chunk (development tool)"""
import os, sys, time
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
vals = rng.standard_normal(nd)
for rep in range(2):
    lib.profile_enable(True)
    t = time.time(); plan.condition(0.0, cent, vals, knodes0 + 1); torch.cuda.synchronize(); wall = time.time() - t
    prof = lib.profile_read(); lib.profile_enable(False)
    print(f"condition wall {wall * 1e3:.1f} ms", {k: round(v["ms"], 2) for k, v in prof.items() if "krige" in k}, flush=True)
z = torch.empty((R, N), dtype=torch.float64, device=dev)
for _ in range(2):
    plan.sample_dev(R, None, 3, 0, 1.0, 0.0, 0, None, z.data_ptr())
print("conditional chunk of 32: ms", lib.last_sample_ms(), flush=True)
zc = z[:2].cpu().numpy()
print("data honoured:", float(np.abs(zc[:, knodes0] - vals[None, :]).max()), "checksum", float(np.abs(zc).sum()))
