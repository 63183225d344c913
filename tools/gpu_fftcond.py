"""Conditional FFTSIM timing (development tool): plan conditioning (weight tables) and per-realization cost vs unconditional."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import aniso3, iso
import torch
lib = gsp.Library()
dev = torch.device("cuda:0")
for dims, nd, R in (((1024, 1024), 1000, 64), ((256, 256, 256), 1000, 32)):
    ndim = len(dims); N = int(np.prod(dims))
    st = aniso3(O.SPHERICAL, 1.0, (40.0, 20.0, 10.0), 30.0) if ndim == 3 else iso(O.GAUSSIAN, 1.0, 50.0, 2)
    rng = np.random.default_rng(1)
    knodes0 = np.sort(rng.choice(N, nd, replace=False))
    cent = np.stack([(knodes0 // int(np.prod(dims[:a]))) % dims[a] + 0.5 for a in range(ndim)], axis=1)
    dvals = rng.standard_normal(nd)
    plan = gsp.FFTPlan(lib, st, dims, [0.0] * ndim, [1.0] * ndim)
    z = torch.empty((R, N), dtype=torch.float64, device=dev)
    for _ in range(2):
        plan.sample_dev(R, None, 3, 0, 1.0, 0.0, 0, None, z.data_ptr())
    t_unc = lib.last_sample_ms()
    lib.profile_enable(True)
    t0 = time.perf_counter(); plan.condition(0.0, cent, dvals, knodes0 + 1); t_cond = time.perf_counter() - t0
    for _ in range(2):
        plan.sample_dev(R, None, 3, 0, 1.0, 0.0, 0, None, z.data_ptr())
    t_con = lib.last_sample_ms()
    prof = lib.profile_read(); lib.profile_enable(False)
    zz = z[0].cpu().numpy()
    print(dims, "nd", nd, f"condition {t_cond * 1e3:.1f} ms | per realization (device RNG): unconditional {t_unc / R * 1e3:.1f} us, conditional {t_con / R * 1e3:.1f} us",
          "| data honoured", float(np.abs(zz[knodes0] - dvals).max()), flush=True)
    print({k: (round(v["ms"] / v["launches"], 3), v["launches"]) for k, v in prof.items() if k.startswith("krige")}, flush=True)
    plan.close()
