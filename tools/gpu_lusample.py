"""LUSIM host-pointer sampling: pipelined (default) vs stream-ordered (GSP_LU_PIPELINE=0), C3 and a 32k-node case (development tool;
run once per setting: the switch is read once per process)."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import iso
import torch
lib = gsp.Library()
for dims, nd, R in (((128, 128), 1000, 1000), ((256, 128), 500, 4096)):
    N = dims[0] * dims[1]
    rng = np.random.default_rng(3)
    dinds = np.sort(rng.choice(N, nd, replace=False)); z1 = rng.standard_normal(nd) * 0.5
    dom = (gsp._lib.make_grid_domain(dims, [0.0, 0.0], [1.0, 1.0]), None)
    plan = gsp.LUPlan(lib, iso(O.EXPONENTIAL, 1.0, 20.0, 2), dom, dinds + 1, z1, 0.0)
    hW = torch.randn((R, plan.Ns), dtype=torch.float64).pin_memory()
    hZ = torch.empty((R, N), dtype=torch.float64).pin_memory()
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        lib.check(lib.lib.gsp_lu_sample(plan.h, R, hW.data_ptr(), 0, 0, 0, math.nan, None, hZ.data_ptr()))
        best = min(best, time.perf_counter() - t)
    Zn = hZ.numpy()
    ok = bool(np.array_equal(Zn[:, dinds], np.repeat(z1[None, :], R, 0)))
    print(f"N={N} R={R}: host-pointer sample wall {best * 1e3:.1f} ms, device {lib.last_sample_ms():.1f} ms, data exact {ok}, checksum {float(np.abs(Zn).sum()):.6e}", flush=True)
    plan.close()
