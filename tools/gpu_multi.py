"""Multi-device context check on a real box (development tool): one process drives 2 GPUs, realizations sharded."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import iso
lib1 = gsp.Library(devices=[0])
lib2 = gsp.Library(devices=[0, 1])
st = iso(O.SPHERICAL, 1.0, 9.0, 3)
dims = (64, 64, 32)
p1 = gsp.FFTPlan(lib1, st, dims, [0.0] * 3, [1.0] * 3)
p2 = gsp.FFTPlan(lib2, st, dims, [0.0] * 3, [1.0] * 3)
w = np.random.default_rng(0).random((7, int(np.prod(dims))))
assert np.array_equal(p1.sample(7, w), p2.sample(7, w)), "fft injected"
assert np.array_equal(p1.sample(7, None, seed=5), p2.sample(7, None, seed=5)), "fft rng"
dom = (gsp._lib.make_grid_domain((48, 40), (0, 0), (1, 1)), None)
rng = np.random.default_rng(1)
dinds = np.sort(rng.choice(1920, 100, replace=False)); z1 = rng.standard_normal(100)
q1 = gsp.LUPlan(lib1, iso(O.EXPONENTIAL, 1.0, 10.0, 2), dom, dinds + 1, z1, 0.0)
q2 = gsp.LUPlan(lib2, iso(O.EXPONENTIAL, 1.0, 10.0, 2), dom, dinds + 1, z1, 0.0)
W = rng.standard_normal((q1.Ns, 301))
assert np.array_equal(q1.sample(301, W), q2.sample(301, W)), "lu injected"
assert np.array_equal(q1.sample(301, None, seed=9), q2.sample(301, None, seed=9)), "lu rng"
# throughput of the host API with 2 devices vs 1 (pinned buffers)
import torch
dims = (256, 256, 256); N = 1 << 24; R = 8
hw = torch.rand((R, N), dtype=torch.float64).pin_memory(); hz = torch.empty((R, N), dtype=torch.float64).pin_memory()
for name, lib in (("1 GPU", lib1), ("2 GPUs", lib2)):
    plan = gsp.FFTPlan(lib, iso(O.SPHERICAL, 1.0, 20.0, 3), dims, [0.0] * 3, [1.0] * 3)
    for _ in range(2):
        t = time.time()
        lib.check(lib.lib.gsp_fft_sample(plan.h, R, hw.data_ptr(), 0, 0, 1.0, 0.0, 0, None, hz.data_ptr()))
        dt = time.time() - t
    print(name, "host-API FFTSIM 256^3:", R / dt, "realizations/s", flush=True)
    plan.close()
print("multi-device OK")
