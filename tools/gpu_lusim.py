"""LUSIM C3-size run for profiling (development tool)."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gsp_b200 as gsp
import torch
lib = gsp.Library()
A = np.zeros((3, 3)); A[0, 0] = A[1, 1] = 1 / 20.0
st = [(2, 1.0, A)]
N = 16384
rng = np.random.default_rng(3)
dinds = np.sort(rng.choice(N, 1000, replace=False)); z1 = rng.standard_normal(1000) * 0.5
dom = (gsp._lib.make_grid_domain((128, 128), [0.0, 0.0], [1.0, 1.0]), None)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(reps):
    t = time.time(); plan = gsp.LUPlan(lib, st, dom, dinds + 1, z1, 0.0); print("plan s", time.time() - t, "stage ms (assemble, factor, solve)", plan.times(), flush=True)
lib.profile_enable(True)
p2 = gsp.LUPlan(lib, st, dom, dinds + 1, z1, 0.0)
print("kernel ms:", {k: (round(v["ms"], 2), v["launches"]) for k, v in lib.profile_read().items()}, "factor ms", p2.times()[1], flush=True)
lib.profile_enable(False)
p2.close()
R = 1000
dev = torch.device("cuda:0")
W = torch.randn((R, plan.Ns), dtype=torch.float64, device=dev); Z = torch.empty((R, N), dtype=torch.float64, device=dev)
for _ in range(2):
    plan.sample_dev(R, W.data_ptr(), plan.Ns, 0, 0, 0, math.nan, None, Z.data_ptr(), N); print("sample ms", lib.last_sample_ms(), flush=True)
