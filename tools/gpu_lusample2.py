"""2-GPU check of the pipelined host-pointer LUSIM sampling: a 2-device context must return what a 1-device context returns (development tool)."""
import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import iso
import torch
G = torch.cuda.device_count()
dims = (64, 48); N = dims[0] * dims[1]; nd = 100; R = 2300
rng = np.random.default_rng(5)
dinds = np.sort(rng.choice(N, nd, replace=False)); z1 = rng.standard_normal(nd)
dom = (gsp._lib.make_grid_domain(dims, [0.0, 0.0], [1.0, 1.0]), None)
st = iso(O.SPHERICAL, 1.0, 10.0, 2)
out = {}
W = None
for name, devs in (("1", [0]), (str(G), list(range(G)))):
    lib = gsp.Library(devices=devs)
    plan = gsp.LUPlan(lib, st, dom, dinds + 1, z1, 0.0)
    if W is None:
        W = rng.standard_normal((plan.Ns, R)); W1 = rng.standard_normal((plan.Ns, R))
    out[name] = (plan.sample(R, W).copy(), plan.sample(R, W, rho=0.6, W1=W1).copy(), plan.sample(R, None, seed=11).copy())
    plan.close(); lib.close()
a, b = out["1"], out[str(G)]
print("GPUs", G, "equal (injected, mixed, device RNG):", [bool(np.array_equal(x, y)) for x, y in zip(a, b)],
      "max diff", [float(np.abs(x - y).max()) for x, y in zip(a, b)])
