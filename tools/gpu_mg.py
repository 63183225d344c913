"""Multi-GPU LUSIM check (development tool): block-cyclic factorization over all visible GPUs vs one GPU."""
import os, sys, time, math
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import iso
import torch
G = torch.cuda.device_count()
print("GPUs:", G, flush=True)
lib1 = gsp.Library(devices=[0])
libG = gsp.Library(devices=list(range(G)))
for dims, nd, kind, R in (((128, 128), 1000, O.EXPONENTIAL, 1000), ((256, 128), 500, O.SPHERICAL, 1024)):
    N = dims[0] * dims[1]
    rng = np.random.default_rng(3)
    dinds = np.sort(rng.choice(N, nd, replace=False)); z1 = rng.standard_normal(nd) * 0.4
    dom = (gsp._lib.make_grid_domain(dims, [0.0, 0.0], [1.0, 1.0]), None)
    st = iso(kind, 1.0, 20.0, 2)
    res = {}
    for name, lib in (("1gpu", lib1), (f"{G}gpu", libG)):
        gsp.LUPlan(lib, st, dom, dinds + 1, z1, 0.0).close()
        t = time.time(); plan = gsp.LUPlan(lib, st, dom, dinds + 1, z1, 0.0); wall = time.time() - t
        tm = plan.times()
        hZ = np.empty((N, R), order="F")
        t = time.time(); plan.sample(R, None, seed=7, out=hZ); ws = time.time() - t
        t = time.time(); plan.sample(R, None, seed=7, out=hZ); ws = time.time() - t
        res[name] = hZ.copy()
        print(f"N={N} {name}: plan wall {wall*1e3:.1f} ms, factor {tm[1]:.1f} ms ({(plan.Ns + (nd+127)//128*128)**3/3/tm[1]/1e9:.1f} TF/s), sample(host, R={R}) wall {ws*1e3:.1f} ms", flush=True)
        plan.close()
    a, b = res["1gpu"], res[f"{G}gpu"]
    print(f"   max rel diff 1 vs {G} GPUs: {np.abs(a-b).max()/np.abs(a).max():.2e}; data exact: {np.array_equal(b[dinds], np.repeat(z1[:,None], R, 1))}", flush=True)
print("done")
