"""C3-size LUSIM plan: factorization time vs SMs kept free of look-ahead GEMMs (GSP_CHOL_RESERVE is read once per process ->
one subprocess per value).  Development tool."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import gsp_b200 as gsp
lib = gsp.Library()
A = np.zeros((3, 3)); A[0, 0] = A[1, 1] = 1 / 20.0
st = [(2, 1.0, A)]
for (dims, nd) in (((128, 128), 1000), ((256, 128), 500)):
    N = dims[0] * dims[1]
    rng = np.random.default_rng(3)
    dinds = np.sort(rng.choice(N, nd, replace=False)); z1 = rng.standard_normal(nd) * 0.5
    dom = (gsp._lib.make_grid_domain(dims, [0.0, 0.0], [1.0, 1.0]), None)
    best = 1e9
    for _ in range(3):
        plan = gsp.LUPlan(lib, st, dom, dinds + 1, z1, 0.0); best = min(best, plan.times()[1]); d2, _ = (None, None); plan.close()
    print(dims, "factor ms", round(best, 2), flush=True)
''' % ROOT
for r in sys.argv[1:] or ["0", "8", "16", "24", "32", "48"]:
    env = dict(os.environ, GSP_CHOL_RESERVE=r)  # GSP_CHOL_PANELS etc. pass through from the caller's environment
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    print("reserve", r, "|", " | ".join(out.stdout.strip().splitlines()), out.stderr[-300:] if out.returncode else "", flush=True)
