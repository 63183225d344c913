"""Resident-ensemble statistics timing (development tool): FFTSIM 256^3 x R realizations, each statistic timed twice."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import aniso3
lib = gsp.Library()
n = int(os.environ.get("ENS_N", "256")); R = int(os.environ.get("ENS_R", "64"))
plan = gsp.FFTPlan(lib, aniso3(O.SPHERICAL, 1.0, (40.0, 20.0, 10.0), 30.0), (n, n, n), [0.0] * 3, [1.0] * 3)
t0 = time.perf_counter(); ens = plan.sample_ensemble(R, None, seed=4); print("simulate", time.perf_counter() - t0, flush=True)
t0 = time.perf_counter(); ens2 = plan.sample_ensemble(R, None, seed=5); print("simulate again", time.perf_counter() - t0, flush=True)
ens2.close()
lib.profile_enable(True)
for name, fn in (("mean", ens.mean), ("var", ens.var), ("cdf", lambda: ens.cdf(0.0)), ("ccdf", lambda: ens.ccdf(0.0)),
                 ("quantile1", lambda: ens.quantile([0.5])), ("quantile3", lambda: ens.quantile([0.1, 0.5, 0.9]))):
    for rep in range(2):
        t0 = time.perf_counter(); out = fn(); dt = time.perf_counter() - t0
        print(f"{name} call {rep}: {dt * 1e3:.1f} ms  (result mean {float(np.mean(out)):.4f})", flush=True)
print(lib.profile_read())
