"""Multi-device context check of resident ensembles, statistics and conditional FFTSIM on a real 2-GPU box (development tool)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import iso
lib1 = gsp.Library(devices=[0])
lib2 = gsp.Library(devices=[0, 1])
rng = np.random.default_rng(0)
# statistics on uploaded data: two devices == one device == numpy
n, R = 30_000, 513
Z = rng.standard_normal((R, n)) * 3.0 + 7.0
for lib in (lib1, lib2):
    e = gsp.DeviceEnsemble(lib, n, R); e.put(Z)
    assert np.array_equal(e.fetch(), Z)
    assert np.abs(e.mean() - Z.mean(axis=0)).max() < 1e-12
    assert np.abs(e.var() - Z.var(axis=0, ddof=1)).max() < 1e-11
    assert np.array_equal(e.cdf(7.5), (Z <= 7.5).sum(axis=0) / R) and np.array_equal(e.ccdf(7.5), (Z > 7.5).sum(axis=0) / R)
    q = e.quantile([0.0, 0.3, 0.5, 1.0])
    assert np.abs(q - np.quantile(Z, [0.0, 0.3, 0.5, 1.0], axis=0)).max() < 1e-13
    e.close()
print("ensemble statistics on 2 devices OK", flush=True)
# resident simulation, unconditional and conditional, 2 devices == 1 device (counter RNG: independent of the sharding)
st = iso(O.SPHERICAL, 1.0, 9.0, 3)
dims = (64, 64, 32); N = int(np.prod(dims))
kn = np.sort(rng.choice(N, 50, replace=False))
cent = np.stack([(kn % 64) + 0.5, ((kn // 64) % 64) + 0.5, (kn // 4096) + 0.5], axis=1)
dv = rng.standard_normal(50)
res = []
for lib in (lib1, lib2):
    p = gsp.FFTPlan(lib, st, dims, [0.0] * 3, [1.0] * 3)
    a = p.sample_ensemble(11, None, seed=5)
    p.condition(0.0, cent, dv, kn + 1)
    b = p.sample_ensemble(11, None, seed=5)
    res.append((a.fetch(), b.fetch(), b.mean(), b.quantile([0.5])[0]))
    assert np.abs(res[-1][1][:, kn] - dv[None, :]).max() < 1e-10
    a.close(); b.close(); p.close()
assert np.array_equal(res[0][0], res[1][0]), "unconditional resident"
assert np.array_equal(res[0][1], res[1][1]), "conditional resident"
assert np.abs(res[0][2] - res[1][2]).max() < 1e-13 and np.abs(res[0][3] - res[1][3]).max() < 1e-13
print("resident / conditional FFTSIM on 2 devices OK", flush=True)
# LUSIM resident on 2 devices
dom = (gsp._lib.make_grid_domain((48, 40), (0, 0), (1, 1)), None)
dinds = np.sort(rng.choice(1920, 100, replace=False)); z1 = rng.standard_normal(100)
outs = []
for lib in (lib1, lib2):
    q = gsp.LUPlan(lib, iso(O.EXPONENTIAL, 1.0, 10.0, 2), dom, dinds + 1, z1, 0.0)
    e = q.sample_ensemble(301, None, seed=9)
    outs.append((e.fetch(), e.mean(), e.var()))
    e.close(); q.close()
assert np.array_equal(outs[0][0], outs[1][0]) and np.abs(outs[0][1] - outs[1][1]).max() < 1e-13 and np.abs(outs[0][2] - outs[1][2]).max() < 1e-12
print("LUSIM resident on 2 devices OK", flush=True)
# big: 256^3 x 32 resident over 2 GPUs, statistics timing
p = gsp.FFTPlan(lib2, iso(O.SPHERICAL, 1.0, 20.0, 3), (256, 256, 256), [0.0] * 3, [1.0] * 3)
e = p.sample_ensemble(32, None, seed=1)
for name, fn in (("simulate 32", lambda: p.sample_ensemble(32, None, seed=2, ens=e)), ("mean", e.mean), ("var", e.var), ("quantile", lambda: e.quantile([0.5]))):
    t = time.time(); fn(); print(f"2 GPUs 256^3 x 32: {name} {1e3 * (time.time() - t):.1f} ms", flush=True)
print("multi-device ensembles OK")
