"""Quick on-GPU sanity + timing sweep (development tool, run under gpurun).  Not a test, not the bench."""
import importlib.util
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
spec = importlib.util.spec_from_file_location("gsplib", os.path.join(ROOT, "geostatsprocesses.jl_b200", "_lib.py"))
L = importlib.util.module_from_spec(spec)
spec.loader.exec_module(L)
import gsp_oracle as O  # noqa: E402
import scipy.linalg  # noqa: E402

out = {}


def relerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def iso(kind, sill, rang, ndim):
    A = np.zeros((3, 3))
    for a in range(ndim):
        A[a, a] = 1.0 / rang
    return [(kind, sill, A)]


lib = L.Library()
print(lib.version(), flush=True)
rng = np.random.default_rng(0)

# ---- potrf parity + timing
for n in ((300,) if '--small' in sys.argv else (300, 1024, 4096)):
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    t = time.time()
    Lc = lib.potrf(S)
    dt = time.time() - t
    Lr = scipy.linalg.cholesky(S, lower=True)
    out[f"potrf_{n}_err"] = relerr(Lc, Lr)
    print("potrf", n, out[f"potrf_{n}_err"], "wall %.3f" % dt, flush=True)

# ---- pairwise
X1 = rng.uniform(0, 30, (1000, 3))
st = iso(O.EXPONENTIAL, 1.0, 9.0, 3)
P = lib.pairwise(st, X1)
out["pairwise_err"] = relerr(P, O.pairwise([O.Structure(*s) for s in st], X1))
print("pairwise", out["pairwise_err"], flush=True)

# ---- LUSIM C1 parity (50x50 spherical range 20, 100 reals)
st = iso(O.SPHERICAL, 1.0, 20.0, 2)
ost = [O.Structure(*s) for s in st]
dims = (50, 50)
coords = O.grid_centroids(dims, [0, 0], [1, 1])
plan = L.LUPlan(lib, st, (L.make_grid_domain(dims, [0, 0], [1, 1]), None), None, None, 0.0)
pre = O.lusim_preprocess(ost, coords, np.zeros(0, dtype=np.int64), np.zeros(0), 0.0)
W = np.random.default_rng(1).standard_normal((2500, 100))
Z = plan.sample(100, W)
out["lusim_c1_err"] = relerr(Z, O.lusim_sample(pre, W))
print("lusim C1", out["lusim_c1_err"], flush=True)
plan.close()

# ---- LUSIM conditional mid-size parity
dims = (48, 40)
coords = O.grid_centroids(dims, [0, 0], [1, 1])
st = iso(O.EXPONENTIAL, 1.0, 10.0, 2)
ost = [O.Structure(*s) for s in st]
dinds = np.sort(rng.choice(coords.shape[0], 200, replace=False))
z1 = rng.standard_normal(200)
plan = L.LUPlan(lib, st, (L.make_grid_domain(dims, [0, 0], [1, 1]), None), dinds + 1, z1, 0.0)
pre = O.lusim_preprocess(ost, coords, dinds, z1, 0.0)
W = rng.standard_normal((plan.Ns, 300))
Z = plan.sample(300, W)
out["lusim_cond_err"] = relerr(Z, O.lusim_sample(pre, W))
out["lusim_cond_exact"] = bool(np.array_equal(Z[dinds], np.repeat(z1[:, None], 300, 1)))
print("lusim cond", out["lusim_cond_err"], out["lusim_cond_exact"], flush=True)
plan.close()

# ---- FFTSIM parity
for dims, kind, rang in (((64, 32), O.SPHERICAL, 9.0), ((32, 16, 8), O.EXPONENTIAL, 5.0), ((100, 100), O.SPHERICAL, 10.0),
                         ((256, 256), O.EXPONENTIAL, 30.0)):
    nd = len(dims)
    st = iso(kind, 1.5, rang, nd)
    ost = [O.Structure(*s) for s in st]
    plan = L.FFTPlan(lib, st, dims, [0.0] * nd, [1.0] * nd)
    Fo = O.fftsim_preprocess(ost, dims, [0.0] * nd, [1.0] * nd)
    eF = relerr(plan.spectrum(), Fo)
    w = rng.random((3, int(np.prod(dims))))
    Zc = plan.sample(3, w, sill=1.5, mu=0.2)
    Zo = np.stack([O.fftsim_sample(Fo, w[r], 1.5, 0.2) for r in range(3)])
    out[f"fft_{'x'.join(map(str, dims))}_err"] = relerr(Zc, Zo)
    print("fft", dims, "F", eF, "Z", relerr(Zc, Zo), flush=True)
    plan.close()

if '--small' in sys.argv:
    print(json.dumps(out))
    sys.exit(0)

# ---- timings with device-resident data
import torch  # noqa: E402

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)

# FP64 GEMM peak (cuBLAS) for the roofline denominator
a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["dgemm_8192_tflops"] = 2 * 8192**3 / best / 1e9
print("cuBLAS dgemm 8192 TF/s", out["dgemm_8192_tflops"], flush=True)
del a, b

# LUSIM C3: 128x128 grid, 1000 data, exponential range 20
dims = (128, 128)
st = iso(O.EXPONENTIAL, 1.0, 20.0, 2)
N = 16384
dinds = np.sort(np.random.default_rng(3).choice(N, 1000, replace=False))
z1 = np.random.default_rng(3).standard_normal(1000) * 0.5
for rep in range(2):
    t = time.time()
    plan = L.LUPlan(lib, st, (L.make_grid_domain(dims, [0, 0], [1, 1]), None), dinds + 1, z1, 0.0)
    dt = time.time() - t
    print("lusim C3 plan wall s", dt, flush=True)
    out["lusim_c3_plan_s"] = dt
    if rep == 0:
        plan.close()
R = 1000
Zd = torch.empty((R, N), dtype=torch.float64, device=dev)
for rep in range(3):
    plan.sample_dev(R, None, plan.Ns, 1234, 0, 0, math.nan, None, Zd.data_ptr(), N)
    print("lusim C3 sample_dev ms (device RNG incl.)", lib.last_sample_ms(), flush=True)
out["lusim_c3_sample_ms"] = lib.last_sample_ms()
Wd = torch.randn((R, plan.Ns), dtype=torch.float64, device=dev)
for rep in range(3):
    plan.sample_dev(R, Wd.data_ptr(), plan.Ns, 0, 0, 0, math.nan, None, Zd.data_ptr(), N)
    print("lusim C3 sample_dev ms (injected W)", lib.last_sample_ms(), flush=True)
out["lusim_c3_sample_inj_ms"] = lib.last_sample_ms()
Zh = Zd.cpu().numpy()
out["lusim_c3_data_exact"] = bool(np.array_equal(Zh[:, dinds], np.repeat(z1[None, :], R, 0)))
out["lusim_c3_var"] = float(Zh.var())
print("C3 exact", out["lusim_c3_data_exact"], "var", out["lusim_c3_var"], flush=True)
plan.close()
del Zd, Wd

# FFTSIM timings
for dims, kind, R in (((1024, 1024), O.GAUSSIAN, 64), ((256, 256, 256), O.SPHERICAL, 16)):
    nd = len(dims)
    N = int(np.prod(dims))
    if nd == 3:
        th = math.radians(30.0)
        Rm = np.array([[math.cos(th), -math.sin(th), 0], [math.sin(th), math.cos(th), 0], [0, 0, 1]])
        A = np.diag([1 / 40.0, 1 / 20.0, 1 / 10.0]) @ Rm.T
        st = [(kind, 1.0, A)]
    else:
        st = iso(kind, 1.0, 50.0, nd)
    t = time.time()
    plan = L.FFTPlan(lib, st, dims, [0.0] * nd, [1.0] * nd)
    print("fft plan", dims, "wall s", time.time() - t, flush=True)
    wd = torch.rand((R, N), dtype=torch.float64, device=dev)
    zd = torch.empty((R, N), dtype=torch.float64, device=dev)
    for rep in range(3):
        plan.sample_dev(R, wd.data_ptr(), 0, 0, 1.0, 0.0, 0, None, zd.data_ptr())
        ms = lib.last_sample_ms()
        print("fft sample_dev", dims, "R", R, "ms", ms, "real/s", R / ms * 1e3, "alg GB/s", 20 * N * R / ms / 1e6, flush=True)
    out[f"fft_{nd}d_ms_per_real"] = ms / R
    z = zd[0].cpu().numpy()
    print("  mean", z.mean(), "var(N-1)", (z * z).sum() / (N - 1), flush=True)
    out[f"fft_{nd}d_mean"] = float(z.mean())
    out[f"fft_{nd}d_var"] = float((z * z).sum() / (N - 1))
    plan.close()
    del wd, zd

print("launches", lib.kernel_launches())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
