"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs; golden fixtures; and size-independent properties at BASELINE.json's full sizes.
Tolerance: north_star's 1e-9 relative (normwise max|dZ|/max|Z|) in Float64; conditioning data bit-exact."""
import math
import os

import numpy as np
import pytest
import scipy.linalg

import gsp_b200 as gsp
import gsp_oracle as O
from helpers import aniso3, iso, ostructs, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def grid_dom(dims):
    nd = len(dims)
    return (gsp._lib.make_grid_domain(dims, [0.0] * nd, [1.0] * nd), None)


# ------------------------------------------------------------------ a1 / a2 pieces
def test_pairwise_models(gpu_lib):
    rng = np.random.default_rng(0)
    for kind in (O.SPHERICAL, O.EXPONENTIAL, O.GAUSSIAN, O.CUBIC, O.PENTASPHERICAL, O.SINEHOLE, O.CIRCULAR):
        st = aniso3(kind, 0.9, (9.0, 4.0, 2.0), 30.0) + [(O.NUGGET, 0.1, np.eye(3))]
        X1, X2 = rng.uniform(0, 12, (301, 3)), rng.uniform(0, 12, (77, 3))
        assert relerr(gpu_lib.pairwise(st, X1, X2), O.pairwise(ostructs(st), X1, X2)) < 1e-13
        assert relerr(gpu_lib.pairwise(st, X1), O.pairwise(ostructs(st), X1)) < 1e-13
    # Matern (K_nu evaluated on the device) against SciPy's kv: absolute and entry-wise relative agreement
    X = np.concatenate([rng.uniform(0, 12, (300, 3)), rng.uniform(0, 0.05, (20, 3)), rng.uniform(0, 200, (20, 3))])
    for nu in (0.3, 0.5, 1.0, 1.5, 2.5, 3.2, 7.0):
        st = aniso3(O.MATERN, 0.9, (9.0, 4.0, 2.0), 30.0, order=nu) + [(O.NUGGET, 0.1, np.eye(3))]
        G, Go = gpu_lib.pairwise(st, X), O.pairwise(ostructs(st), X)
        assert np.abs(G - Go).max() < 1e-13, nu
        big = Go > 1e-200
        assert np.abs(G[big] / Go[big] - 1).max() < 2e-12, nu


@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 1000, 2048])
def test_potrf(gpu_lib, n):
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    L = gpu_lib.potrf(S)
    assert relerr(L, scipy.linalg.cholesky(S, lower=True)) < 1e-12
    assert np.all(np.triu(L, 1) == 0.0)


def test_potrf_not_positive_definite(gpu_lib):
    A = np.eye(300)
    A[200, 200] = -2.0
    with pytest.raises(gsp.PosDefException) as ei:
        gpu_lib.potrf(A)
    assert ei.value.info == 201


# ------------------------------------------------------------------ LUSIM
def test_lusim_c1_config(gpu_lib):
    """BASELINE configs[0]: unconditional 50x50, SphericalCovariance(range=20), 100 realizations, seed 1."""
    st = iso(O.SPHERICAL, 1.0, 20.0, 2)
    coords = O.grid_centroids((50, 50), [0, 0], [1, 1])
    plan = gsp.LUPlan(gpu_lib, st, grid_dom((50, 50)), None, None, 0.0)
    pre = O.lusim_preprocess(ostructs(st), coords, np.zeros(0, dtype=np.int64), np.zeros(0), 0.0)
    W = np.random.default_rng(1).standard_normal((2500, 100))
    Z = plan.sample(100, W)
    Zo = O.lusim_sample(pre, W)
    assert max(relerr(Z[:, r], Zo[:, r]) for r in range(100)) < TOL
    d2, L22 = plan.get()
    assert relerr(L22, pre.L22) < 1e-11 and np.all(d2 == 0.0)
    plan.close()


@pytest.mark.parametrize("kind,nd,mu", [(O.EXPONENTIAL, 300, 0.0), (O.SPHERICAL, 17, 2.5), (O.CUBIC, 128, -1.0)])
def test_lusim_conditional(gpu_lib, kind, nd, mu):
    rng = np.random.default_rng(nd)
    dims = (64, 48)
    st = iso(kind, 1.4, 12.0, 2) + [(O.NUGGET, 0.05, np.eye(3))]
    coords = O.grid_centroids(dims, [0, 0], [1, 1])
    dinds = np.sort(rng.choice(coords.shape[0], nd, replace=False))
    z1 = rng.standard_normal(nd)
    plan = gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, z1, mu)
    pre = O.lusim_preprocess(ostructs(st), coords, dinds, z1, mu)
    R = 200
    W = rng.standard_normal((plan.Ns, R))
    Z = plan.sample(R, W)
    Zo = O.lusim_sample(pre, W)
    assert max(relerr(Z[:, r], Zo[:, r]) for r in range(R)) < TOL
    assert np.array_equal(Z[dinds], np.repeat(z1[:, None], R, 1))  # conditioning data honoured exactly
    d2, L22 = plan.get()
    assert relerr(L22, pre.L22) < 1e-10 and relerr(d2, pre.d2) < 1e-10
    plan.close()


def test_matern_lusim_and_fftsim(gpu_lib):
    """Matern structures (order 1.5 and 0.7, nested with a nugget) through the whole LUSIM and FFTSIM paths, injected noise."""
    rng = np.random.default_rng(41)
    dims = (48, 40)
    A = np.zeros((3, 3)); A[0, 0] = 1 / 14.0; A[1, 1] = 1 / 9.0
    st = [(O.MATERN, 0.9, A, 1.5), (O.MATERN, 0.4, A * 2.0, 0.7), (O.NUGGET, 0.05, np.eye(3))]
    coords = O.grid_centroids(dims, [0, 0], [1, 1])
    dinds = np.sort(rng.choice(coords.shape[0], 150, replace=False))
    z1 = rng.standard_normal(150)
    plan = gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, z1, 0.0)
    pre = O.lusim_preprocess(ostructs(st), coords, dinds, z1, 0.0)
    W = rng.standard_normal((plan.Ns, 50))
    Z = plan.sample(50, W)
    Zo = O.lusim_sample(pre, W)
    assert max(relerr(Z[:, r], Zo[:, r]) for r in range(50)) < TOL
    assert np.array_equal(Z[dinds], np.repeat(z1[:, None], 50, 1))
    plan.close()
    fdims = (64, 64, 32)
    A3 = np.diag([1 / 12.0, 1 / 8.0, 1 / 5.0])
    fst = [(O.MATERN, 1.0, A3, 2.5)]
    fplan = gsp.FFTPlan(gpu_lib, fst, fdims, [0.0] * 3, [1.0] * 3)
    Fo = O.fftsim_preprocess(ostructs(fst), fdims, [0.0] * 3, [1.0] * 3)
    w = rng.random((2, int(np.prod(fdims))))
    Zf = fplan.sample(2, w, sill=1.0, mu=0.0)
    for r in range(2):
        assert relerr(Zf[r], O.fftsim_sample(Fo, w[r], 1.0, 0.0)) < TOL
    fplan.close()


def test_lusim_bivariate(gpu_lib):
    """cosimulation (lusim.jl:112-126,164): rho-mixing of the noises, each variable with its own marginal plan."""
    rng = np.random.default_rng(5)
    dims = (40, 32)
    C = np.array([[1.0, 0.7], [0.7, 1.0]])
    A = np.eye(3) / 20.0
    mv = [(O.SPHERICAL, C, A)]
    coords = O.grid_centroids(dims, [0, 0], [1, 1])
    dinds = np.sort(rng.choice(1280, 50, replace=False))
    z = [rng.standard_normal(50), rng.standard_normal(50)]
    rho = O.rho_mv(mv)
    plans, pres = [], []
    for j in range(2):
        m = O.marginalize(mv, j)
        st = [(s.kind, s.sill, s.A) for s in m]
        plans.append(gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, z[j], 0.1 * (j + 1)))
        pres.append(O.lusim_preprocess(m, coords, dinds, z[j], 0.1 * (j + 1)))
    R = 64
    W1, W2 = rng.standard_normal((plans[0].Ns, R)), rng.standard_normal((plans[0].Ns, R))
    Z1 = plans[0].sample(R, W1)
    Z2 = plans[1].sample(R, W2, rho=rho, W1=W1)
    assert relerr(Z1, O.lusim_sample(pres[0], W1)) < TOL
    assert relerr(Z2, O.lusim_sample(pres[1], W2, rho, W1)) < TOL
    assert np.array_equal(Z2[dinds], np.repeat(z[1][:, None], R, 1))
    for p in plans:
        p.close()


def test_lusim_rejects_w1_without_w(gpu_lib):
    """both noises are injected or both come from the device RNG (a lone W1 used to be ignored silently)"""
    plan = gsp.LUPlan(gpu_lib, iso(O.SPHERICAL, 1.0, 5.0, 2), grid_dom((16, 12)), None, None, 0.0)
    W1 = np.random.default_rng(0).standard_normal((plan.Ns, 3))
    with pytest.raises(ValueError):
        plan.sample(3, None, rho=0.5, W1=W1)
    plan.close()


def test_lusim_shared_factor_gpu(gpu_lib):
    """gsp_lu_plan_create_like on the GPU: the second variable of a cosimulation shares the first one's factor; fields are bit-identical
    to those of a plan that factored on its own, and equal the oracle (lusim.jl:66-107,164)"""
    rng = np.random.default_rng(21)
    dims = (96, 80)                                  # 7,680 nodes: the panel algorithm (60 blocks)
    N, nd, R = 7680, 200, 16
    C = np.array([[1.0, 0.7], [0.7, 1.0]])
    mv = [(O.SPHERICAL, C, np.eye(3) / 15.0)]
    m0 = O.marginalize(mv, 0)
    st = [(s_.kind, s_.sill, s_.A) for s_ in m0]
    coords = O.grid_centroids(dims, [0, 0], [1, 1])
    dinds = np.sort(rng.choice(N, nd, replace=False))
    z = [rng.standard_normal(nd) * 0.4, rng.standard_normal(nd) * 0.4]
    base = gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, z[0], 0.0)
    shared = gsp.LUPlan(gpu_lib, None, None, dinds + 1, z[1], 0.0, like=base)
    own = gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, z[1], 0.0)
    W1, W2 = rng.standard_normal((base.Ns, R)), rng.standard_normal((base.Ns, R))
    Z2 = shared.sample(R, W2, rho=0.7, W1=W1)
    assert np.array_equal(Z2, own.sample(R, W2, rho=0.7, W1=W1))
    pre = O.lusim_preprocess(m0, coords, dinds, z[1], 0.0)
    assert relerr(Z2, O.lusim_sample(pre, W2, 0.7, W1)) < TOL
    base.close()
    assert np.array_equal(shared.sample(R, W2, rho=0.7, W1=W1), Z2)
    assert shared.times()[1] == 0.0 and shared.times()[2] > 0.0
    shared.close(), own.close()


def test_lusim_pointset_3d(gpu_lib):
    rng = np.random.default_rng(8)
    X = rng.uniform(0, 20, (700, 3))
    st = aniso3(O.EXPONENTIAL, 1.0, (8.0, 6.0, 3.0), 30.0)
    dinds = np.sort(rng.choice(700, 30, replace=False))
    z1 = rng.standard_normal(30)
    plan = gsp.LUPlan(gpu_lib, st, gsp._lib.make_point_domain(X), dinds + 1, z1, 0.0)
    pre = O.lusim_preprocess(ostructs(st), X, dinds, z1, 0.0)
    W = rng.standard_normal((670, 10))
    assert relerr(plan.sample(10, W), O.lusim_sample(pre, W)) < TOL
    plan.close()


def test_lusim_golden(gpu_lib):
    g = np.load(os.path.join(GOLD, "golden_small.npz"))
    st = iso(O.SPHERICAL, 1.0, 20.0, 2)
    plan = gsp.LUPlan(gpu_lib, st, grid_dom((12, 10)), g["lu_dinds"] + 1, g["lu_z1"], 0.0)
    assert relerr(plan.sample(4, g["lu_W"]), g["lu_Z"]) < TOL
    plan.close()


def test_lusim_c3_size_properties(gpu_lib):
    """BASELINE configs[2] at full size (16,384 nodes + 1,000 data, 1,000 realizations),
    1,000 device-RNG realizations: data honoured bit-exactly, d2 == K21 K11^-1 z1 on probes, ensemble moments
    (the factor itself is compared entry by entry in test_lusim_c3_full_size_oracle_parity)."""
    rng = np.random.default_rng(3)
    dims = (128, 128)
    st = iso(O.EXPONENTIAL, 1.0, 20.0, 2)
    coords = O.grid_centroids(dims, [0, 0], [1, 1])
    N, nd, R = 16384, 1000, 1000
    dinds = np.sort(rng.choice(N, nd, replace=False))
    pre0 = O.lusim_preprocess(ostructs(st), coords[dinds], np.zeros(0, dtype=np.int64), np.zeros(0), 0.0)
    z1 = O.lusim_sample(pre0, rng.standard_normal((nd, 1)))[:, 0]  # data consistent with the model
    plan = gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, z1, 0.0)
    Z = plan.sample(R, None, seed=33)
    assert np.array_equal(Z[dinds], np.repeat(z1[:, None], R, 1))
    sinds = np.setdiff1d(np.arange(N), dinds)
    # conditional mean: d2 = C21 C11^-1 z1
    C11 = O.pairwise(ostructs(st), coords[dinds])
    probe = rng.choice(len(sinds), 400, replace=False)
    C21 = O.pairwise(ostructs(st), coords[sinds[probe]], coords[dinds])
    d2_ref = C21 @ scipy.linalg.cho_solve(scipy.linalg.cho_factor(C11), z1)
    d2 = np.empty(plan.Ns)
    plan.lib.check(plan.lib.lib.gsp_lu_plan_get(plan.h, d2.ctypes.data, None))
    assert np.abs(d2[probe] - d2_ref).max() < 1e-9
    # ensemble mean -> d2 and conditional variance <= sill within sampling tolerance
    m = Z[sinds].mean(axis=1)
    assert np.abs(m - d2).max() < 6.0 / math.sqrt(R)
    v = Z[sinds].var(axis=1, ddof=1)
    assert v.max() < 1.35 and v.min() > 0.0
    plan.close()


def _c3_setup():
    rng = np.random.default_rng(3)
    dims = (128, 128)
    st = iso(O.EXPONENTIAL, 1.0, 20.0, 2)
    coords = O.grid_centroids(dims, [0, 0], [1, 1])
    N, nd = 16384, 1000
    dinds = np.sort(rng.choice(N, nd, replace=False))
    pre0 = O.lusim_preprocess(ostructs(st), coords[dinds], np.zeros(0, dtype=np.int64), np.zeros(0), 0.0)
    z1 = O.lusim_sample(pre0, rng.standard_normal((nd, 1)))[:, 0]
    return rng, dims, st, coords, N, nd, dinds, z1


def test_lusim_c3_full_size_oracle_parity(gpu_lib):
    """BASELINE configs[2] at FULL size against the full oracle (lusim.jl:95-103,160-169): the whole L22 (15,384^2, every
    recursion level, look-ahead stream and tile shape of the factorization), d2, and 8 injected-noise realizations at 1e-9."""
    rng, dims, st, coords, N, nd, dinds, z1 = _c3_setup()
    pre = O.lusim_preprocess(ostructs(st), coords, dinds, z1, 0.0)   # SciPy potrf/trsm at 16k: tens of seconds
    plan = gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, z1, 0.0)
    d2, L22 = plan.get()
    assert np.all(np.triu(L22, 1) == 0.0) and L22.diagonal().min() > 0.0
    assert relerr(d2, pre.d2) < 1e-10
    assert relerr(L22, pre.L22) < 1e-10
    # probe form of the same statement, independent of the oracle's own factorization:
    # L22 (L22' v) == (C22 - C21 C11^-1 C12) v for random v
    sinds = pre.sinds
    C11 = O.pairwise(ostructs(st), coords[dinds])
    C21 = O.pairwise(ostructs(st), coords[sinds], coords[dinds])
    C22 = O.pairwise(ostructs(st), coords[sinds])
    V = rng.standard_normal((len(sinds), 8))
    rhs = C22 @ V - C21 @ scipy.linalg.cho_solve(scipy.linalg.cho_factor(C11), C21.T @ V)
    lhs = L22 @ (L22.T @ V)
    assert np.abs(lhs - rhs).max() / np.abs(rhs).max() < 1e-10
    del C22, C21, lhs, rhs
    R = 8
    W = rng.standard_normal((plan.Ns, R))
    Z = plan.sample(R, W)
    Zo = O.lusim_sample(pre, W)
    assert max(relerr(Z[:, r], Zo[:, r]) for r in range(R)) < TOL
    assert np.array_equal(Z[dinds], np.repeat(z1[:, None], R, 1))
    plan.close()


def test_lusim_c5_full_size_oracle_parity(gpu_lib):
    """BASELINE configs[4] at FULL size (bivariate rho = 0.7, 32,768 nodes, 500 shared data nodes) against the full oracle:
    4 injected-noise realizations of both variables at 1e-9 (oracle.lusim_sample(pre, W2, rho, W1)), d2 and the whole L22.
    Both marginals are the same SphericalCovariance, so the oracle factorises once and only d2 differs per variable."""
    import dataclasses
    rng = np.random.default_rng(5)
    dims = (256, 128)
    N, nd, R = 32768, 500, 4
    C = np.array([[1.0, 0.7], [0.7, 1.0]])
    mv = [(O.SPHERICAL, C, np.eye(3) / 20.0)]
    coords = O.grid_centroids(dims, [0, 0], [1, 1])
    dinds = np.sort(rng.choice(N, nd, replace=False))
    z = [rng.standard_normal(nd) * 0.3, rng.standard_normal(nd) * 0.3]
    rho = O.rho_mv(mv)
    m0, m1 = O.marginalize(mv, 0), O.marginalize(mv, 1)
    assert [(a.kind, a.sill) for a in m0] == [(a.kind, a.sill) for a in m1] and all(np.array_equal(a.A, b.A) for a, b in zip(m0, m1))
    pre1 = O.lusim_preprocess(m0, coords, dinds, z[0], 0.0)
    C11 = O.pairwise(m0, coords[dinds])
    C21 = O.pairwise(m0, coords[pre1.sinds], coords[dinds])
    pre2 = dataclasses.replace(pre1, z1=z[1], d2=C21 @ scipy.linalg.cho_solve(scipy.linalg.cho_factor(C11), z[1]))
    del C21
    plans = [gsp.LUPlan(gpu_lib, [(s_.kind, s_.sill, s_.A) for s_ in m], grid_dom(dims), dinds + 1, z[j], 0.0) for j, m in enumerate((m0, m1))]
    W1, W2 = rng.standard_normal((plans[0].Ns, R)), rng.standard_normal((plans[0].Ns, R))
    Z1 = plans[0].sample(R, W1)
    Z2 = plans[1].sample(R, W2, rho=rho, W1=W1)
    Zo1 = O.lusim_sample(pre1, W1)
    Zo2 = O.lusim_sample(pre2, W2, rho, W1)
    assert max(relerr(Z1[:, r], Zo1[:, r]) for r in range(R)) < TOL
    assert max(relerr(Z2[:, r], Zo2[:, r]) for r in range(R)) < TOL
    assert np.array_equal(Z1[dinds], np.repeat(z[0][:, None], R, 1)) and np.array_equal(Z2[dinds], np.repeat(z[1][:, None], R, 1))
    d2a = np.empty(plans[1].Ns)
    plans[1].lib.check(plans[1].lib.lib.gsp_lu_plan_get(plans[1].h, d2a.ctypes.data, None))
    assert relerr(d2a, pre2.d2) < 1e-9
    plans[1].close()
    d2, L22 = plans[0].get()
    assert relerr(d2, pre1.d2) < 1e-9
    assert np.all(np.triu(L22, 1) == 0.0)
    err = 0.0
    for c0 in range(0, L22.shape[1], 2048):   # blockwise: no second 8 GB temporary
        err = max(err, float(np.abs(L22[:, c0:c0 + 2048] - pre1.L22[:, c0:c0 + 2048]).max()))
    assert err / np.abs(pre1.L22).max() < 1e-10
    plans[0].close()


def test_lusim_host_pipeline_matches_device_path(gpu_lib):
    """host-pointer sampling with several chunks per device (H2D / GEMM / D2H of consecutive chunks overlap on three streams,
    two buffer slots) returns exactly what the device-pointer path returns for the same injected noise, bivariate mixing included."""
    import torch
    rng = np.random.default_rng(9)
    dims = (48, 40)
    N, nd, R = 1920, 60, 1300          # 3 chunks of 512 columns
    st = iso(O.SPHERICAL, 1.0, 9.0, 2)
    dinds = np.sort(rng.choice(N, nd, replace=False))
    z1 = rng.standard_normal(nd)
    plan = gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, z1, 0.0)
    W = rng.standard_normal((plan.Ns, R))
    W1 = rng.standard_normal((plan.Ns, R))
    dev = torch.device("cuda:0")
    for rho, w1 in ((math.nan, None), (0.7, W1)):
        Zh = plan.sample(R, W, rho=rho, W1=w1)
        dW = torch.from_numpy(np.ascontiguousarray(W.T)).to(dev)
        dW1 = torch.from_numpy(np.ascontiguousarray(W1.T)).to(dev) if w1 is not None else None
        dZ = torch.empty((R, N), dtype=torch.float64, device=dev)
        plan.sample_dev(R, dW.data_ptr(), plan.Ns, 0, 0, 0, rho, dW1.data_ptr() if dW1 is not None else None, dZ.data_ptr(), N)
        torch.cuda.synchronize()
        assert np.array_equal(Zh, dZ.cpu().numpy().T)
        assert np.array_equal(Zh[dinds], np.repeat(z1[:, None], R, 1))
    plan.close()


def test_lusim_c5_size_properties(gpu_lib):
    """BASELINE configs[4] at full size: bivariate (rho = 0.7) LUSIM on 32,768 nodes with 500 shared data nodes.
    The oracle would need ~100 s per variable here, so check properties: data honoured bit-exactly for both variables,
    cross-correlation of the two simulated fields ~ rho, marginal variance <= sill, d2 against a direct solve."""
    rng = np.random.default_rng(5)
    dims = (256, 128)
    N, nd, R = 32768, 500, 192
    C = np.array([[1.0, 0.7], [0.7, 1.0]])
    A = np.eye(3) / 20.0
    mv = [(O.SPHERICAL, C, A)]
    coords = O.grid_centroids(dims, [0, 0], [1, 1])
    dinds = np.sort(rng.choice(N, nd, replace=False))
    z = [rng.standard_normal(nd) * 0.3, rng.standard_normal(nd) * 0.3]
    rho = O.rho_mv(mv)
    plans = []
    for j in range(2):
        m = O.marginalize(mv, j)
        plans.append(gsp.LUPlan(gpu_lib, [(s_.kind, s_.sill, s_.A) for s_ in m], grid_dom(dims), dinds + 1, z[j], 0.0))
    Z1 = plans[0].sample(R, None, seed=55, stream=0)
    Z2 = plans[1].sample(R, None, seed=55, stream=1, rho=rho)
    assert np.array_equal(Z1[dinds], np.repeat(z[0][:, None], R, 1))
    assert np.array_equal(Z2[dinds], np.repeat(z[1][:, None], R, 1))
    sinds = np.setdiff1d(np.arange(N), dinds)
    d2 = [np.empty(plans[j].Ns) for j in range(2)]
    for j in range(2):
        plans[j].lib.check(plans[j].lib.lib.gsp_lu_plan_get(plans[j].h, d2[j].ctypes.data, None))
    st1 = O.marginalize(mv, 0)
    C11 = O.pairwise(st1, coords[dinds])
    probe = rng.choice(len(sinds), 300, replace=False)
    C21 = O.pairwise(st1, coords[sinds[probe]], coords[dinds])
    assert np.abs(d2[0][probe] - C21 @ np.linalg.solve(C11, z[0])).max() < 1e-8
    r1 = Z1[sinds] - d2[0][:, None]
    r2 = Z2[sinds] - d2[1][:, None]
    far = np.ones(len(sinds), dtype=bool)  # nodes far from data: conditional residual correlation ~ rho
    corr = (r1[far] * r2[far]).mean() / np.sqrt((r1[far] ** 2).mean() * (r2[far] ** 2).mean())
    assert abs(corr - rho) < 0.02, corr
    assert r1.var(axis=1).max() < 1.6 and r1.var(axis=1).mean() < 1.0
    t = plans[0].times()
    assert t[1] > 0.0
    for p_ in plans:
        p_.close()


def test_lusim_statistics(gpu_lib):
    """ensemble covariance of unconditional LUSIM reproduces C within sampling tolerance (device RNG)."""
    st = iso(O.SPHERICAL, 2.0, 10.0, 2)
    dims = (24, 20)
    plan = gsp.LUPlan(gpu_lib, st, grid_dom(dims), None, None, 1.0)
    R = 20000
    Z = plan.sample(R, None, seed=2024)
    assert abs(Z.mean() - 1.0) < 0.05
    C = O.pairwise(ostructs(st), O.grid_centroids(dims, [0, 0], [1, 1]))
    Zc = Z - 1.0
    emp = (Zc[:60] @ Zc.T) / R
    assert np.abs(emp - C[:60]).max() < 0.12
    plan.close()


# ------------------------------------------------------------------ FFTSIM
@pytest.mark.parametrize("dims,kind,rang", [((100, 100), O.SPHERICAL, 10.0), ((256, 128), O.EXPONENTIAL, 25.0), ((64, 32, 16), O.SPHERICAL, 8.0),
                                            ((45, 30, 7), O.EXPONENTIAL, 5.0), ((1000,), O.SPHERICAL, 30.0), ((33, 27), O.CUBIC, 6.0)])
def test_fftsim_parity(gpu_lib, dims, kind, rang):
    rng = np.random.default_rng(sum(dims))
    nd = len(dims)
    st = iso(kind, 1.5, rang, nd)
    plan = gsp.FFTPlan(gpu_lib, st, dims, [0.0] * nd, [1.0] * nd)
    Fo = O.fftsim_preprocess(ostructs(st), dims, [0.0] * nd, [1.0] * nd)
    assert relerr(plan.spectrum(), Fo) < 1e-10
    N = int(np.prod(dims))
    w = rng.random((4, N))
    Z = plan.sample(4, w, sill=1.5, mu=0.2)
    for r in range(4):
        assert relerr(Z[r], O.fftsim_sample(Fo, w[r], 1.5, 0.2)) < TOL
    plan.close()


def test_fftsim_large_prime_extents(gpu_lib):
    """prime factors > 13 in the grid extents (101 x 101 as in a reference user's grid, 2 * 211, 3-D with 17 / 29 / 31)."""
    rng = np.random.default_rng(23)
    for dims in ((101, 101), (422, 37), (34, 29, 31), (1009,)):
        nd = len(dims)
        st = iso(O.EXPONENTIAL, 1.0, 7.0, nd)
        plan = gsp.FFTPlan(gpu_lib, st, dims, [0.0] * nd, [1.0] * nd)
        Fo = O.fftsim_preprocess(ostructs(st), dims, [0.0] * nd, [1.0] * nd)
        assert relerr(plan.spectrum(), Fo) < 1e-11, dims
        w = rng.random((3, int(np.prod(dims))))
        Z = plan.sample(3, w, sill=1.0, mu=0.1)
        Zo = np.stack([O.fftsim_sample(Fo, w[r], 1.0, 0.1) for r in range(3)])
        assert relerr(Z, Zo) < 1e-9, dims
        plan.close()


def test_fftsim_anisotropic_3d_and_view(gpu_lib):
    dims = (64, 64, 32)
    st = aniso3(O.SPHERICAL, 1.0, (20.0, 10.0, 5.0), 30.0)
    plan = gsp.FFTPlan(gpu_lib, st, dims, [0.0] * 3, [1.0] * 3)
    Fo = O.fftsim_preprocess(ostructs(st), dims, [0.0] * 3, [1.0] * 3)
    N = int(np.prod(dims))
    w = np.random.default_rng(4).random((1, N))
    inds = np.arange(0, N, 7)
    Zs = plan.sample(1, w, sill=1.0, mu=0.0, inds1=inds + 1)
    assert relerr(Zs[0], O.fftsim_sample(Fo, w[0], 1.0, 0.0, inds)) < TOL
    plan.close()


def test_fftsim_golden(gpu_lib):
    g = np.load(os.path.join(GOLD, "golden_small.npz"))
    plan = gsp.FFTPlan(gpu_lib, iso(O.EXPONENTIAL, 1.0, 5.0, 3), (8, 6, 4), [0.0] * 3, [1.0] * 3)
    assert relerr(plan.sample(1, g["fft_w"][None, :], sill=1.0, mu=0.5)[0], g["fft_Z"]) < TOL
    plan.close()


def test_fftsim_c2_config(gpu_lib):
    """BASELINE configs[1]: 1024x1024, GaussianCovariance(range=50), 64 realizations.  The Gaussian spectrum underflows to
    rounding noise (sqrt of ~1e-16 relative garbage), so F parity is asserted on F^2 = |fft(C)| and the fields
    are compared with a looser, stated bound; invariants are exact."""
    dims = (1024, 1024)
    st = iso(O.GAUSSIAN, 1.0, 50.0, 2)
    plan = gsp.FFTPlan(gpu_lib, st, dims, [0.0, 0.0], [1.0, 1.0])
    Fo = O.fftsim_preprocess(ostructs(st), dims, [0.0, 0.0], [1.0, 1.0])
    F = plan.spectrum()
    assert relerr(F ** 2, Fo ** 2) < 1e-12
    N = 1 << 20
    w = np.random.default_rng(2).random((64, N))
    Z = plan.sample(64, w, sill=1.0, mu=0.0)
    assert np.abs(Z.mean(axis=1)).max() < 1e-12
    assert np.abs((Z ** 2).sum(axis=1) / (N - 1) - 1.0).max() < 1e-12
    for r in (0, 63):
        assert relerr(Z[r], O.fftsim_sample(Fo, w[r], 1.0, 0.0)) < 1e-6  # conditioning-limited, see docstring
    plan.close()


def test_fftsim_c2_size_well_conditioned_parity(gpu_lib):
    """the C2 grid (1024 x 1024, 64 realizations) with a covariance whose spectrum does NOT underflow (Exponential, range 50): every
    realization at 1e-9 - an indexing error at this size cannot hide under the 1e-6 that the Gaussian model's conditioning forces above"""
    dims = (1024, 1024)
    st = iso(O.EXPONENTIAL, 1.0, 50.0, 2)
    plan = gsp.FFTPlan(gpu_lib, st, dims, [0.0, 0.0], [1.0, 1.0])
    Fo = O.fftsim_preprocess(ostructs(st), dims, [0.0, 0.0], [1.0, 1.0])
    assert relerr(plan.spectrum(), Fo) < 1e-10
    N = 1 << 20
    w = np.random.default_rng(22).random((64, N))
    Z = plan.sample(64, w, sill=1.0, mu=-0.5)
    for r in range(64):
        assert relerr(Z[r], O.fftsim_sample(Fo, w[r], 1.0, -0.5)) < TOL, r
    plan.close()


def test_fftsim_c4_size_properties(gpu_lib):
    """BASELINE configs[3] at full size (256^3, anisotropic spherical): one oracle realization for parity plus the
    size-independent invariants (exact mean, exact variance, determinism, shard invariance of the device RNG)."""
    dims = (256, 256, 256)
    st = aniso3(O.SPHERICAL, 1.0, (40.0, 20.0, 10.0), 30.0)
    plan = gsp.FFTPlan(gpu_lib, st, dims, [0.0] * 3, [1.0] * 3)
    N = 1 << 24
    w = np.random.default_rng(4).random((1, N))
    Z = plan.sample(1, w, sill=1.0, mu=0.25)
    assert abs(Z[0].mean() - 0.25) < 1e-12
    assert abs(((Z[0] - 0.25) ** 2).sum() / (N - 1) - 1.0) < 1e-11
    Fo = O.fftsim_preprocess(ostructs(st), dims, [0.0] * 3, [1.0] * 3)
    assert relerr(Z[0], O.fftsim_sample(Fo, w[0], 1.0, 0.25)) < TOL
    a = plan.sample(3, None, seed=77)
    b = np.concatenate([plan.sample(2, None, seed=77, first_real=0), plan.sample(1, None, seed=77, first_real=2)])
    assert np.array_equal(a, b)
    plan.close()


def test_fftsim_variogram_reproduction(gpu_lib):
    """empirical variogram along x of an ensemble vs the model gamma(h) = sill - C(h), h << grid."""
    dims = (256, 256)
    st = iso(O.SPHERICAL, 1.0, 20.0, 2)
    plan = gsp.FFTPlan(gpu_lib, st, dims, [0.0, 0.0], [1.0, 1.0])
    Z = plan.sample(32, None, seed=5).reshape(32, 256, 256)  # [r][y][x]
    for h in (1, 3, 6, 10):
        emp = 0.5 * np.mean((Z[:, :, h:] - Z[:, :, :-h]) ** 2)
        model = 1.0 - float(O.corr(O.SPHERICAL, np.array(h / 20.0)))
        assert abs(emp - model) < 0.08, (h, emp, model)
    plan.close()


def test_nearest_init_gpu(gpu_lib):
    """NearestInit on the device (nearest.jl:12-34) == the oracle, at the conditional-FFTSIM scale too
    (256^3 grid: the reference would build a KD-tree over 16.7 M centroids for these 5,000 data)"""
    rng = np.random.default_rng(11)
    for dims, nd in (((128, 128), 1000), ((256, 256, 256), 5000), ((50,), 300)):
        dim = len(dims)
        X = rng.uniform(-2.0, np.asarray(dims) + 2.0, (nd, dim))
        v = rng.standard_normal(nd)
        v[rng.choice(nd, nd // 9, replace=False)] = np.nan
        X[nd // 2:nd // 2 + nd // 4] = X[:nd // 4]
        dinds, z1 = gpu_lib.nearest_init(dims, [0.0] * dim, [1.0] * dim, X, v)
        do, zo = O.nearest_init(dims, [0.0] * dim, [1.0] * dim, X, v)
        assert np.array_equal(dinds, do) and np.array_equal(z1, zo)
    # through rand(): conditional LUSIM with the data table snapped on the device honours the data exactly
    proc = gsp.GaussianProcess(gsp.SphericalCovariance(range=10.0))
    grid = gsp.CartesianGrid(40, 30)
    pts = [(2.5, 2.5), (10.2, 7.9), (35.5, 12.5), (10.4, 7.6)]   # the last one falls on the node of the second: later wins
    data = gsp.georef({"Z": [0.3, -1.1, 0.8, 1.9]}, pts)
    real = gsp.rand(proc, grid, rng=np.random.default_rng(1), method=gsp.LUSIM(library=gpu_lib), data=data)
    assert real.Z[grid.nearest(np.array([10.4, 7.6]))] == 1.9 and real.Z[grid.nearest(np.array([2.5, 2.5]))] == 0.3


# ------------------------------------------------------------------ reference-facing API on the GPU
def test_rand_api_gpu(gpu_lib):
    rng = np.random.default_rng(123)
    proc = gsp.GaussianProcess(gsp.SphericalCovariance(range=10.0))
    grid = gsp.CartesianGrid(100, 100)
    ens = gsp.rand(proc, grid, 3, rng=rng, method=gsp.LUSIM(library=gpu_lib))
    assert len(ens) == 3 and ens[0].field.shape == (10000,)
    proc = gsp.GaussianProcess(gsp.GaussianVariogram(range=10.0))
    vgrid = grid.view(range(1, 5001))
    real = gsp.rand(proc, vgrid, rng=rng, method=gsp.FFTSIM(library=gpu_lib))
    assert real.domain == vgrid and real.nrow == 5000


def test_expectation_gpu(gpu_lib):
    """mean(process, domain; data) / expectedvalue (src/expectation/field.jl:17-44, field/gaussian.jl:21-25) - the reference's own
    "Expectation" test (test/field.jl:156-176) on the device Kriging, plus the oracle's Kriging on every element"""
    proc = gsp.GaussianProcess(gsp.SphericalVariogram(range=35.0))
    grid = gsp.CartesianGrid((0.5, 0.5), (100.5, 100.5), dims=(100, 100))
    assert np.all(gsp.mean(proc, grid).field == 0.0)
    pts = [(25.0, 25.0), (50.0, 75.0), (75.0, 50.0)]
    data = gsp.georef({"Z": [1.0, 0.0, 1.0]}, pts)
    mval = gsp.mean(proc, grid, data=data, library=gpu_lib)
    Z = mval.Z.reshape(100, 100)
    assert abs(Z[24, 24] - 1.0) < 1e-12 and abs(Z[74, 49]) < 1e-12 and abs(Z[49, 74] - 1.0) < 1e-12
    st = ostructs(iso(O.SPHERICAL, 1.0, 35.0, 2))
    cent = O.grid_centroids((100, 100), [0.5, 0.5], [1.0, 1.0])
    nbr, lam = O.krige_neighbors_weights(st, cent, np.array(pts), 10)
    assert np.abs(mval.Z - (lam * np.array([1.0, 0.0, 1.0])[nbr]).sum(axis=1)).max() < 1e-10
    assert np.array_equal(gsp.expectedvalue(proc, grid, data=data, library=gpu_lib).Z, mval.Z)


# ------------------------------------------------------------------ §8f rank 1: device-resident ensembles (src/ensembles.jl:42-52)
def test_ensemble_reference_pins_gpu(gpu_lib):
    """the reference's value-level ensemble test (test/ensembles.jl:24-59) on the CUDA kernels"""
    ens = gsp.DeviceEnsemble(gpu_lib, 9, 3)
    ens.put(np.stack([i * np.ones(9) for i in (1.0, 2.0, 3.0)]))
    ones = np.ones(9)
    assert np.array_equal(ens.mean(), 2.0 * ones) and np.array_equal(ens.var(), ones)
    for i in (1, 2, 3):
        assert np.array_equal(ens.cdf(i), i / 3 * ones)
        assert np.allclose(ens.ccdf(i), 1 - ens.cdf(i), rtol=1e-15, atol=1e-16)
    q = ens.quantile([0.0, 0.5, 1.0])
    assert np.array_equal(q, np.stack([ones, 2 * ones, 3 * ones]))
    ens.close()


@pytest.mark.parametrize("n,R", [(1000, 1), (4097, 2), (50_000, 257), (20_000, 1000), (300, 4096), (64, 10_000)])
def test_ensemble_statistics_gpu(gpu_lib, n, R):
    rng = np.random.default_rng(n + R)
    Z = rng.standard_normal((R, n)) * 2.5 - 17.0
    ens = gsp.DeviceEnsemble(gpu_lib, n, R)
    ens.put(Z)
    assert np.array_equal(ens.fetch(R - 1, 1)[0], Z[R - 1])
    assert relerr(ens.mean(), Z.mean(axis=0)) < 1e-13
    if R > 1:
        assert relerr(ens.var(), Z.var(axis=0, ddof=1)) < 1e-12
    x = -16.2
    assert np.array_equal(ens.cdf(x), (Z <= x).sum(axis=0) / R) and np.array_equal(ens.ccdf(x), (Z > x).sum(axis=0) / R)
    ps = [0.0, 0.05, 0.5, 0.777, 1.0]
    q = ens.quantile(ps)
    sub = slice(0, min(n, 64))  # scalar oracle (Julia's formula) on a few nodes, numpy on all of them
    qo = np.array([[O.julia_quantile(Z[:, i], p) for i in range(n)[sub]] for p in ps])
    assert relerr(q[:, sub], qo) < 1e-15
    assert relerr(q, np.quantile(Z, ps, axis=0)) < 1e-14
    ens.close()


def test_ensemble_resident_simulation_gpu(gpu_lib):
    """realizations simulated straight into a resident ensemble == the host-path fields; statistics of a conditional LUSIM
    ensemble: data nodes have mean z1 and variance 0, free nodes converge to the simple-kriging mean d2."""
    st = iso(O.SPHERICAL, 1.0, 8.0, 3)
    plan = gsp.FFTPlan(gpu_lib, st, (64, 64, 32), [0.0] * 3, [1.0] * 3)
    e = plan.sample_ensemble(9, None, seed=21, sill=1.0, mu=2.0)   # 9 realizations over 4 lanes: uneven last chunk
    assert np.array_equal(e.fetch(), plan.sample(9, None, seed=21, sill=1.0, mu=2.0))
    w = np.random.default_rng(2).random((3, 64 * 64 * 32))
    e2 = plan.sample_ensemble(3, w, sill=1.0, mu=2.0)
    assert np.array_equal(e2.fetch(), plan.sample(3, w, sill=1.0, mu=2.0))
    assert abs(e.mean().mean() - 2.0) < 1e-12  # every realization has exact spatial mean mu (fftsim.jl:91)
    e.close(), e2.close(), plan.close()
    dims = (40, 40)
    N = 1600
    rng = np.random.default_rng(8)
    dinds = np.sort(rng.choice(N, 60, replace=False))
    z1 = rng.standard_normal(60)
    st2 = iso(O.EXPONENTIAL, 1.0, 12.0, 2)
    lp = gsp.LUPlan(gpu_lib, st2, grid_dom(dims), dinds + 1, z1, 0.0)
    R = 4000
    ens = lp.sample_ensemble(R, None, seed=3)
    mean, var = ens.mean(), ens.var()
    assert np.array_equal(mean[dinds], z1) and np.all(var[dinds] == 0.0)
    d2, L22 = lp.get()
    sinds = np.setdiff1d(np.arange(N), dinds)
    cvar = (L22 ** 2).sum(axis=1)  # conditional variance = diag(L22 L22')
    assert np.abs(mean[sinds] - d2).max() < 5.0 * np.sqrt(cvar.max() / R)
    assert np.abs(var[sinds] / cvar - 1.0).max() < 0.15
    med = ens.quantile([0.5])[0]
    assert np.abs(med[sinds] - d2).max() < 6.0 * np.sqrt(cvar.max() / R)
    ens.close(), lp.close()


# ------------------------------------------------------------------ §8f rank 2: conditional FFTSIM (fftsim.jl:94-101,140-153)
@pytest.mark.parametrize("dims,nd,kind,maxn,view", [((64, 64), 100, O.EXPONENTIAL, 26, False), ((32, 32, 16), 80, O.SPHERICAL, 26, False),
                                                    ((48, 40), 60, O.CUBIC, 8, True), ((40, 40), 12, O.SPHERICAL, 26, False),
                                                    # more data than the warp-level pruning caches (1,056): full-scan path of the weights kernel
                                                    ((32, 32, 12), 1200, O.EXPONENTIAL, 26, False)])
def test_fftsim_conditional_parity(gpu_lib, dims, nd, kind, maxn, view):
    # (no GaussianCovariance here: its Kriging matrices are numerically singular, kappa ~ 1e12+, so the weights of two correct
    #  implementations agree to ~1e-6 only - the same caveat SURVEY §7 records for LUSIM + Gaussian)
    rng = np.random.default_rng(nd)
    nd_ = len(dims)
    st = iso(kind, 1.2, 6.0, nd_)
    N = int(np.prod(dims))
    mu = 0.4
    inds0 = np.sort(rng.choice(N, N // 2, replace=False)) if view else None
    cent = O.grid_centroids(dims, [0.0] * nd_, [1.0] * nd_)
    tg = cent if inds0 is None else cent[inds0]
    knodes0 = np.sort(rng.choice(tg.shape[0], nd, replace=False))
    dcoords = tg[knodes0] + rng.uniform(-0.4, 0.4, (nd, nd_))
    dvals = rng.standard_normal(nd) + mu
    plan = gsp.FFTPlan(gpu_lib, st, dims, [0.0] * nd_, [1.0] * nd_)
    inds1 = None if inds0 is None else inds0 + 1
    plan.condition(mu, dcoords, dvals, knodes0 + 1, inds1, maxneighbors=maxn)
    cond = O.fftsim_condition(ostructs(st), dims, [0.0] * nd_, [1.0] * nd_, dcoords, dvals, knodes0, mu, maxn, inds0)
    assert relerr(plan.condmean(), cond.zbar) < 1e-10
    Fo = O.fftsim_preprocess(ostructs(st), dims, [0.0] * nd_, [1.0] * nd_)
    w = rng.random((5, N))
    Z = plan.sample(5, w, sill=1.2, mu=mu, inds1=inds1)  # 5 realizations: one full chunk of 4 lanes + a remainder (3-D)
    for r in range(5):
        assert relerr(Z[r], O.fftsim_sample_conditional(Fo, w[r], 1.2, cond, inds0)) < TOL
    plan.close()


def test_fftsim_conditional_properties_large(gpu_lib):
    """128^3 grid, 500 data at node centroids, default 26 neighbours: data honoured, ensemble mean -> zbar, far field unconditional"""
    dims = (128, 128, 128)
    N = 128 ** 3
    rng = np.random.default_rng(17)
    st = aniso3(O.SPHERICAL, 1.0, (20.0, 10.0, 5.0), 30.0)
    knodes0 = np.sort(rng.choice(N, 500, replace=False))
    cent_k = np.stack([(knodes0 % 128) + 0.5, ((knodes0 // 128) % 128) + 0.5, (knodes0 // 16384) + 0.5], axis=1)
    dvals = rng.standard_normal(500)
    plan = gsp.FFTPlan(gpu_lib, st, dims, [0.0] * 3, [1.0] * 3)
    plan.condition(0.0, cent_k, dvals, knodes0 + 1)
    zbar = plan.condmean()
    assert np.abs(zbar[knodes0] - dvals).max() < 1e-10            # simple Kriging interpolates the data
    ens = plan.sample_ensemble(48, None, seed=9, sill=1.0, mu=0.0)
    Z3 = ens.fetch(0, 3)
    assert np.abs(Z3[:, knodes0] - dvals[None, :]).max() < 1e-10  # every realization honours the data
    mean, var = ens.mean(), ens.var()
    assert np.abs(mean - zbar).max() < 6.0 / np.sqrt(48) + 0.05     # E[z] = zbar
    assert var[knodes0].max() < 1e-18                             # no spread at the data
    assert 0.8 < var.mean() < 1.05                                # most nodes are farther than a range from any datum
    ens.close(), plan.close()
