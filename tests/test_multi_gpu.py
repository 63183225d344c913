"""Real multi-GPU tests (-m gpu; skipped on a 1-GPU box): one context over G devices against the 1-device context.

LUSIM: the distributed Cholesky (row panels owned by different devices, panels multicast over NVLink peer stores) and the
realization-sharded sampling must reproduce the single-device fields (<= 1e-12 relative: the summation order of the trailing
updates differs) and the oracle (1e-9, lusim.jl:95-103,160-169).  FFTSIM / ensembles: realizations are independent and the
device RNG is counter-based, so the multi-device fields are bit-identical (fftsim.jl:124-135)."""
import math

import numpy as np
import pytest

import gsp_b200 as gsp
import gsp_oracle as O
from helpers import aniso3, iso, ostructs, relerr

pytestmark = pytest.mark.gpu


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def need(g):
    return pytest.mark.skipif(ngpus() < g, reason=f"needs {g} GPUs")


GS = [pytest.param(g, marks=need(g)) for g in (2, 4, 8)]


def grid_dom(dims):
    return (gsp._lib.make_grid_domain(dims, [0.0] * len(dims), [1.0] * len(dims)), None)


@pytest.mark.parametrize("G", GS)
def test_lusim_distributed_factorization_matches_single_device(gpu_lib, G):
    """96 x 96 grid + 300 data (Np = 9,472 -> 74 blocks, above the distribution threshold): factor over G devices"""
    rng = np.random.default_rng(40 + G)
    dims = (96, 96)
    N, nd, R = 9216, 300, 37                      # R not a multiple of G: uneven realization shards
    st = iso(O.EXPONENTIAL, 1.2, 15.0, 2) + [(O.NUGGET, 0.02, np.eye(3))]
    coords = O.grid_centroids(dims, [0, 0], [1, 1])
    dinds = np.sort(rng.choice(N, nd, replace=False))
    z1 = rng.standard_normal(nd)
    libG = gsp.Library(devices=list(range(G)))
    p1 = gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, z1, 0.0)
    pG = gsp.LUPlan(libG, st, grid_dom(dims), dinds + 1, z1, 0.0)
    W = rng.standard_normal((p1.Ns, R))
    Z1, ZG = p1.sample(R, W), pG.sample(R, W)
    assert relerr(ZG, Z1) < 1e-12
    assert np.array_equal(ZG[dinds], np.repeat(z1[:, None], R, 1))
    pre = O.lusim_preprocess(ostructs(st), coords, dinds, z1, 0.0)
    Zo = O.lusim_sample(pre, W)
    assert max(relerr(ZG[:, r], Zo[:, r]) for r in range(R)) < 1e-9
    d2, L22 = pG.get()
    assert relerr(L22, pre.L22) < 1e-10 and relerr(d2, pre.d2) < 1e-10
    # device RNG: sharding over G devices does not change a realization
    assert relerr(pG.sample(R, None, seed=5), p1.sample(R, None, seed=5)) < 1e-12
    # resident ensemble sharded over the devices: statistics merge across devices
    e1, eG = p1.sample_ensemble(R, W), pG.sample_ensemble(R, W)
    assert relerr(eG.mean(), e1.mean()) < 1e-12 and relerr(eG.var(), e1.var()) < 1e-11
    assert relerr(eG.quantile([0.25, 0.5])[1], e1.quantile([0.25, 0.5])[1]) < 1e-12
    e1.close(), eG.close(), p1.close(), pG.close(), libG.close()


@pytest.mark.parametrize("G", GS)
def test_lusim_distributed_not_positive_definite(gpu_lib, G):
    """a non-positive pivot met by ANY panel owner surfaces as PosDefException with the reference's index"""
    libG = gsp.Library(devices=list(range(G)))
    # 7,000 mutually uncorrelated points (identity block), then 2,000 points 0.01 apart under a Gaussian model of range 30:
    # numerically singular from the cluster's 3rd-4th point on, i.e. on a panel that device 0 does not own
    i = np.arange(7000)
    X = np.concatenate([np.stack([(i % 100) * 200.0, (i // 100) * 200.0], 1),
                        np.stack([1e5 + 0.01 * np.arange(2000), np.full(2000, 1e5)], 1)])
    st = iso(O.GAUSSIAN, 1.0, 30.0, 2)
    with pytest.raises(gsp.PosDefException) as ei:
        gsp.LUPlan(libG, st, gsp._lib.make_point_domain(X), None, None, 0.0)
    assert 7002 <= ei.value.info <= 7064, ei.value.info
    with pytest.raises(gsp.PosDefException) as e1:
        gsp.LUPlan(gpu_lib, st, gsp._lib.make_point_domain(X), None, None, 0.0)
    assert abs(e1.value.info - ei.value.info) <= 8   # the first rounding-noise pivot may differ by a few rows between summation orders
    libG.close()


@pytest.mark.parametrize("G", GS)
def test_fftsim_and_ensemble_over_devices_bit_identical(gpu_lib, G):
    dims = (64, 64, 32)
    st = aniso3(O.SPHERICAL, 1.0, (20.0, 10.0, 5.0), 30.0)
    libG = gsp.Library(devices=list(range(G)))
    p1 = gsp.FFTPlan(gpu_lib, st, dims, [0.0] * 3, [1.0] * 3)
    pG = gsp.FFTPlan(libG, st, dims, [0.0] * 3, [1.0] * 3)
    R = 4 * G + 3
    w = np.random.default_rng(G).random((R, int(np.prod(dims))))
    assert np.array_equal(pG.sample(R, w, sill=1.0, mu=0.3), p1.sample(R, w, sill=1.0, mu=0.3))
    assert np.array_equal(pG.sample(R, None, seed=9), p1.sample(R, None, seed=9))
    e1, eG = p1.sample_ensemble(R, None, seed=9), pG.sample_ensemble(R, None, seed=9)
    assert np.array_equal(eG.fetch(), e1.fetch())
    assert relerr(eG.mean(), e1.mean()) < 1e-13 and relerr(eG.var(), e1.var()) < 1e-12
    assert np.array_equal(eG.cdf(0.1), e1.cdf(0.1))
    assert np.array_equal(eG.quantile([0.1, 0.9]), e1.quantile([0.1, 0.9]))
    e1.close(), eG.close(), p1.close(), pG.close(), libG.close()


@pytest.mark.parametrize("G", GS)
def test_shared_factor_plan_over_devices(gpu_lib, G):
    """gsp_lu_plan_create_like on a multi-device context: the second variable shares the distributed factor on every device"""
    rng = np.random.default_rng(60 + G)
    dims = (96, 96)
    N, nd, R = 9216, 150, 4 * G + 1
    st = iso(O.SPHERICAL, 1.0, 12.0, 2)
    dinds = np.sort(rng.choice(N, nd, replace=False))
    za, zb = rng.standard_normal(nd), rng.standard_normal(nd)
    libG = gsp.Library(devices=list(range(G)))
    base = gsp.LUPlan(libG, st, grid_dom(dims), dinds + 1, za, 0.0)
    shared = gsp.LUPlan(libG, None, None, dinds + 1, zb, 0.0, like=base)
    own = gsp.LUPlan(libG, st, grid_dom(dims), dinds + 1, zb, 0.0)
    W1, W2 = rng.standard_normal((base.Ns, R)), rng.standard_normal((base.Ns, R))
    Zs = shared.sample(R, W2, rho=0.7, W1=W1)
    assert relerr(Zs, own.sample(R, W2, rho=0.7, W1=W1)) < 1e-12   # two factorizations over G devices: not bit-identical to each other
    assert np.array_equal(Zs[dinds], np.repeat(zb[:, None], R, 1))
    one = gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, zb, 0.0)
    assert relerr(Zs, one.sample(R, W2, rho=0.7, W1=W1)) < 1e-12
    base.close()
    assert np.array_equal(shared.sample(R, W2, rho=0.7, W1=W1), Zs)
    shared.close(), own.close(), one.close(), libG.close()


@need(2)
def test_c3_over_all_devices(gpu_lib):
    """BASELINE configs[2] (16,384 nodes + 1,000 data) factored over every visible GPU == the single-device plan"""
    G = ngpus()
    rng = np.random.default_rng(3)
    dims = (128, 128)
    N, nd, R = 16384, 1000, 16
    st = iso(O.EXPONENTIAL, 1.0, 20.0, 2)
    dinds = np.sort(rng.choice(N, nd, replace=False))
    z1 = rng.standard_normal(nd) * 0.5
    libG = gsp.Library(devices=list(range(G)))
    p1 = gsp.LUPlan(gpu_lib, st, grid_dom(dims), dinds + 1, z1, 0.0)
    pG = gsp.LUPlan(libG, st, grid_dom(dims), dinds + 1, z1, 0.0)
    W = rng.standard_normal((p1.Ns, R))
    Z1, ZG = p1.sample(R, W), pG.sample(R, W)
    assert relerr(ZG, Z1) < 1e-12
    assert np.array_equal(ZG[dinds], np.repeat(z1[:, None], R, 1))
    assert math.isfinite(pG.times()[1]) and pG.times()[1] > 0
    p1.close(), pG.close(), libG.close()
