"""Conditional FFTSIM (fftsim.jl:94-101,140-153): simple Kriging of residuals with k-nearest neighbourhoods - SURVEY §8f rank 2.
Emulated build here; the CUDA versions of the same checks are in test_gpu_parity.py."""
import numpy as np
import pytest
import scipy.linalg

import gsp_b200 as gsp
import gsp_oracle as O
from helpers import aniso3, iso, ostructs, relerr

TOL = 1e-9


def test_oracle_kriging_known_answers():
    """pins of the restated Kriging: exact interpolation at a sample, weights of an isolated pair, global solve when k = all"""
    st = ostructs(iso(O.EXPONENTIAL, 2.0, 5.0, 2))
    X = np.array([[1.0, 1.0], [4.0, 2.0], [2.5, 6.0], [7.0, 7.0]])
    nbr, lam = O.krige_neighbors_weights(st, X[[2]], X, 26)          # target = sample 2
    assert nbr[0, 0] == 2 and np.allclose(lam[0], [1.0, 0.0, 0.0, 0.0], atol=1e-14)
    t = np.array([[3.0, 3.0]])
    nbr, lam = O.krige_neighbors_weights(st, t, X, 4)
    C = O.pairwise(st, X)
    c0 = O.pairwise(st, X, t)[:, 0]
    full = np.linalg.solve(C, c0)
    assert np.allclose(lam[0], full[nbr[0]], rtol=1e-12)            # all samples: the global simple-Kriging system
    nbr2, lam2 = O.krige_neighbors_weights(st, t, X, 2)
    d = np.linalg.norm(X - t, axis=1)
    assert list(nbr2[0]) == list(np.argsort(d, kind="stable")[:2])
    # ties: four samples at the same distance, k = 2 -> the two with the lowest index
    Xs = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0]])
    nbr3, _ = O.krige_neighbors_weights(st, np.zeros((1, 2)), Xs, 2)
    assert list(nbr3[0]) == [0, 1]


def _case(lib, dims, nd, R, kind, rang, mu, maxn, seed, view=False, snapped=False):
    rng = np.random.default_rng(seed)
    nd_ = len(dims)
    st = iso(kind, 1.4, rang, nd_)
    N = int(np.prod(dims))
    inds0 = np.sort(rng.choice(N, N // 2, replace=False)) if view else None
    cent = O.grid_centroids(dims, [0.0] * nd_, [1.0] * nd_)
    tg = cent if inds0 is None else cent[inds0]
    n = tg.shape[0]
    knodes0 = np.sort(rng.choice(n, nd, replace=False))              # dinds = findall(mask), positions within sdom
    if snapped:
        dcoords = tg[knodes0].copy()                                 # data exactly at the centroids of their nodes
    else:
        dcoords = tg[knodes0] + rng.uniform(-0.45, 0.45, (nd, nd_))  # data somewhere inside their cells
    dvals = rng.standard_normal(nd) + mu
    plan = gsp.FFTPlan(lib, st, dims, [0.0] * nd_, [1.0] * nd_)
    inds1 = None if inds0 is None else inds0 + 1
    plan.condition(mu, dcoords, dvals, knodes0 + 1, inds1, maxneighbors=maxn)
    cond = O.fftsim_condition(ostructs(st), dims, [0.0] * nd_, [1.0] * nd_, dcoords, dvals, knodes0, mu, maxn, inds0)
    assert relerr(plan.condmean(), cond.zbar) < 1e-10
    Fo = O.fftsim_preprocess(ostructs(st), dims, [0.0] * nd_, [1.0] * nd_)
    w = rng.random((R, N))
    Z = plan.sample(R, w, sill=1.4, mu=mu, inds1=inds1)
    for r in range(R):
        assert relerr(Z[r], O.fftsim_sample_conditional(Fo, w[r], 1.4, cond, inds0)) < TOL
    if snapped:
        # data at the node centroids are honoured (Kriging is an exact interpolator; roundoff of a 26x26 solve)
        assert np.abs(Z[:, knodes0] - dvals[None, :]).max() < 1e-11
    # resident ensemble + device RNG go through the same conditioning
    e = plan.sample_ensemble(3, None, seed=5, sill=1.4, mu=mu, inds1=inds1)
    assert np.array_equal(e.fetch(), plan.sample(3, None, seed=5, sill=1.4, mu=mu, inds1=inds1))
    e.close()
    with pytest.raises(ValueError):
        plan.sample(1, w[:1], sill=1.4, mu=mu + 1.0, inds1=inds1)   # a different mean than the one conditioned on
    plan.close()


@pytest.mark.parametrize("dims,nd,kind,maxn,view,snapped", [
    ((16, 12), 9, O.SPHERICAL, 26, False, False),    # fewer data than maxneighbors: global neighbourhoods
    ((16, 12), 40, O.EXPONENTIAL, 26, False, True),  # the reference's default k = 26
    ((12, 10), 30, O.GAUSSIAN, 5, True, False),      # view of the grid, small neighbourhoods
    ((8, 6, 4), 35, O.SPHERICAL, 12, False, True),   # 3-D
    ((32,), 6, O.EXPONENTIAL, 3, False, False),      # 1-D
])
def test_conditional_fftsim_vs_oracle(emu_lib, dims, nd, kind, maxn, view, snapped):
    _case(emu_lib, dims, nd, 2, kind, 4.0, 0.7, maxn, seed=len(dims) * 100 + nd, view=view, snapped=snapped)


def test_conditional_without_covariance_table(emu_lib, monkeypatch):
    """GSP_KRIGE_TABLE=0: the weights kernel evaluates the sample-to-sample covariances itself (what it does beyond 4,096 samples)."""
    monkeypatch.setenv("GSP_KRIGE_TABLE", "0")
    _case(emu_lib, (16, 12), 40, 2, O.EXPONENTIAL, 5.0, 0.3, 26, 5)


def test_conditional_two_devices_and_errors(emu_lib):
    lib2 = gsp.Library(emu_lib.path, devices=[0, 0])
    _case(lib2, (12, 8), 20, 3, O.SPHERICAL, 3.0, -0.2, 8, seed=77)
    st = iso(O.SPHERICAL, 1.0, 3.0, 2)
    plan = gsp.FFTPlan(emu_lib, st, (8, 8), [0.0, 0.0], [1.0, 1.0])
    X = np.array([[1.5, 1.5], [1.5, 1.5]])  # coincident data: singular Kriging matrix -> the reference's cholesky throws
    with pytest.raises(gsp.GspError):
        plan.condition(0.0, X, np.array([0.1, 0.2]), np.array([10]))
    with pytest.raises(ValueError):
        plan.condition(0.0, X[:1], np.array([0.1]), np.array([10, 3]))  # knodes not ascending
    with pytest.raises(gsp.GspError):
        plan.condmean()
    plan.close()
    lib2.close()


def test_rand_fftsim_with_data(emu_lib):
    """rand(process, grid; data, method=FFTSIM()) - test/field.jl:134-154 restated (types/shapes) + the data are honoured"""
    grid = gsp.CartesianGrid(20, 20)
    proc = gsp.GaussianProcess(gsp.GaussianVariogram(range=5.0), 0.3)
    data = gsp.georef({"z": [1.0, 0.0, 1.0]}, [(5.5, 5.5), (10.5, 15.5), (15.5, 10.5)])  # at cell centroids
    ens = gsp.rand(proc, grid, 3, rng=np.random.default_rng(1), data=data, method=gsp.FFTSIM(library=emu_lib))
    assert len(ens) == 3 and ens.variables() == ("z",) and ens[0].z.shape == (400,)
    for pt, v in zip([(5.5, 5.5), (10.5, 15.5), (15.5, 10.5)], [1.0, 0.0, 1.0]):
        j = grid.nearest(np.array(pt))
        assert all(abs(r.z[j] - v) < 1e-10 for r in ens)
    single = gsp.rand(proc, grid, rng=np.random.default_rng(1), data=data, method=gsp.FFTSIM(library=emu_lib, maxneighbors=2))
    assert single.z.shape == (400,)
    # a view of the grid
    vgrid = grid.view(range(1, 201))
    real = gsp.rand(proc, vgrid, rng=np.random.default_rng(2), data=data, method=gsp.FFTSIM(library=emu_lib))
    assert real.domain == vgrid and real.nrow == 200


def test_conditional_plan_rejects_another_view(emu_lib):
    """a conditional plan carries zbar and weight tables of ONE simulation domain: sampling it with a different view of the same
    length (or a different mean) is an argument error, not silently wrong fields"""
    dims = (12, 10)
    st = iso(O.SPHERICAL, 1.0, 4.0, 2)
    rng = np.random.default_rng(0)
    inds0 = np.sort(rng.choice(120, 60, replace=False))
    other0 = np.sort(rng.choice(120, 60, replace=False))
    assert not np.array_equal(inds0, other0)
    cent = O.grid_centroids(dims, [0.0, 0.0], [1.0, 1.0])[inds0]
    knodes0 = np.sort(rng.choice(60, 8, replace=False))
    plan = gsp.FFTPlan(emu_lib, st, dims, [0.0, 0.0], [1.0, 1.0])
    plan.condition(0.3, cent[knodes0], rng.standard_normal(8), knodes0 + 1, inds0 + 1, maxneighbors=4)
    plan.sample(1, None, seed=1, sill=1.0, mu=0.3, inds1=inds0 + 1)
    with pytest.raises(ValueError):
        plan.sample(1, None, seed=1, sill=1.0, mu=0.3, inds1=other0 + 1)
    with pytest.raises(ValueError):
        plan.sample(1, None, seed=1, sill=1.0, mu=0.4, inds1=inds0 + 1)
    plan.close()
