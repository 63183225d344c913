"""Pins the oracle (CPU restatement): index conventions against the reference's own fixtures,
mathematical known answers, and the committed golden vectors (tests/golden)."""
import os

import numpy as np
import scipy.linalg

import gsp_oracle as O
from helpers import iso, ostructs, relerr

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_linear_index_convention_matches_reference_fixture():
    # test/initialization.jl:16-21: grid (-0.5,-0.5)-(99.5,99.5) dims (100,100); data2D.tsv rows
    # (25,25),(50,75),(75,50) land on 1-based elements 2526, 7551, 5076
    pts = np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]])
    dinds, z1 = O.nearest_init((100, 100), (-0.5, -0.5), (1.0, 1.0), pts, np.array([1.0, 2.0, 3.0]))
    assert sorted((dinds + 1).tolist()) == [2526, 5076, 7551]
    assert z1.tolist() == [1.0, 3.0, 2.0]  # ascending node order, not table order (lusim.jl:71)
    # test/initialization.jl:9-14 (1-D): x = 0,10,...,100 on (-0.5..99.5) -> elements 1,11,...,91,100
    x = np.arange(0.0, 101.0, 10.0)[:, None]
    d1, _ = O.nearest_init((100,), (-0.5,), (1.0,), x, np.arange(11.0))
    assert (d1 + 1).tolist() == [1, 11, 21, 31, 41, 51, 61, 71, 81, 91, 100]


def test_centroids_column_major():
    c = O.grid_centroids((3, 2), (0.0, 10.0), (1.0, 2.0))
    assert c.tolist() == [[0.5, 11.0], [1.5, 11.0], [2.5, 11.0], [0.5, 13.0], [1.5, 13.0], [2.5, 13.0]]


def test_model_known_values():
    u = np.array([0.0, 0.5, 1.0, 2.0])
    assert np.allclose(O.corr(O.SPHERICAL, u), [1.0, 1 - 0.75 + 0.0625, 0.0, 0.0])
    assert np.allclose(O.corr(O.EXPONENTIAL, u), np.exp(-3 * u))
    assert np.allclose(O.corr(O.GAUSSIAN, u), np.exp(-3 * u * u))
    assert O.corr(O.NUGGET, u).tolist() == [1.0, 0.0, 0.0, 0.0]
    for k in (O.CUBIC, O.PENTASPHERICAL, O.CIRCULAR):
        c = O.corr(k, u)
        assert c[0] == 1.0 and abs(c[2]) < 1e-15 and c[3] == 0.0
    # circular: area of the lens of two unit-diameter discs at distance u, / (pi/4); sine hole: first zero at u = 1
    assert abs(O.corr(O.CIRCULAR, np.array(0.5)) - (2 / np.pi) * (np.pi / 3 - 0.5 * np.sqrt(0.75))) < 1e-15
    assert np.allclose(O.corr(O.SINEHOLE, u), [1.0, 2 / np.pi, 0.0, 0.0], atol=1e-15)


def test_matern_known_values():
    """Matern closed forms at half-integer orders (delta = sqrt(2 nu) 3u): nu = 1/2 IS the exponential model of the same range."""
    u = np.array([0.0, 1e-9, 0.01, 0.2, 0.5, 1.0, 2.0, 10.0])
    assert np.allclose(O.corr(O.MATERN, u, 0.5), O.corr(O.EXPONENTIAL, u), rtol=1e-13, atol=1e-300)
    d = np.sqrt(3.0) * 3 * u
    assert np.allclose(O.corr(O.MATERN, u, 1.5), (1 + d) * np.exp(-d), rtol=1e-12, atol=1e-300)
    d = np.sqrt(5.0) * 3 * u
    assert np.allclose(O.corr(O.MATERN, u, 2.5), (1 + d + d * d / 3) * np.exp(-d), rtol=1e-12, atol=1e-300)
    c = O.corr(O.MATERN, u, 1.0)   # GeoStatsFunctions' default order
    assert c[0] == 1.0 and np.all(np.diff(c) < 0) and c[-1] < 1e-15
    assert O.corr(O.MATERN, np.array([300.0]), 0.7)[0] == 0.0   # K underflows: exactly zero, not NaN


def test_lusim_joint_cholesky_identity():
    """lusim.jl:95-103 equals the blocks of ONE Cholesky of the joint matrix [data; sim]."""
    rng = np.random.default_rng(0)
    st = ostructs(iso(O.EXPONENTIAL, 1.0, 7.0, 2))
    coords = O.grid_centroids((12, 9), (0, 0), (1, 1))
    dinds = np.sort(rng.choice(108, 11, replace=False))
    z1 = rng.standard_normal(11)
    pre = O.lusim_preprocess(st, coords, dinds, z1, 0.3)
    order = np.concatenate([dinds, pre.sinds])
    K = O.pairwise(st, coords[order])
    L = scipy.linalg.cholesky(K, lower=True)
    assert relerr(L[11:, 11:], pre.L22) < 1e-12
    y = scipy.linalg.solve_triangular(L[:11, :11], z1, lower=True)
    assert np.allclose(L[11:, :11] @ y, pre.d2, atol=1e-12)
    # d2 is the simple-kriging mean with ZERO mean (mu ignored when data exist, lusim.jl:102,172)
    C11 = O.pairwise(st, coords[dinds])
    C21 = O.pairwise(st, coords[pre.sinds], coords[dinds])
    assert np.allclose(pre.d2, C21 @ np.linalg.solve(C11, z1), atol=1e-11)
    Z = O.lusim_sample(pre, rng.standard_normal((len(pre.sinds), 3)))
    assert np.array_equal(Z[dinds], np.repeat(z1[:, None], 3, 1))


def test_lusim_unconditional_adds_mean_only_without_data():
    st = ostructs(iso(O.SPHERICAL, 2.0, 5.0, 1))
    coords = O.grid_centroids((30,), (0,), (1,))
    pre = O.lusim_preprocess(st, coords, np.zeros(0, dtype=np.int64), np.zeros(0), 1.5)
    Z = O.lusim_sample(pre, np.zeros((30, 1)))
    assert np.allclose(Z, 1.5)
    assert np.allclose(pre.L22 @ pre.L22.T, O.pairwise(st, coords), atol=1e-13)


def test_fftsim_invariants_and_parseval():
    """fftsim.jl:91,131-132: exact spatial mean mu, var(N-1, mean 0) == sill; sigma^2 is noise-independent."""
    rng = np.random.default_rng(1)
    for dims in ((32, 20), (8, 6, 10), (15,)):
        nd = len(dims)
        st = ostructs(iso(O.SPHERICAL, 1.7, 4.0, nd))
        F = O.fftsim_preprocess(st, dims, [0.0] * nd, [1.0] * nd)
        n = int(np.prod(dims))
        assert F.reshape(-1)[0] == 0.0
        w = rng.random(n)
        z = O.fftsim_sample(F, w, 1.7, 0.25)
        assert abs(z.mean() - 0.25) < 1e-12
        assert abs(((z - 0.25) ** 2).sum() / (n - 1) - 1.7) < 1e-12
        # Parseval: var before scaling == sum(F^2) / (N (N-1))
        Wh = np.fft.fftn(w.reshape(F.shape))
        Zr = np.real(np.fft.ifftn(F * np.exp(1j * np.angle(Wh))))
        assert abs((Zr ** 2).sum() / (n - 1) - (F ** 2).sum() / (n * (n - 1.0))) < 1e-15 * max(1.0, (F ** 2).sum())


def test_marginalize_and_rho():
    A = np.eye(3) / 10.0
    mv = [(O.GAUSSIAN, np.array([[1.0, 0.95], [0.95, 1.0]]), A)]
    assert abs(O.rho_mv(mv) - 0.95) < 1e-15
    m1 = O.marginalize(mv, 1)
    assert len(m1) == 1 and m1[0].kind == O.GAUSSIAN and m1[0].sill == 1.0


def test_golden_vectors():
    """Regression pin: fixtures written by tests/golden/make_golden.py (oracle outputs on seeded inputs)."""
    g = np.load(os.path.join(GOLD, "golden_small.npz"))
    st = ostructs(iso(O.SPHERICAL, 1.0, 20.0, 2))
    coords = O.grid_centroids((12, 10), (0, 0), (1, 1))
    pre = O.lusim_preprocess(st, coords, g["lu_dinds"], g["lu_z1"], 0.0)
    assert relerr(O.lusim_sample(pre, g["lu_W"]), g["lu_Z"]) < 1e-12
    st = ostructs(iso(O.EXPONENTIAL, 1.0, 5.0, 3))
    F = O.fftsim_preprocess(st, (8, 6, 4), [0, 0, 0], [1, 1, 1])
    assert relerr(O.fftsim_sample(F, g["fft_w"], 1.0, 0.5), g["fft_Z"]) < 1e-12
