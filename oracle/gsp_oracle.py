"""CPU oracle for the LUSIM / FFTSIM hot path of GeoStatsProcesses.jl v0.13.0.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
`geostatsprocesses.jl_b200/`, the C-ABI library) may import or call this file.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs use it,
and there only as the checker / the timed CPU baseline.

PARITY UNPINNED.  The reference cannot run here (no Julia; its arithmetic lives in
un-vendored packages GeoStatsFunctions 0.15, Meshes 0.57, FFTW 1.7, LinearAlgebra/
OpenBLAS) and the reference's own tests for this path (test/field.jl:17-71,115-154)
hold no golden vectors - only eltype/unit/length assertions.  This file is therefore
a line-by-line *restatement* in NumPy/SciPy float64 of

  src/simulation/field/lusim.jl:38-175      (preprocess, randsingle, _marginalize, _rho, _lusim)
  src/simulation/field/fftsim.jl:54-92,109-139 (unconditional preprocess + randsingle)
  src/simulation/field/fftsim.jl:94-101,140-153 (conditioning: simple Kriging of residuals; GeoStatsModels' k-nearest
                                            neighbourhood search restated, ties -> lower sample index)
  src/ensembles.jl:42-52                    (ensemble statistics; pinned by the reference's test/ensembles.jl:24-59)
  src/utils.jl:50-62                        (_pairwise: sill - gamma for variograms)
  src/processes/field.jl:43-58, src/initialization/nearest.jl:12-34 (dinds ordering)

plus the published model formulas of GeoStatsFunctions (practical-range convention)
and the Meshes CartesianGrid centroid rule origin + (ijk - 1/2)*spacing with
column-major linear indices (pinned by the reference's test/initialization.jl:16-21,
which tests/test_oracle.py re-checks).  The oracle is additionally pinned to
mathematics by known-answer tests in tests/test_oracle.py (L L' = C, conditional
mean/covariance identities, FFT-MA invariants).

What WOULD pin it: tests/golden/make_golden.jl runs the real GeoStatsProcesses v0.13.0 (Julia needed, absent here),
captures the noise it drew and writes reference vectors that tests/test_reference_fixtures.py compares this file
with at 1e-12 (strict-xfail until the vectors are committed).

All noise is INJECTED (normals for LUSIM, uniforms for FFTSIM) so that the GPU path
and the oracle consume identical arrays.
"""
from __future__ import annotations

import math

import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import scipy.fft
import scipy.linalg

# ----------------------------------------------------------------------------
# covariance models (GeoStatsFunctions formulas, restated from the published
# definitions; the dependency is not vendored in /root/reference)
# ----------------------------------------------------------------------------
NUGGET, SPHERICAL, EXPONENTIAL, GAUSSIAN, CUBIC, PENTASPHERICAL, SINEHOLE, CIRCULAR, MATERN = 0, 1, 2, 3, 4, 5, 6, 7, 8


@dataclass
class Structure:
    """One nested structure: contribution `sill` * rho(|A @ delta|).

    `A` is the dim x dim matrix mapping a coordinate difference to the unit-range
    Mahalanobis frame (isotropic range r: A = I / r; MetricBall(radii, R):
    A = diag(1/radii) @ R')."""
    kind: int
    sill: float
    A: np.ndarray = field(default_factory=lambda: np.eye(3))
    param: float = 0.0   # Matern: order nu


def corr(kind: int, u: np.ndarray, param: float = 0.0) -> np.ndarray:
    """Correlation rho(u) of the basic models at normalised lag u = h / range."""
    u = np.asarray(u, dtype=np.float64)
    if kind == MATERN:
        # GeoStatsFunctions MaternVariogram: delta = sqrt(2 nu) 3 h/r, Omega = 2^(1-nu)/Gamma(nu) delta^nu, gamma = s (1 - Omega K_nu(delta)).
        # The reference evaluates at h + eps() to avoid 0 * Inf at the origin; rho(0) = 1 here (difference <= 1e-15).
        import scipy.special
        nu = float(param)
        d = np.sqrt(2.0 * nu) * 3.0 * u
        with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
            r = (2.0 ** (1.0 - nu) / scipy.special.gamma(nu)) * d**nu * scipy.special.kv(nu, d)
        r = np.where(np.isfinite(r), r, 1.0)
        return np.where(u == 0, 1.0, np.minimum(r, 1.0))
    if kind == NUGGET:
        return (u == 0).astype(np.float64)
    if kind == SPHERICAL:
        return np.where(u < 1, 1.0 - 1.5 * u + 0.5 * u**3, 0.0)
    if kind == EXPONENTIAL:
        return np.exp(-3.0 * u)
    if kind == GAUSSIAN:
        return np.exp(-3.0 * u * u)
    if kind == CUBIC:
        g = 7 * u**2 - 8.75 * u**3 + 3.5 * u**5 - 0.75 * u**7
        return np.where(u < 1, 1.0 - g, 0.0)
    if kind == PENTASPHERICAL:
        g = 1.875 * u - 1.25 * u**3 + 0.375 * u**5
        return np.where(u < 1, 1.0 - g, 0.0)
    if kind == SINEHOLE:  # gamma = s (1 - sin(pi u) / (pi u))
        return np.sinc(u)   # numpy's sinc is sin(pi x) / (pi x), 1 at 0
    if kind == CIRCULAR:  # gamma = s (1 - (2/pi) acos(u) + (2u/pi) sqrt(1 - u^2)) for u < 1
        uc = np.minimum(u, 1.0)
        return np.where(u < 1, (2.0 / np.pi) * (np.arccos(uc) - uc * np.sqrt(1.0 - uc * uc)), 0.0)
    raise ValueError(f"unknown model kind {kind}")


def model_sill(structs: Sequence[Structure]) -> float:
    return float(sum(s.sill for s in structs))


def cov_eval(structs: Sequence[Structure], delta: np.ndarray) -> np.ndarray:
    """C(delta) for coordinate differences delta[..., dim].

    Variograms are handled by the caller's convention: utils.jl:50-62 turns a
    variogram gamma into sill - gamma, which equals this covariance form."""
    delta = np.asarray(delta, dtype=np.float64)
    dim = delta.shape[-1]
    out = np.zeros(delta.shape[:-1])
    for s in structs:
        A = np.asarray(s.A, dtype=np.float64)[:dim, :dim]
        t = delta @ A.T
        u = np.sqrt(np.sum(t * t, axis=-1))
        out = out + s.sill * corr(s.kind, u, s.param)
    return out


def pairwise(structs: Sequence[Structure], X1: np.ndarray, X2: Optional[np.ndarray] = None,
             block: int = 1024) -> np.ndarray:
    """_pairwise(fun, dom1[, dom2]) - src/utils.jl:50-62.  X: (n, dim)."""
    X1 = np.asarray(X1, dtype=np.float64)
    X2 = X1 if X2 is None else np.asarray(X2, dtype=np.float64)
    out = np.empty((X1.shape[0], X2.shape[0]))
    for i0 in range(0, X1.shape[0], block):
        d = X1[i0:i0 + block, None, :] - X2[None, :, :]
        out[i0:i0 + block] = cov_eval(structs, d)
    return out


# ----------------------------------------------------------------------------
# CartesianGrid conventions (Meshes): centroid = origin + (ijk - 1/2) * spacing,
# linear index column-major (x fastest) - test/initialization.jl:16-21
# ----------------------------------------------------------------------------
def grid_centroids(dims: Sequence[int], origin: Sequence[float], spacing: Sequence[float]) -> np.ndarray:
    dims = tuple(int(d) for d in dims)
    axes = [origin[a] + (np.arange(dims[a]) + 0.5) * spacing[a] for a in range(len(dims))]
    mesh = np.meshgrid(*axes, indexing="ij")
    # column-major flattening: first axis fastest
    return np.stack([m.reshape(-1, order="F") for m in mesh], axis=1)


def nearest_init(dims, origin, spacing, data_coords: np.ndarray, data_vals: np.ndarray):
    """initialize + NearestInit (src/processes/field.jl:43-58, src/initialization/nearest.jl:12-34)
    for a CartesianGrid: each datum snaps to the nearest element centroid, later data
    overwrite earlier ones, NaN (= missing) is skipped.  Returns (dinds0 ascending 0-based, z1)."""
    dims = tuple(int(d) for d in dims)
    n = int(np.prod(dims))
    vals = np.zeros(n)
    mask = np.zeros(n, dtype=bool)
    data_coords = np.atleast_2d(np.asarray(data_coords, dtype=np.float64))
    for p, v in zip(data_coords, data_vals):
        if np.isnan(v):
            continue
        lin, stride = 0, 1
        for a in range(len(dims)):
            i = int(np.floor((p[a] - origin[a]) / spacing[a]))
            i = min(max(i, 0), dims[a] - 1)
            lin += i * stride
            stride *= dims[a]
        vals[lin] = v
        mask[lin] = True
    dinds = np.flatnonzero(mask)  # findall(mask): ascending node order (lusim.jl:71)
    return dinds, vals[dinds]


# ----------------------------------------------------------------------------
# LUSIM - src/simulation/field/lusim.jl
# ----------------------------------------------------------------------------
@dataclass
class LUPre:
    z1: np.ndarray
    mu1: float
    d2: np.ndarray
    L22: np.ndarray
    dinds: np.ndarray
    sinds: np.ndarray


def marginalize(structs_mv, j: int):
    """_marginalize (lusim.jl:132-137): NuggetEffect(c0[j,j]) + sum_k c_k[j,j] * cov_k.
    `structs_mv` is a list of (kind, Cmat (nv x nv), A)."""
    out = []
    for kind, C, A in structs_mv:
        c = float(np.asarray(C)[j, j])
        if kind == NUGGET and c == 0.0:
            continue
        out.append(Structure(kind, c, A))
    return out


def rho_mv(structs_mv) -> float:
    """_rho (lusim.jl:139-143): C12(0) / sqrt(s1 * s2)."""
    c12 = sum(float(np.asarray(C)[0, 1]) for _, C, _ in structs_mv)
    s1 = sum(float(np.asarray(C)[0, 0]) for _, C, _ in structs_mv)
    s2 = sum(float(np.asarray(C)[1, 1]) for _, C, _ in structs_mv)
    return c12 / np.sqrt(s1 * s2)


def lusim_preprocess(structs, coords: np.ndarray, dinds: np.ndarray, z1: np.ndarray, mu1: float) -> LUPre:
    """preprocess for ONE variable (the body of the `map` at lusim.jl:66-107).
    coords (N, dim); dinds 0-based ascending."""
    n = coords.shape[0]
    dinds = np.asarray(dinds, dtype=np.int64)
    sinds = np.setdiff1d(np.arange(n), dinds)                   # lusim.jl:72
    ddom, sdom = coords[dinds], coords[sinds]                   # lusim.jl:81-82
    C22 = pairwise(structs, sdom)                               # lusim.jl:88
    if len(dinds) == 0:                                         # lusim.jl:90-92
        d2 = np.zeros(len(sinds))
        L22 = scipy.linalg.cholesky(C22, lower=True)
    else:
        C11 = pairwise(structs, ddom)                           # lusim.jl:95
        C12 = pairwise(structs, ddom, sdom)                     # lusim.jl:96
        L11 = scipy.linalg.cholesky(C11, lower=True)            # lusim.jl:98
        B12 = scipy.linalg.solve_triangular(L11, C12, lower=True)   # lusim.jl:99
        A21 = B12.T                                             # lusim.jl:100
        d2 = A21 @ scipy.linalg.solve_triangular(L11, np.asarray(z1, float), lower=True)  # lusim.jl:102
        L22 = scipy.linalg.cholesky(C22 - A21 @ B12, lower=True)    # lusim.jl:103
    return LUPre(np.asarray(z1, float), float(mu1), d2, L22, dinds, sinds)


def lusim_sample(pre: LUPre, W: np.ndarray, rho: Optional[float] = None, W1: Optional[np.ndarray] = None) -> np.ndarray:
    """_lusim (lusim.jl:145-175) batched over the columns of W (Ns x R).  Returns Z (N x R)."""
    W = np.asarray(W, dtype=np.float64)
    if W.ndim == 1:
        W = W[:, None]
    n = len(pre.dinds) + len(pre.sinds)
    if rho is None:
        z2 = pre.d2[:, None] + pre.L22 @ W                                            # lusim.jl:162
    else:
        W1 = np.asarray(W1, dtype=np.float64).reshape(W.shape)
        z2 = pre.d2[:, None] + pre.L22 @ (rho * W1 + np.sqrt(1 - rho**2) * W)         # lusim.jl:164
    Z = np.empty((n, W.shape[1]))
    Z[pre.dinds, :] = pre.z1[:, None]                                                 # lusim.jl:168
    Z[pre.sinds, :] = z2                                                              # lusim.jl:169
    if len(pre.dinds) == 0:                                                           # lusim.jl:172
        Z += pre.mu1
    return Z


# ----------------------------------------------------------------------------
# FFTSIM - src/simulation/field/fftsim.jl (unconditional path)
# ----------------------------------------------------------------------------
def _workers():
    return os.cpu_count() or 1


def fftsim_preprocess(structs, dims, origin, spacing) -> np.ndarray:
    """fftsim.jl:77-91.  Returns F with shape dims[::-1] in C order == column-major dims."""
    dims = tuple(int(d) for d in dims)
    cent = grid_centroids(dims, origin, spacing)
    cind = [d // 2 for d in dims]                        # CartesianIndex(dims .÷ 2), 1-based (fftsim.jl:80)
    # 1-based index (c1,c2,..) -> 0-based (c1-1, ...); dims .÷ 2 >= 1 is required by Julia
    lin, stride = 0, 1
    for a, d in enumerate(dims):
        if cind[a] < 1:
            raise ValueError("grid dimension < 2: CartesianIndex(dims .÷ 2) is out of bounds in the reference")
        lin += (cind[a] - 1) * stride
        stride *= d
    covs = cov_eval(structs, cent - cent[lin][None, :])  # _pairwise(f, cdom, gdom) fftsim.jl:84-86
    C = covs.reshape(dims[::-1])                         # column-major reshape == C-order with reversed dims
    Cs = scipy.fft.fftshift(C)                           # fftsim.jl:90
    F = np.sqrt(np.abs(scipy.fft.fftn(Cs, workers=_workers())))
    F.reshape(-1)[0] = 0.0                               # fftsim.jl:91
    return F


def fftsim_sample(F: np.ndarray, w: np.ndarray, sill: float, mu: float,
                  inds: Optional[np.ndarray] = None) -> np.ndarray:
    """fftsim.jl:124-135 for one realization.  `w` uniform [0,1) with F's shape
    (C-order reversed dims).  Returns the flat column-major field (or its subset `inds`, 0-based)."""
    w = np.asarray(w, dtype=np.float64).reshape(F.shape)
    Wh = scipy.fft.fftn(w, workers=_workers())
    P = F * np.exp(1j * np.angle(Wh))                    # fftsim.jl:125
    Z = np.real(scipy.fft.ifftn(P, workers=_workers()))  # fftsim.jl:128
    n = Z.size
    s2 = np.sum(Z * Z) / (n - 1)                         # var(Z, mean=0), corrected (fftsim.jl:131)
    Z = np.sqrt(sill / s2) * Z + mu                      # fftsim.jl:132
    z = Z.reshape(-1)
    return z if inds is None else z[np.asarray(inds)]    # fftsim.jl:135


# ---------------------------------------------------------------------------------------------- conditional FFTSIM
def krige_neighbors_weights(structs, targets: np.ndarray, scoords: np.ndarray, maxneighbors: int):
    """GeoStatsModels.fitpredict's neighbourhood path for Kriging(f, mu) = simple Kriging, restated (the package is not
    vendored): per target the `maxneighbors` nearest samples (KNearestSearch, Euclidean, sorted by distance; ties -> lower
    sample index, i.e. a stable sort), covariance matrix C of those samples factorised with Cholesky, weights C^-1 c0.
    maxneighbors outside [1, nobs] is reset to nobs like fitpredict does.  Returns (nbr (n, k) int, lam (n, k))."""
    targets = np.atleast_2d(np.asarray(targets, dtype=np.float64))
    scoords = np.atleast_2d(np.asarray(scoords, dtype=np.float64))
    ns = scoords.shape[0]
    k = maxneighbors if 1 <= maxneighbors <= ns else ns
    n = targets.shape[0]
    nbr = np.zeros((n, k), dtype=np.int64)
    lam = np.zeros((n, k))
    for i in range(n):
        d2 = ((scoords - targets[i]) ** 2).sum(axis=1)
        idx = np.argsort(d2, kind="stable")[:k]
        X = scoords[idx]
        C = cov_eval(structs, (X[:, None, :] - X[None, :, :]).reshape(-1, X.shape[1])).reshape(k, k)
        c0 = cov_eval(structs, X - targets[i])
        cf = scipy.linalg.cho_factor(C, lower=True)
        nbr[i], lam[i] = idx, scipy.linalg.cho_solve(cf, c0)
    return nbr, lam


class FFTCond:
    """what fftsim.jl:94-106 keeps for conditional simulation: zbar and dinds (+ the restated weight table)"""

    def __init__(self, zbar, knodes, nbr, lam, mu):
        self.zbar, self.knodes, self.nbr, self.lam, self.mu = zbar, knodes, nbr, lam, mu


def fftsim_condition(structs, dims, origin, spacing, dcoords, dvals, knodes0, mu: float, maxneighbors: int = 26,
                     inds0: Optional[np.ndarray] = None) -> FFTCond:
    """fftsim.jl:94-104.  dcoords/dvals: the conditioning table where it is; knodes0: dinds = findall(mask), 0-based positions
    within the simulation domain sdom (= the grid, or its view `inds0`).  zbar = mu + sum lambda (z - mu) per element of sdom."""
    cent = grid_centroids(dims, origin, spacing)
    tg = cent if inds0 is None else cent[np.asarray(inds0)]
    nbr, lam = krige_neighbors_weights(structs, tg, dcoords, maxneighbors)
    dvals = np.asarray(dvals, dtype=np.float64)
    zbar = mu + (lam * (dvals[nbr] - mu)).sum(axis=1)
    knodes0 = np.asarray(knodes0)
    nbr2, lam2 = krige_neighbors_weights(structs, tg, tg[knodes0], maxneighbors)  # samples at the data nodes' centroids (:143)
    return FFTCond(zbar, knodes0, nbr2, lam2, mu)


def fftsim_sample_conditional(F: np.ndarray, w: np.ndarray, sill: float, cond: FFTCond, inds0: Optional[np.ndarray] = None) -> np.ndarray:
    """fftsim.jl:124-153: unconditional realization zu, Kriging of zu[dinds] (same neighbourhoods, same weights), z = zbar + (zu - zbaru)."""
    zu = fftsim_sample(F, w, sill, cond.mu, inds0)
    zk = zu[cond.knodes]
    zbaru = cond.mu + (cond.lam * (zk[cond.nbr] - cond.mu)).sum(axis=1)
    return cond.zbar + (zu - zbaru)


# ---------------------------------------------------------------------------------------------- ensemble statistics
def julia_quantile(v: np.ndarray, p: float) -> float:
    """Statistics.quantile(v, p) with Julia's defaults alpha = beta = 1 (definition 7), the call at src/ensembles.jl:50:
    sort; m = alpha + p (1 - alpha - beta); aleph = n p + m; j = clamp(trunc(aleph), 1, n-1); g = clamp(aleph - j, 0, 1);
    n == 1 ? v[1] : v[j] + g (v[j+1] - v[j])."""
    v = np.sort(np.asarray(v, dtype=np.float64))
    n = len(v)
    if not (0.0 <= p <= 1.0):
        raise ValueError("input probability out of [0,1] range")
    m = 1.0 + p * (1.0 - 1.0 - 1.0)
    aleph = n * p + m
    j = int(min(max(math.trunc(aleph), 1), n - 1)) if n > 1 else 1
    g = min(max(aleph - j, 0.0), 1.0)
    if n == 1:
        a = b = v[0]
    else:
        a, b = v[j - 1], v[j]
    return a + g * (b - a)


def ensemble_stats(Z: np.ndarray, x: float, ps: Sequence[float]):
    """src/ensembles.jl:42-52 on an (R, n) array of realizations (row = realization): mean, var (corrected), cdf(x) =
    count(<= x)/R, ccdf(x) = count(> x)/R and quantile(p) for p in ps, each per node.  Scalar loops like `ereduce`
    (ensembles.jl:76-85) for the quantile; use small n."""
    Z = np.asarray(Z, dtype=np.float64)
    R, n = Z.shape
    mean = Z.sum(axis=0) / R
    var = ((Z - mean) ** 2).sum(axis=0) / (R - 1) if R > 1 else np.full(n, np.nan)
    cdf = (Z <= x).sum(axis=0) / R
    ccdf = (Z > x).sum(axis=0) / R
    q = np.array([[julia_quantile(Z[:, i], p) for i in range(n)] for p in ps])
    return mean, var, cdf, ccdf, q

