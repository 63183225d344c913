"""Reference-facing host API for the GPU methods: GaussianProcess, LUSIM, FFTSIM, rand, Ensemble.

Mirrors (same names, argument meaning and error behaviour):
  src/processes/field/gaussian.jl:22-48   GaussianProcess, defaultschema
  src/processes/field.jl:43-58            initialize
  src/initialization/nearest.jl:12-34, explicit.jl:12-46   NearestInit / ExplicitInit
  src/simulation/field.jl:47-124,152-166  rand, defaultsimulation
  src/simulation/field/lusim.jl:38-126    LUSIM preprocess / randsingle (validity checks here, math on the GPU)
  src/simulation/field/fftsim.jl:54-139   FFTSIM preprocess / randsingle (unconditional path)
  src/ensembles.jl:10-85                  Ensemble
Everything numerical is a call into libgspb200 through _lib.py; nothing here computes fields on the CPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Union

import numpy as np

from . import _lib
from .domains import CartesianGrid, GeoTable, GridView, PointSet, georef
from .functions import GeoStatsFunction

_default_library: Optional[_lib.Library] = None


def default_library() -> _lib.Library:
    """The process-wide CUDA library/context (device 0 unless `set_devices` was called)."""
    global _default_library
    if _default_library is None:
        _default_library = _lib.Library()
    return _default_library


def set_devices(devices: Sequence[int]) -> _lib.Library:
    """Use these CUDA ordinals for subsequent `rand` calls (realizations are sharded over them)."""
    global _default_library
    if _default_library is not None:
        _default_library.close()
    _default_library = _lib.Library(devices=list(devices))
    return _default_library


# ------------------------------------------------------------------ process
class GaussianProcess:
    """GaussianProcess(func, mean=0) - gaussian.jl:22-35."""

    def __init__(self, func: GeoStatsFunction, mean=None):
        nf = func.nvariables()
        if mean is None:
            mean = np.zeros(nf) if nf > 1 else 0.0
        nm = np.size(mean)
        assert nm == nf, f"mean must have {nf} components, received {nm}"
        self.func = func
        self.mean = mean

    def mean_of(self, j: int) -> float:
        return float(np.atleast_1d(self.mean)[j])

    def defaultschema(self):
        nv = self.func.nvariables()
        return tuple(f"field{i + 1}" for i in range(nv)) if nv > 1 else ("field",)


# ------------------------------------------------------------------ initialization
class NearestInit:
    """nearest.jl:10-34: each datum goes to the nearest element, later data overwrite, NaN = missing.
    On a CartesianGrid with a CUDA library at hand the search runs on the device (gsp_nearest_init: grid arithmetic instead of
    the reference's KD-tree over all centroids); views and point sets are searched on the host like the reference does."""

    def apply(self, real, mask, dom, data: GeoTable, lib: Optional[_lib.Library] = None):
        dcoords = data.domain.centroids()
        if lib is not None and isinstance(dom, CartesianGrid) and dcoords.shape[0] > 0:
            for var in real:
                vals = np.array([np.nan if (v is None or (isinstance(v, float) and math.isnan(v))) else float(v) for v in data[var]])
                dinds0, z1 = lib.nearest_init(dom.dims, dom.origin, dom.spacing, dcoords, vals)
                real[var][dinds0] = z1
                mask[var][dinds0] = True
            return
        for i in range(dcoords.shape[0]):
            j = dom.nearest(dcoords[i])
            for var in real:
                v = data[var][i]
                if v is not None and not (isinstance(v, float) and math.isnan(v)):
                    real[var][j] = v
                    mask[var][j] = True


class ExplicitInit:
    """explicit.jl:12-46 with 1-based `orig` / `dest` like the reference."""

    def __init__(self, *args):
        if len(args) == 1:
            self.orig, self.dest = None, args[0]
        else:
            self.orig, self.dest = args

    def apply(self, real, mask, dom, data: GeoTable, lib=None):
        dest = list(self.dest)
        orig = list(range(1, data.domain.nelements() + 1)) if self.orig is None else list(self.orig)
        assert len(orig) == len(dest), "invalid explicit initialization"
        for i, j in zip(orig, dest):
            for var in real:
                v = data[var][i - 1]
                if v is not None and not (isinstance(v, float) and math.isnan(v)):
                    real[var][j - 1] = v
                    mask[var][j - 1] = True


def initialize(process: GaussianProcess, domain, data: Optional[GeoTable], init, lib: Optional[_lib.Library] = None):
    """field.jl:43-58 -> (real, mask) dicts keyed by variable name.  `lib`: lets NearestInit search on the device."""
    names = process.defaultschema() if data is None else data.names()
    n = domain.nelements()
    real = {v: np.zeros(n) for v in names}
    mask = {v: np.zeros(n, dtype=bool) for v in names}
    if data is not None:
        init.apply(real, mask, domain, data, lib)
    return real, mask


# ------------------------------------------------------------------ methods
class FieldSimulationMethod:
    pass


@dataclass
class LUSIM(FieldSimulationMethod):
    """GPU LUSIM (lusim.jl:36).  `library` selects the context (default: device 0).  `share_factor`: variables with the same marginal
    covariance and data nodes share one Cholesky factor (the results are identical to factoring twice)."""
    library: Optional[_lib.Library] = None
    share_factor: bool = True


@dataclass
class FFTSIM(FieldSimulationMethod):
    """GPU FFTSIM (fftsim.jl:47-52); neighbour options only matter for conditional simulation."""
    minneighbors: int = 1
    maxneighbors: int = 26
    neighborhood: object = None
    distance: object = None
    library: Optional[_lib.Library] = None


def defaultsimulation(process: GaussianProcess, domain, data=None):
    """field.jl:152-166 (SEQSIM is outside this engine's scope)."""
    f = process.func
    p = domain.parent()
    if isinstance(p, CartesianGrid) and f.isstationary() and f.nvariables() == 1 and f.range() <= min(p.sides()) / 3 and data is None:
        return FFTSIM()
    if domain.nelements() < 100 * 100 and f.isstationary() and f.issymmetric() and f.isbanded():
        return LUSIM()
    raise NotImplementedError("the reference would pick SEQSIM here; this engine provides LUSIM and FFTSIM only - pass method=")


def _domain_handle(domain):
    if isinstance(domain, CartesianGrid):
        return (_lib.make_grid_domain(domain.dims, domain.origin, domain.spacing), None)
    return _lib.make_point_domain(domain.centroids())


class _LUPre:
    def __init__(self, plans, names, rho):
        self.plans, self.names, self.rho = plans, names, rho


def preprocess_lusim(process: GaussianProcess, method: LUSIM, init, domain, data) -> _LUPre:
    """lusim.jl:38-110."""
    f = process.func
    if not (f.isstationary() and f.issymmetric() and f.isbanded()):
        raise ValueError("LUSIM requires a geostatistical function that is stationary, symmetric and banded. "
                         "Covariances or composite functions of covariances satisfy these properties.")
    lib = method.library or default_library()
    real, mask = initialize(process, domain, data, init, lib)
    names = tuple(real.keys())
    assert len(names) == f.nvariables(), "incompatible number of variables for geostatistical function"
    assert len(names) in (1, 2), "LUSIM only supports univariate and bivariate simulation"
    dom = _domain_handle(domain)
    plans, keys = [], []
    for j, var in enumerate(names):
        dinds0 = np.flatnonzero(mask[var])
        z1 = real[var][dinds0]
        marg = f.marginal(j)
        # lusim.jl:66-107 assembles and factors once per variable; when the marginal covariance and the data nodes of a variable equal
        # those of an earlier one (e.g. [1 rho; rho 1] * cov with shared data locations) the factor is shared and only d2 is computed
        key = ([(int(m[0]), float(m[1]), np.asarray(m[2], dtype=np.float64).tobytes(), float(m[3]) if len(m) > 3 else 0.0) for m in marg],
               dinds0.tobytes())
        like = next((plans[i] for i, k in enumerate(keys) if k == key), None) if method.share_factor else None
        plans.append(_lib.LUPlan(lib, marg, dom, dinds0 + 1 if len(dinds0) else None, z1 if len(dinds0) else None,
                                 process.mean_of(j), like=like))
        keys.append(key)
    if len(plans) == 2 and plans[0].Ns != plans[1].Ns:
        raise ValueError("DimensionMismatch: both variables must have the same number of simulation nodes (lusim.jl:164)")
    rho = f.rho() if len(names) == 2 else math.nan
    return _LUPre(plans, names, rho)


def rand_lusim(pre: _LUPre, nreals: int, rng, seed: int, resident: bool = False) -> Dict[str, np.ndarray]:
    """randsingle + _lusim (lusim.jl:112-175) for all realizations at once.  Returns var -> (N, R) arrays, or with
    `resident` var -> DeviceEnsemble (the fields stay in HBM)."""
    p1 = pre.plans[0]
    out = {}
    if resident:
        W1 = W2 = None
        nv = len(pre.plans)
        if rng is not None:
            W = rng.standard_normal((nreals, nv, p1.Ns))
            W1 = np.asfortranarray(W[:, 0, :].T)
            W2 = np.asfortranarray(W[:, 1, :].T) if nv == 2 else None
        out[pre.names[0]] = p1.sample_ensemble(nreals, W1, seed=seed, stream=0)
        if nv == 2:
            out[pre.names[1]] = pre.plans[1].sample_ensemble(nreals, W2, seed=seed, stream=1, rho=pre.rho, W1=W1)
        return out
    if rng is not None:
        # the reference's draw order: per realization w1 (Ns normals), then w2 (lusim.jl:160, randsingle :114-119)
        nv = len(pre.plans)
        W = rng.standard_normal((nreals, nv, p1.Ns))
        W1 = np.asfortranarray(W[:, 0, :].T)
        out[pre.names[0]] = p1.sample(nreals, W1)
        if nv == 2:
            W2 = np.asfortranarray(W[:, 1, :].T)
            out[pre.names[1]] = pre.plans[1].sample(nreals, W2, rho=pre.rho, W1=W1)
    else:
        out[pre.names[0]] = p1.sample(nreals, None, seed=seed, stream=0)
        if len(pre.plans) == 2:
            out[pre.names[1]] = pre.plans[1].sample(nreals, None, seed=seed, stream=1, rho=pre.rho)
    return out


class _FFTPre:
    def __init__(self, plan, var, inds1, sill):
        self.plan, self.var, self.inds1, self.sill = plan, var, inds1, sill


def preprocess_fftsim(process: GaussianProcess, method: FFTSIM, init, domain, data) -> _FFTPre:
    """fftsim.jl:54-107 (unconditional part)."""
    f = process.func
    assert f.isstationary(), "geostatistical function must be stationary"
    lib = method.library or default_library()
    real, mask = initialize(process, domain, data, init, lib)
    assert len(real) == 1, "FFTSIM does not support multivariate simulation"
    var = next(iter(real))
    grid = domain.parent()
    if not isinstance(grid, CartesianGrid):
        raise ValueError("FFTSIM requires a (view of a) CartesianGrid")
    plan = _lib.FFTPlan(lib, f.flat(), grid.dims, grid.origin, grid.spacing)
    inds1 = domain.parentindices()
    if data is not None:
        # fftsim.jl:94-104: zbar = simple Kriging of the data (where they are) onto sdom; dinds = findall(mask[var]);
        # the per-realization Kriging of fftsim.jl:140-149 is prepared here as a weight table (geometry only)
        if method.neighborhood is not None or method.distance is not None:
            raise NotImplementedError("FFTSIM conditioning on the GPU supports the default search only "
                                      "(k nearest neighbours, Euclidean distance)")
        vals = np.asarray(data[var], dtype=np.float64)
        keep = ~np.isnan(vals)
        dinds0 = np.flatnonzero(mask[var])
        plan.condition(process.mean_of(0), data.domain.centroids()[keep], vals[keep], dinds0 + 1, inds1,
                       minneighbors=method.minneighbors, maxneighbors=method.maxneighbors)
    return _FFTPre(plan, var, inds1, float(f.sill()))


def rand_fftsim(pre: _FFTPre, process: GaussianProcess, nreals: int, rng, seed: int, resident: bool = False) -> Dict[str, np.ndarray]:
    """fftsim.jl:109-139 for all realizations.  Returns var -> (n, R), or with `resident` var -> DeviceEnsemble."""
    w = None
    if rng is not None:
        w = rng.random((nreals, pre.plan.N))  # rand(rng, Float64, dims) per realization (fftsim.jl:124), column-major dims
    if resident:
        return {pre.var: pre.plan.sample_ensemble(nreals, w, seed=seed, sill=pre.sill, mu=process.mean_of(0), inds1=pre.inds1)}
    Z = pre.plan.sample(nreals, w, seed=seed, sill=pre.sill, mu=process.mean_of(0), inds1=pre.inds1)
    return {pre.var: Z.T}


# ------------------------------------------------------------------ ensemble
class Ensemble:
    """ensembles.jl:10-85.  `reals[var]` is an (n, R) host array, or a DeviceEnsemble when the realizations are resident
    on the GPUs (`rand(..., resident=True)`): then `e[i]` fetches one realization (the reference's `fetch` hook,
    ensembles.jl:16,27-31) and the statistics (ensembles.jl:42-52) run in HBM - only n-vectors come back.
    `e[i]` is the i-th realization (0-based) as a GeoTable."""

    def __init__(self, domain, reals: Dict[str, Union[np.ndarray, _lib.DeviceEnsemble]]):
        self.domain = domain
        self.reals = reals

    @staticmethod
    def _resident(a) -> bool:
        return isinstance(a, _lib.DeviceEnsemble)

    def __len__(self):
        a = next(iter(self.reals.values()))
        return a.R if self._resident(a) else a.shape[1]

    def __getitem__(self, i):
        if isinstance(i, (list, tuple, np.ndarray, range)):
            return [self[k] for k in i]
        if i < 0 or i >= len(self):
            raise IndexError(i)
        return georef({v: (a.fetch(i, 1)[0] if self._resident(a) else a[:, i]) for v, a in self.reals.items()}, self.domain)

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def variables(self):
        return tuple(self.reals.keys())

    def _reduce(self, host_fn, dev_fn):
        return georef({v: (dev_fn(a) if self._resident(a) else host_fn(a)) for v, a in self.reals.items()}, self.domain)

    def mean(self):
        return self._reduce(lambda a: a.mean(axis=1), lambda d: d.mean())

    def var(self):
        return self._reduce(lambda a: a.var(axis=1, ddof=1), lambda d: d.var())

    def cdf(self, x: float):
        return self._reduce(lambda a: (a <= x).mean(axis=1), lambda d: d.cdf(x))

    def ccdf(self, x: float):
        return self._reduce(lambda a: (a > x).mean(axis=1), lambda d: d.ccdf(x))

    def quantile(self, p):
        if np.ndim(p) > 0:
            return [self.quantile(q) for q in p]
        return self._reduce(lambda a: np.quantile(a, p, axis=1), lambda d: d.quantile([p])[0])

    def close(self):
        """release the device memory of a resident ensemble (also done when the object is collected)"""
        for a in self.reals.values():
            if self._resident(a):
                a.close()

    def __repr__(self):
        return f"{self.domain.ndim}D Ensemble\n  domain:    {self.domain}\n  variables: {', '.join(self.variables())}\n  N° reals:  {len(self)}"


def merge_moments(parts):
    """Chan's update over per-rank partials [(count, mean, m2), ...] (gsp_ensemble_moments of every rank's shard, gathered
    with torch.distributed.all_gather_object or an all-gather of device tensors): -> (R, mean, var) of the whole ensemble."""
    cnt, mean, m2 = parts[0]
    mean, m2 = np.array(mean, dtype=np.float64), np.array(m2, dtype=np.float64)
    for c, m, q in parts[1:]:
        tot = cnt + c
        d = np.asarray(m) - mean
        mean = mean + d * (c / tot)
        m2 = m2 + np.asarray(q) + d * d * (cnt * c / tot)
        cnt = tot
    return cnt, mean, m2 / (cnt - 1)


def mean(process: GaussianProcess, domain, *, data: Optional[GeoTable] = None, init=None, minneighbors: int = 1, maxneighbors: int = 10,
         neighborhood=None, distance=None, library: Optional[_lib.Library] = None) -> GeoTable:
    """mean(process, domain; data, init, kwargs...) - src/expectation/field.jl:43-44.
    Without data: priormean (expectation/field/gaussian.jl:5-19), the process mean on every element.  With data: posteriormean
    (gaussian.jl:21-25) = simple Kriging of the data where they are onto the domain, GeoStatsModels.fitpredict's defaults
    (k nearest neighbours, maxneighbors = 10, Euclidean) - computed on the device by the Kriging kernel that conditions FFTSIM
    (gsp_fft_plan_condition -> gsp_fft_plan_condmean).  Grids and views of grids, univariate functions; `init` places nothing here
    (the Kriging uses the data's own coordinates) and is accepted for signature parity."""
    f = process.func
    if data is None:
        names = process.defaultschema()
        n = domain.nelements()
        return georef({v: np.full(n, process.mean_of(j)) for j, v in enumerate(names)}, domain)
    if neighborhood is not None or distance is not None:
        raise NotImplementedError("posterior mean on the GPU supports the default search only (k nearest neighbours, Euclidean distance)")
    assert f.nvariables() == 1, "the posterior mean is offloaded for univariate functions only"
    grid = domain.parent()
    if not isinstance(grid, CartesianGrid):
        raise ValueError("the posterior mean on the GPU requires a (view of a) CartesianGrid")
    lib = library or default_library()
    names = data.names()
    plan = _lib.FFTPlan(lib, f.flat(), grid.dims, grid.origin, grid.spacing)
    try:
        out = {}
        inds1 = domain.parentindices()
        for var in names:
            vals = np.asarray(data[var], dtype=np.float64)
            keep = ~np.isnan(vals)
            real, mask = initialize(process, domain, GeoTable({var: vals}, data.domain), init or NearestInit(), lib)
            knodes0 = np.flatnonzero(mask[var])
            plan.condition(process.mean_of(0), data.domain.centroids()[keep], vals[keep], knodes0 + 1, inds1, minneighbors=minneighbors,
                           maxneighbors=maxneighbors)
            out[var] = plan.condmean()
    finally:
        plan.close()
    return georef(out, domain)


def expectedvalue(process: GaussianProcess, domain, **kwargs) -> GeoTable:
    """expectedvalue(process, domain; kwargs...) - src/expectation/field.jl:17-23: Gaussian processes are continuous and analytical,
    so this is `mean`."""
    return mean(process, domain, **kwargs)


def _fresh_seed() -> int:
    """64 fresh bits per unseeded call (os entropy through numpy's SeedSequence)."""
    return int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0])


def rand(process: GaussianProcess, domain, nreals: Optional[int] = None, *, rng: Union[None, int, np.random.Generator] = None,
         data: Optional[GeoTable] = None, method: Optional[FieldSimulationMethod] = None, init=None, resident: bool = False):
    """rand([rng], process, domain, [n]; data, method, init) - field.jl:47-124.

    rng: a numpy Generator -> noise is drawn on the host in the reference's order and injected
    (parity mode); an int seed -> on-device counter RNG, reproducible; None -> on-device counter RNG with a FRESH
    64-bit seed per call (the reference draws from Random.default_rng(): successive calls are independent).
    Without `nreals` a single GeoTable is returned, with it an Ensemble.  resident=True keeps the realizations on the
    GPUs (Ensemble with the `fetch` hook, ensembles.jl:16): statistics run in HBM, `e[i]` downloads one realization."""
    init = init or NearestInit()
    smethod = method if method is not None else defaultsimulation(process, domain, data)
    gen = rng if isinstance(rng, np.random.Generator) else None
    if isinstance(rng, (int, np.integer)):
        seed = int(rng) & 0xFFFFFFFFFFFFFFFF
    else:
        seed = _fresh_seed()  # unused when a Generator injects the noise
    n = 1 if nreals is None else int(nreals)
    if isinstance(smethod, LUSIM):
        pre = preprocess_lusim(process, smethod, init, domain, data)
        reals = rand_lusim(pre, n, gen, seed, resident and nreals is not None)
        for p in pre.plans:
            p.close()
    elif isinstance(smethod, FFTSIM):
        pre = preprocess_fftsim(process, smethod, init, domain, data)
        reals = rand_fftsim(pre, process, n, gen, seed, resident and nreals is not None)
        pre.plan.close()
    else:
        raise TypeError(f"unsupported simulation method {smethod!r}")
    ens = Ensemble(domain, reals)
    return ens[0] if nreals is None else ens
