"""ctypes binding of the C ABI declared in include/gsp_b200.h (libgspb200.so).

This is the stand-in for the Julia `ccall` glue (julia/GeoStatsProcessesB200.jl):
Julia is not available in this image, so the reference-facing host layer is mirrored in Python.
There is NO CPU fallback: if the CUDA library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSP_B200_LIB selects another BUILD of the same CUDA library (A/B experiments); there is still no non-CUDA alternative
DEFAULT_LIB = os.environ.get("GSP_B200_LIB") or os.path.join(_HERE, "libgspb200.so")

GSP_E_CUDA, GSP_E_UNSUPPORTED, GSP_E_NOMEM, GSP_E_STATE = -1001, -1002, -1003, -1004
MAX_STRUCTS = 8


class GspError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gsp_b200 error {code}: {msg}")
        self.code = code


class PosDefException(ArithmeticError):
    """Mirror of LinearAlgebra.PosDefException thrown by `cholesky` (lusim.jl:92,98,103)."""

    def __init__(self, info: int):
        super().__init__(f"matrix is not positive definite; Cholesky factorization failed (info={info})")
        self.info = info


class _Structure(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("sill", C.c_double), ("A", C.c_double * 9), ("param", C.c_double)]


class _CovModel(C.Structure):
    _fields_ = [("nstruct", C.c_int32), ("reserved", C.c_int32), ("structs", C.POINTER(_Structure))]


class _Domain(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("nelems", C.c_int64), ("coords", C.POINTER(C.c_double)),
                ("dims", C.c_int64 * 3), ("origin", C.c_double * 3), ("spacing", C.c_double * 3)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol of include/gsp_b200.h (tests check this)
SIGNATURES = {
    "gsp_version": (C.c_char_p, []),
    "gsp_ctx_create": (C.c_int, [C.c_int32, C.POINTER(C.c_int32), C.POINTER(_vp)]),
    "gsp_ctx_destroy": (C.c_int, [_vp]),
    "gsp_last_error": (C.c_char_p, [_vp]),
    "gsp_ctx_ndev": (C.c_int, [_vp]),
    "gsp_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_int64]),
    "gsp_host_free": (C.c_int, [_vp]),
    "gsp_pairwise": (C.c_int, [_vp, C.POINTER(_CovModel), C.c_int32, C.c_int64, _vp, C.c_int64, _vp, _vp]),
    "gsp_potrf": (C.c_int, [_vp, C.c_int64, _vp]),
    "gsp_nearest_init": (C.c_int, [_vp, C.POINTER(_Domain), C.c_int64, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64)]),
    "gsp_lu_plan_create": (C.c_int, [_vp, C.POINTER(_CovModel), C.POINTER(_Domain), C.c_int64, _vp, _vp, C.c_double, C.POINTER(_vp)]),
    "gsp_lu_plan_create_like": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_double, C.POINTER(_vp)]),
    "gsp_lu_plan_destroy": (C.c_int, [_vp]),
    "gsp_lu_plan_sizes": (C.c_int, [_vp, C.POINTER(C.c_int64 * 3)]),
    "gsp_lu_plan_times": (C.c_int, [_vp, C.POINTER(C.c_double * 3)]),
    "gsp_lu_plan_get": (C.c_int, [_vp, _vp, _vp]),
    "gsp_lu_sample": (C.c_int, [_vp, C.c_int64, _vp, C.c_uint64, C.c_int32, C.c_int64, C.c_double, _vp, _vp]),
    "gsp_lu_sample_dev": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, C.c_uint64, C.c_int32, C.c_int64, C.c_double, _vp, _vp, C.c_int64]),
    "gsp_fft_plan_create": (C.c_int, [_vp, C.POINTER(_CovModel), C.POINTER(_Domain), C.POINTER(_vp)]),
    "gsp_fft_plan_destroy": (C.c_int, [_vp]),
    "gsp_fft_plan_get": (C.c_int, [_vp, _vp]),
    "gsp_fft_sample": (C.c_int, [_vp, C.c_int64, _vp, C.c_uint64, C.c_int64, C.c_double, C.c_double, C.c_int64, _vp, _vp]),
    "gsp_fft_sample_dev": (C.c_int, [_vp, C.c_int64, _vp, C.c_uint64, C.c_int64, C.c_double, C.c_double, C.c_int64, _vp, _vp]),
    "gsp_fft_plan_condition": (C.c_int, [_vp, C.c_double, C.c_int32, C.c_int32, C.c_int64, _vp, _vp, C.c_int64, _vp, C.c_int64, _vp]),
    "gsp_fft_plan_condmean": (C.c_int, [_vp, _vp]),
    "gsp_ensemble_create": (C.c_int, [_vp, C.c_int64, C.c_int64, C.POINTER(_vp)]),
    "gsp_ensemble_destroy": (C.c_int, [_vp]),
    "gsp_ensemble_sizes": (C.c_int, [_vp, C.POINTER(C.c_int64 * 2)]),
    "gsp_ensemble_put": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp]),
    "gsp_ensemble_fetch": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp]),
    "gsp_fft_sample_ensemble": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_int64, C.c_double, C.c_double, C.c_int64, _vp]),
    "gsp_lu_sample_ensemble": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_int32, C.c_int64, C.c_double, _vp]),
    "gsp_ensemble_mean": (C.c_int, [_vp, _vp]),
    "gsp_ensemble_var": (C.c_int, [_vp, _vp]),
    "gsp_ensemble_cdf": (C.c_int, [_vp, C.c_double, _vp]),
    "gsp_ensemble_ccdf": (C.c_int, [_vp, C.c_double, _vp]),
    "gsp_ensemble_quantile": (C.c_int, [_vp, C.c_int64, _vp, _vp]),
    "gsp_ensemble_moments": (C.c_int, [_vp, _vp, _vp]),
    "gsp_profile_enable": (C.c_int, [_vp, C.c_int32]),
    "gsp_profile_read": (C.c_int64, [_vp, C.c_char_p, C.c_int64]),
    "gsp_kernel_launches": (C.c_int64, []),
    "gsp_last_sample_ms": (C.c_double, [_vp]),
}


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_vp)


def make_cov(structs: Sequence[tuple]) -> tuple:
    """structs: sequence of (kind, sill, A 3x3 row-major ndarray[, param]); param = Matern order.  Returns (model, keepalive)."""
    n = len(structs)
    arr = (_Structure * n)()
    for i, st in enumerate(structs):
        kind, sill, A = st[0], st[1], st[2]
        arr[i].param = float(st[3]) if len(st) > 3 else 0.0
        arr[i].kind = int(kind)
        arr[i].sill = float(sill)
        A = np.asarray(A, dtype=np.float64).reshape(3, 3)
        for k in range(9):
            arr[i].A[k] = float(A.flat[k])
    m = _CovModel(n, 0, C.cast(arr, C.POINTER(_Structure)))
    return m, arr


def make_grid_domain(dims, origin, spacing) -> _Domain:
    d = _Domain()
    d.kind, d.dim = 1, len(dims)
    n = 1
    for a in range(3):
        d.dims[a] = int(dims[a]) if a < len(dims) else 1
        d.origin[a] = float(origin[a]) if a < len(dims) else 0.0
        d.spacing[a] = float(spacing[a]) if a < len(dims) else 1.0
        n *= d.dims[a]
    d.nelems = n
    d.coords = None
    return d


def make_point_domain(coords: np.ndarray) -> tuple:
    """coords: (n, dim) array of centroids.  Returns (domain, keepalive)."""
    X = np.ascontiguousarray(np.asarray(coords, dtype=np.float64))  # (n, dim) C-order == dim x n column-major
    d = _Domain()
    d.kind, d.dim, d.nelems = 0, X.shape[1], X.shape[0]
    d.coords = X.ctypes.data_as(_dp)
    for a in range(3):
        d.dims[a] = 1
    return d, X


class Library:
    """A loaded libgspb200 with one context.  `path=None` loads the in-tree CUDA build and raises if
    it is absent (the product never substitutes anything else)."""

    def __init__(self, path: Optional[str] = None, devices: Optional[Sequence[int]] = None):
        self.path = path or DEFAULT_LIB
        if not os.path.exists(self.path):
            raise FileNotFoundError(
                f"{self.path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'); "
                "there is no CPU fallback")
        self.lib = C.CDLL(self.path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.lib, name)  # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        devices = [0] if devices is None else list(devices)
        arr = (C.c_int32 * len(devices))(*devices)
        ctx = _vp()
        rc = self.lib.gsp_ctx_create(len(devices), arr, C.byref(ctx))
        if rc != 0:
            raise GspError(rc, "gsp_ctx_create failed (is a CUDA device visible?)")
        self.ctx = ctx
        self.devices = devices

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.gsp_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def version(self) -> str:
        return self.lib.gsp_version().decode()

    def check(self, rc: int, posdef: bool = False):
        if rc == 0:
            return
        if rc > 0 and posdef:
            raise PosDefException(rc)
        msg = self.lib.gsp_last_error(self.ctx).decode(errors="replace")
        if -32 <= rc < 0:
            raise ValueError(f"gsp_b200: invalid argument {-rc}: {msg}")
        raise GspError(rc, msg)

    def kernel_launches(self) -> int:
        return int(self.lib.gsp_kernel_launches())

    def profile_enable(self, on: bool = True):
        self.check(self.lib.gsp_profile_enable(self.ctx, 1 if on else 0))

    def profile_read(self) -> dict:
        import json
        buf = C.create_string_buffer(1 << 16)
        n = self.lib.gsp_profile_read(self.ctx, buf, len(buf))
        if n < 0:
            raise GspError(int(n), "profile buffer too small")
        return json.loads(buf.value.decode())

    def last_sample_ms(self) -> float:
        return float(self.lib.gsp_last_sample_ms(self.ctx))

    # ------------------------------------------------------------------ a1 / a2 entry points
    def pairwise(self, structs, X1: np.ndarray, X2: Optional[np.ndarray] = None) -> np.ndarray:
        X1 = np.ascontiguousarray(np.asarray(X1, dtype=np.float64))
        n1, dim = X1.shape
        X2c = None if X2 is None else np.ascontiguousarray(np.asarray(X2, dtype=np.float64))
        n2 = n1 if X2c is None else X2c.shape[0]
        out = np.empty((n1, n2), dtype=np.float64, order="F")
        m, keep = make_cov(structs)
        self.check(self.lib.gsp_pairwise(self.ctx, C.byref(m), dim, n1, _ptr(X1), n2, _ptr(X2c), _ptr(out)))
        return out

    def nearest_init(self, dims, origin, spacing, dcoords: np.ndarray, dvals: np.ndarray):
        """initialize + NearestInit on a CartesianGrid (nearest.jl:12-34) on the device -> (dinds0 ascending 0-based, z1)."""
        if len(dvals) == 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0)
        X = np.ascontiguousarray(np.asarray(dcoords, dtype=np.float64).reshape(len(dvals), -1))
        v = np.ascontiguousarray(dvals, dtype=np.float64)
        dom = make_grid_domain(dims, origin, spacing)
        dinds = np.empty(max(len(v), 1), dtype=np.int64)
        z1 = np.empty(max(len(v), 1))
        cnt = C.c_int64(0)
        self.check(self.lib.gsp_nearest_init(self.ctx, C.byref(dom), len(v), _ptr(X), _ptr(v), _ptr(dinds), _ptr(z1), C.byref(cnt)))
        n = int(cnt.value)
        return dinds[:n] - 1, z1[:n].copy()

    def potrf(self, A: np.ndarray) -> np.ndarray:
        L = np.array(A, dtype=np.float64, order="F", copy=True)
        n = L.shape[0]
        self.check(self.lib.gsp_potrf(self.ctx, n, _ptr(L)), posdef=True)
        return L


class LUPlan:
    def __init__(self, lib: Library, structs, domain, dinds1: Optional[np.ndarray], z1: Optional[np.ndarray], mu: float,
                 like: Optional["LUPlan"] = None):
        """`like`: a plan with the same marginal covariance and data nodes - its factor is shared (gsp_lu_plan_create_like), only
        d2 is computed; `structs` / `domain` are then ignored."""
        self.lib = lib
        dinds1 = np.zeros(0, dtype=np.int64) if dinds1 is None else np.ascontiguousarray(dinds1, dtype=np.int64)
        z1 = np.zeros(0) if z1 is None else np.ascontiguousarray(z1, dtype=np.float64)
        h = _vp()
        if like is not None:
            rc = lib.lib.gsp_lu_plan_create_like(like.h, len(dinds1), _ptr(dinds1) if len(dinds1) else None, _ptr(z1) if len(z1) else None,
                                                 float(mu), C.byref(h))
            lib.check(rc)
        else:
            m, keep = make_cov(structs)
            dom, keep2 = domain
            rc = lib.lib.gsp_lu_plan_create(lib.ctx, C.byref(m), C.byref(dom), len(dinds1), _ptr(dinds1) if len(dinds1) else None,
                                            _ptr(z1) if len(z1) else None, float(mu), C.byref(h))
            lib.check(rc, posdef=True)
        self.h = h
        sizes = (C.c_int64 * 3)()
        lib.check(lib.lib.gsp_lu_plan_sizes(h, C.byref(sizes)))
        self.N, self.Nd, self.Ns = int(sizes[0]), int(sizes[1]), int(sizes[2])

    def times(self):
        """device ms of (assembly, Cholesky, d2 solve)"""
        ms = (C.c_double * 3)()
        self.lib.check(self.lib.lib.gsp_lu_plan_times(self.h, C.byref(ms)))
        return float(ms[0]), float(ms[1]), float(ms[2])

    def get(self):
        d2 = np.empty(self.Ns)
        L22 = np.empty((self.Ns, self.Ns), order="F")
        self.lib.check(self.lib.lib.gsp_lu_plan_get(self.h, _ptr(d2), _ptr(L22)))
        return d2, L22

    def sample(self, R: int, W: Optional[np.ndarray] = None, seed: int = 0, stream: int = 0, first_real: int = 0,
               rho: float = math.nan, W1: Optional[np.ndarray] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        if W is not None:
            W = np.asfortranarray(W, dtype=np.float64).reshape(self.Ns, R, order="F")
        if W1 is not None:
            W1 = np.asfortranarray(W1, dtype=np.float64).reshape(self.Ns, R, order="F")
        Z = np.empty((self.N, R), order="F") if out is None else out
        rc = self.lib.lib.gsp_lu_sample(self.h, R, _ptr(W), seed, stream, first_real, float(rho), _ptr(W1), _ptr(Z))
        self.lib.check(rc)
        return Z

    def sample_ensemble(self, R: int, W: Optional[np.ndarray] = None, seed: int = 0, stream: int = 0, first_real: int = 0,
                        rho: float = math.nan, W1: Optional[np.ndarray] = None, ens: Optional["DeviceEnsemble"] = None) -> "DeviceEnsemble":
        """like `sample`, but the realizations stay on the devices (`ens`: refill an existing ensemble of the same shape)"""
        if W is not None:
            W = np.asfortranarray(W, dtype=np.float64).reshape(self.Ns, R, order="F")
        if W1 is not None:
            W1 = np.asfortranarray(W1, dtype=np.float64).reshape(self.Ns, R, order="F")
        if ens is None:
            ens = DeviceEnsemble(self.lib, self.N, R)
        rc = self.lib.lib.gsp_lu_sample_ensemble(self.h, ens.h, _ptr(W), seed, stream, first_real, float(rho), _ptr(W1))
        self.lib.check(rc)
        return ens

    def sample_dev(self, R, W_ptr, ldw, seed, stream, first_real, rho, W1_ptr, Z_ptr, ldz):
        rc = self.lib.lib.gsp_lu_sample_dev(self.h, R, W_ptr, ldw, seed, stream, first_real, float(rho), W1_ptr, Z_ptr, ldz)
        self.lib.check(rc)

    def close(self):
        if getattr(self, "h", None):
            self.lib.lib.gsp_lu_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceEnsemble:
    """R realizations of n values resident on the context's devices (gsp_ensemble_*): the `reals` + `fetch` of
    Ensemble(domain, reals; fetch) (src/ensembles.jl:10-16) with the statistics of ensembles.jl:42-52 computed in HBM."""

    def __init__(self, lib: Library, n: int, R: int):
        self.lib, self.n, self.R = lib, int(n), int(R)
        h = _vp()
        lib.check(lib.lib.gsp_ensemble_create(lib.ctx, self.n, self.R, C.byref(h)))
        self.h = h

    def put(self, Z: np.ndarray, r0: int = 0):
        """Z: (nr, n) C-order (row r = realization r0 + r)."""
        Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(-1, self.n)
        self.lib.check(self.lib.lib.gsp_ensemble_put(self.h, r0, Z.shape[0], _ptr(Z)))

    def fetch(self, r0: int = 0, nr: Optional[int] = None) -> np.ndarray:
        nr = self.R - r0 if nr is None else nr
        Z = np.empty((nr, self.n))
        self.lib.check(self.lib.lib.gsp_ensemble_fetch(self.h, r0, nr, _ptr(Z)))
        return Z

    def _vec(self, fn, *args) -> np.ndarray:
        out = np.empty(self.n)
        self.lib.check(fn(self.h, *args, _ptr(out)))
        return out

    def mean(self) -> np.ndarray:
        return self._vec(self.lib.lib.gsp_ensemble_mean)

    def var(self) -> np.ndarray:
        return self._vec(self.lib.lib.gsp_ensemble_var)

    def cdf(self, x: float) -> np.ndarray:
        return self._vec(self.lib.lib.gsp_ensemble_cdf, float(x))

    def ccdf(self, x: float) -> np.ndarray:
        return self._vec(self.lib.lib.gsp_ensemble_ccdf, float(x))

    def quantile(self, ps) -> np.ndarray:
        """ps: sequence of probabilities -> (len(ps), n)."""
        ps = np.ascontiguousarray(np.atleast_1d(ps), dtype=np.float64)
        out = np.empty((len(ps), self.n))
        self.lib.check(self.lib.lib.gsp_ensemble_quantile(self.h, len(ps), _ptr(ps), _ptr(out)))
        return out

    def moments(self):
        mean, m2 = np.empty(self.n), np.empty(self.n)
        self.lib.check(self.lib.lib.gsp_ensemble_moments(self.h, _ptr(mean), _ptr(m2)))
        return mean, m2

    def close(self):
        if getattr(self, "h", None):
            self.lib.lib.gsp_ensemble_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FFTPlan:
    def __init__(self, lib: Library, structs, dims, origin, spacing):
        self.lib = lib
        m, keep = make_cov(structs)
        dom = make_grid_domain(dims, origin, spacing)
        h = _vp()
        lib.check(lib.lib.gsp_fft_plan_create(lib.ctx, C.byref(m), C.byref(dom), C.byref(h)))
        self.h = h
        self.dims = tuple(int(d) for d in dims)
        self.N = int(np.prod(self.dims))

    def condition(self, mu: float, dcoords: np.ndarray, dvals: np.ndarray, knodes1: np.ndarray, inds1: Optional[np.ndarray] = None,
                  minneighbors: int = 1, maxneighbors: int = 26):
        """fftsim.jl:94-101: dcoords (nd, dim), dvals (nd); knodes1 = findall(mask) within the simulation domain (1-based);
        inds1 = parentindices of the view (None: whole grid).  Afterwards `sample*` returns conditional realizations."""
        X = np.ascontiguousarray(np.asarray(dcoords, dtype=np.float64).reshape(len(dvals), -1))
        v = np.ascontiguousarray(dvals, dtype=np.float64)
        kn = np.ascontiguousarray(knodes1, dtype=np.int64)
        ii = None if inds1 is None else np.ascontiguousarray(inds1, dtype=np.int64)
        rc = self.lib.lib.gsp_fft_plan_condition(self.h, float(mu), int(minneighbors), int(maxneighbors), len(v), _ptr(X), _ptr(v), len(kn),
                                                 _ptr(kn), 0 if ii is None else len(ii), _ptr(ii))
        self.lib.check(rc)
        self.cond_n = self.N if ii is None else len(ii)

    def condmean(self) -> np.ndarray:
        z = np.empty(getattr(self, "cond_n", self.N))
        self.lib.check(self.lib.lib.gsp_fft_plan_condmean(self.h, _ptr(z)))
        return z

    def spectrum(self) -> np.ndarray:
        F = np.empty(self.N)
        self.lib.check(self.lib.lib.gsp_fft_plan_get(self.h, _ptr(F)))
        return F.reshape(self.dims[::-1])

    def sample(self, R: int, w: Optional[np.ndarray] = None, seed: int = 0, first_real: int = 0, sill: float = 1.0, mu: float = 0.0,
               inds1: Optional[np.ndarray] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        if w is not None:
            w = np.ascontiguousarray(w, dtype=np.float64).reshape(R, self.N)
        n_inds = 0 if inds1 is None else len(inds1)
        inds1 = None if inds1 is None else np.ascontiguousarray(inds1, dtype=np.int64)
        nout = n_inds if n_inds else self.N
        Z = np.empty((R, nout)) if out is None else out  # C-order (R, nout) == column-major nout x R
        rc = self.lib.lib.gsp_fft_sample(self.h, R, _ptr(w), seed, first_real, float(sill), float(mu), n_inds, _ptr(inds1), _ptr(Z))
        self.lib.check(rc)
        return Z

    def sample_ensemble(self, R: int, w: Optional[np.ndarray] = None, seed: int = 0, first_real: int = 0, sill: float = 1.0,
                        mu: float = 0.0, inds1: Optional[np.ndarray] = None, ens: Optional["DeviceEnsemble"] = None) -> "DeviceEnsemble":
        """like `sample`, but the realizations stay on the devices (`ens`: refill an existing ensemble of the same shape)"""
        if w is not None:
            w = np.ascontiguousarray(w, dtype=np.float64).reshape(R, self.N)
        n_inds = 0 if inds1 is None else len(inds1)
        inds1 = None if inds1 is None else np.ascontiguousarray(inds1, dtype=np.int64)
        if ens is None:
            ens = DeviceEnsemble(self.lib, n_inds if n_inds else self.N, R)
        rc = self.lib.lib.gsp_fft_sample_ensemble(self.h, ens.h, _ptr(w), seed, first_real, float(sill), float(mu), n_inds, _ptr(inds1))
        self.lib.check(rc)
        return ens

    def sample_dev(self, R, w_ptr, seed, first_real, sill, mu, n_inds, inds_ptr, out_ptr):
        rc = self.lib.lib.gsp_fft_sample_dev(self.h, R, w_ptr, seed, first_real, float(sill), float(mu), n_inds, inds_ptr, out_ptr)
        self.lib.check(rc)

    def close(self):
        if getattr(self, "h", None):
            self.lib.lib.gsp_fft_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
