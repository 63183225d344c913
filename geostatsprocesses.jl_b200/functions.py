"""Host-side mirror of the GeoStatsFunctions objects the LUSIM/FFTSIM path consumes.

Only what crosses the boundary is modelled: nested structures (kind, contribution matrix,
anisotropy metric), `sill`, `range`, `nvariables`, `isstationary/issymmetric/isbanded`
(lusim.jl:44, fftsim.jl:60, field.jl:159-161), `structures` (lusim.jl:133) and C(0) (lusim.jl:140).
The formulas themselves are evaluated on the GPU (csrc/cov.cuh).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

NUGGET, SPHERICAL, EXPONENTIAL, GAUSSIAN, CUBIC, PENTASPHERICAL, SINEHOLE, CIRCULAR, MATERN = 0, 1, 2, 3, 4, 5, 6, 7, 8
_NAMES = {MATERN: "Matern", NUGGET: "NuggetEffect", SPHERICAL: "Spherical", EXPONENTIAL: "Exponential", GAUSSIAN: "Gaussian", CUBIC: "Cubic",
          PENTASPHERICAL: "Pentaspherical", SINEHOLE: "SineHole", CIRCULAR: "Circular"}


def metric_matrix(range: float = 1.0, ranges: Optional[Sequence[float]] = None, rotation=None) -> np.ndarray:
    """3x3 row-major A with u = |A @ delta|: isotropic -> I/range; MetricBall(radii, R) -> diag(1/radii) @ R'."""
    A = np.zeros((3, 3))
    if ranges is None:
        np.fill_diagonal(A, 1.0 / float(range))
        return A
    radii = np.asarray(ranges, dtype=np.float64)
    d = len(radii)
    R = np.eye(d) if rotation is None else np.asarray(rotation, dtype=np.float64)
    if np.ndim(R) == 0:  # 2-D rotation angle (radians, counter-clockwise)
        c, s = np.cos(float(R)), np.sin(float(R))
        R = np.array([[c, -s], [s, c]])
    A[:d, :d] = np.diag(1.0 / radii) @ R.T
    return A


@dataclass
class _Struct:
    kind: int
    C: np.ndarray          # nv x nv contribution
    A: np.ndarray          # 3x3 metric
    maxrange: float
    param: float = 0.0     # Matern: order nu

    def flat(self, sill: float) -> tuple:
        return (self.kind, sill, self.A, self.param) if self.kind == MATERN else (self.kind, sill, self.A)


@dataclass
class GeoStatsFunction:
    """A (possibly nested, possibly multivariate) stationary covariance or variogram."""
    structs: List[_Struct] = field(default_factory=list)
    variogram: bool = False
    __array_ufunc__ = None  # let `ndarray * f` dispatch to __rmul__

    # --- traits used by the reference's checks
    def nvariables(self) -> int:
        return int(self.structs[0].C.shape[0])

    def isstationary(self) -> bool:
        return True

    def issymmetric(self) -> bool:
        return all(np.allclose(s.C, s.C.T) for s in self.structs)

    def isbanded(self) -> bool:
        return not self.variogram  # covariances are banded, variograms are not (utils.jl:53)

    def sill(self):
        S = sum(s.C for s in self.structs)
        return float(S[0, 0]) if self.nvariables() == 1 else S

    def range(self) -> float:
        return max((s.maxrange for s in self.structs if s.kind != NUGGET), default=0.0)

    def at_zero(self) -> np.ndarray:
        """cov(0) as a matrix (lusim.jl:140)."""
        return sum(s.C for s in self.structs)

    # --- algebra: c * f, M * f, f + g
    def __rmul__(self, c):
        c = np.asarray(c, dtype=np.float64)
        out = []
        for s in self.structs:
            if c.ndim == 0:
                C = float(c) * s.C
            else:
                if s.C.shape != (1, 1):
                    raise ValueError("matrix scaling needs a univariate function")
                C = c * s.C[0, 0]
            out.append(_Struct(s.kind, np.atleast_2d(C), s.A, s.maxrange, s.param))
        return GeoStatsFunction(out, self.variogram)

    def __add__(self, other: "GeoStatsFunction"):
        if self.variogram != other.variogram:
            raise ValueError("cannot add a variogram and a covariance")
        if self.nvariables() != other.nvariables():
            raise ValueError("incompatible number of variables")
        return GeoStatsFunction(self.structs + other.structs, self.variogram)

    # --- boundary flattening
    def marginal(self, j: int) -> list:
        """_marginalize (lusim.jl:132-137) flattened to [(kind, sill, A[, order])]."""
        out = [s.flat(float(s.C[j, j])) for s in self.structs if not (s.kind == NUGGET and s.C[j, j] == 0.0)]
        return out

    def flat(self) -> list:
        if self.nvariables() != 1:
            raise ValueError("univariate function expected")
        return self.marginal(0)

    def rho(self) -> float:
        """_rho (lusim.jl:139-143)."""
        C0 = self.at_zero()
        S = self.sill()
        return float(C0[0, 1] / np.sqrt(S[0, 0] * S[1, 1]))

    def __repr__(self):
        kind = "Variogram" if self.variogram else "Covariance"
        parts = [f"{_NAMES[s.kind]}{'' if s.kind == NUGGET else kind}(sill={s.C.tolist()}, range={s.maxrange})" for s in self.structs]
        return " + ".join(parts)


def _basic(kind: int, variogram: bool, range=1.0, sill=1.0, nugget=0.0, ranges=None, rotation=None, order=None) -> GeoStatsFunction:
    A = metric_matrix(range, ranges, rotation)
    mr = float(range) if ranges is None else float(max(ranges))
    if kind == MATERN:
        order = 1.0 if order is None else float(order)  # GeoStatsFunctions' default
        if not order > 0.0:
            raise ValueError("Matern order must be positive")
    elif order is not None:
        raise TypeError("`order` is a Matern parameter")
    structs = [_Struct(kind, np.array([[float(sill) - float(nugget)]]), A, mr, order or 0.0)]
    if nugget != 0.0:
        structs.append(_Struct(NUGGET, np.array([[float(nugget)]]), np.eye(3), 0.0))
    return GeoStatsFunction(structs, variogram)


def SphericalCovariance(**kw): return _basic(SPHERICAL, False, **kw)
def ExponentialCovariance(**kw): return _basic(EXPONENTIAL, False, **kw)
def GaussianCovariance(**kw): return _basic(GAUSSIAN, False, **kw)
def CubicCovariance(**kw): return _basic(CUBIC, False, **kw)
def PentasphericalCovariance(**kw): return _basic(PENTASPHERICAL, False, **kw)
def SineHoleCovariance(**kw): return _basic(SINEHOLE, False, **kw)
def CircularCovariance(**kw): return _basic(CIRCULAR, False, **kw)
def MaternCovariance(**kw): return _basic(MATERN, False, **kw)
def MaternVariogram(**kw): return _basic(MATERN, True, **kw)
def SineHoleVariogram(**kw): return _basic(SINEHOLE, True, **kw)
def CircularVariogram(**kw): return _basic(CIRCULAR, True, **kw)
def SphericalVariogram(**kw): return _basic(SPHERICAL, True, **kw)
def ExponentialVariogram(**kw): return _basic(EXPONENTIAL, True, **kw)
def GaussianVariogram(**kw): return _basic(GAUSSIAN, True, **kw)
def CubicVariogram(**kw): return _basic(CUBIC, True, **kw)
def PentasphericalVariogram(**kw): return _basic(PENTASPHERICAL, True, **kw)


def NuggetEffect(nugget: float = 1.0, variogram: bool = False) -> GeoStatsFunction:
    return GeoStatsFunction([_Struct(NUGGET, np.array([[float(nugget)]]), np.eye(3), 0.0)], variogram)
