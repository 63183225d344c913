import numpy as np

import gsp_oracle as O


def relerr(a, b):
    """normwise max|a-b| / max|b| (SURVEY §8c acceptance metric)."""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max())


def iso(kind, sill, rang, ndim=3, order=None):
    A = np.zeros((3, 3))
    for a in range(ndim):
        A[a, a] = 1.0 / rang
    return [(kind, float(sill), A) if order is None else (kind, float(sill), A, float(order))]


def ostructs(structs):
    return [O.Structure(*st) for st in structs]


def aniso3(kind, sill, ranges, angle_deg, order=None):
    th = np.radians(angle_deg)
    R = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
    A = np.diag(1.0 / np.asarray(ranges, dtype=float)) @ R.T
    return [(kind, float(sill), A) if order is None else (kind, float(sill), A, float(order))]
