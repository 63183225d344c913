"""Oracle (and, under -m gpu, the CUDA path) against vectors produced by the REAL GeoStatsProcesses.jl v0.13.0.

The fixtures are written by tests/golden/make_golden.jl (needs Julia, absent from this image) into tests/golden/julia/.
Until they exist every test here is xfail(strict): the suite stays green, and the day the fixtures are committed a still-xfail
mark turns into a failure, i.e. the marks must come off and the assertions hold: oracle == reference at 1e-12 relative
(normwise), which pins what DESIGN.md lists as "our reading" (model formulas, index conventions, lusim.jl block algebra,
fftsim.jl scaling, GeoStatsModels' neighbour ties and the support of the second Kriging)."""
import os
import tomllib

import numpy as np
import pytest

import gsp_oracle as O
from helpers import aniso3, iso, ostructs, relerr

JDIR = os.path.join(os.path.dirname(__file__), "golden", "julia")
HAVE = os.path.exists(os.path.join(JDIR, "manifest.toml"))
needs_fixtures = pytest.mark.xfail(condition=not HAVE, strict=True, raises=FileNotFoundError,
                                   reason="reference fixtures absent: run tests/golden/make_golden.jl with Julia (none in this image)")
KINDS = {"spherical": O.SPHERICAL, "exponential": O.EXPONENTIAL, "gaussian": O.GAUSSIAN, "cubic": O.CUBIC,
         "pentaspherical": O.PENTASPHERICAL, "sinehole": O.SINEHOLE, "circular": O.CIRCULAR, "matern": O.MATERN}
TOL = 1e-12


class Fixtures:
    def __init__(self):
        with open(os.path.join(JDIR, "manifest.toml"), "rb") as f:  # FileNotFoundError while the fixtures are absent
            self.m = tomllib.load(f)["cases"]

    def arr(self, case, name):
        meta = self.m[case][name]
        dt = {"Float64": "<f8", "Int64": "<i8"}[meta["dtype"]]
        a = np.fromfile(os.path.join(JDIR, f"{case}.{name}.bin"), dtype=dt)
        return a.reshape(meta["shape"], order="F")


def with_nugget(st, nug):
    return st + ([(O.NUGGET, float(nug), np.eye(3))] if nug else [])


@needs_fixtures
def test_pairwise_models_against_reference():
    fx = Fixtures()
    cases = [c for c in fx.m if c.startswith("pairwise_")]
    assert len(cases) >= 10
    for c in cases:
        md = fx.m[c]
        st = with_nugget(aniso3(KINDS[md["kind"]], md["sill"], md["ranges"], md["angle_deg"], order=md.get("order")), md["nugget"])
        X = fx.arr(c, "X").T  # dim x n column-major -> (n, dim)
        assert relerr(O.pairwise(ostructs(st), X), fx.arr(c, "C")) < TOL, c


def _lusim_oracle(fx, case, structs_per_var, means, rho):
    md = fx.m[case]
    dims = md["dims"]
    coords = O.grid_centroids(dims, [0.0] * len(dims), [1.0] * len(dims))
    out = []
    W1 = None
    for j, var in enumerate(md["vars"]):
        if "dcoords" in md:
            dinds, z1 = O.nearest_init(dims, [0.0] * len(dims), [1.0] * len(dims), fx.arr(case, "dcoords").T, fx.arr(case, f"dvals_{var}"))
        else:
            dinds, z1 = np.zeros(0, dtype=np.int64), np.zeros(0)
        pre = O.lusim_preprocess(structs_per_var[j], coords, dinds, z1, means[j])
        W = fx.arr(case, f"W{j + 1}")
        Z = O.lusim_sample(pre, W) if j == 0 else O.lusim_sample(pre, W, rho, W1)
        W1 = W
        out.append((Z[:, 0], fx.arr(case, f"Z{j + 1}")))
    return out


@needs_fixtures
def test_lusim_against_reference():
    fx = Fixtures()
    mu = fx.m["lusim_uni_meta"]
    st = ostructs(with_nugget(iso(KINDS[mu["kind"]], mu["sill"], mu["range"], 2), mu["nugget"]))
    for case in ("lusim_uni", "lusim_cond"):
        for z, zref in _lusim_oracle(fx, case, [st], [mu["mean"]], None):
            assert relerr(z, zref) < TOL, case
    mb = fx.m["lusim_bi_meta"]
    C = np.asarray(mb["C"]).reshape(2, 2)
    mv = [(KINDS[mb["kind"]], C, np.diag([1 / mb["range"], 1 / mb["range"], 0.0]))]
    for z, zref in _lusim_oracle(fx, "lusim_bi", [O.marginalize(mv, 0), O.marginalize(mv, 1)], mb["mean"], O.rho_mv(mv)):
        assert relerr(z, zref) < TOL


@needs_fixtures
def test_fftsim_against_reference():
    fx = Fixtures()
    mf = fx.m["fftsim_meta"]
    for case in ("fftsim_2d", "fftsim_3d", "fftsim_view", "fftsim_cond_k3", "fftsim_cond_k26"):
        md = fx.m[case]
        dims = md["dims"]
        nd = len(dims)
        st = ostructs(iso(KINDS[mf["kind"]], mf["sill"], mf["range"], nd))
        F = O.fftsim_preprocess(st, dims, [0.0] * nd, [1.0] * nd)
        w = fx.arr(case, "w").reshape(-1, order="F")
        inds0 = fx.arr(case, "inds") - 1 if "inds" in md else None
        if "dcoords" in md:
            dco, dv = fx.arr(case, "dcoords").T, fx.arr(case, "dvals")
            knodes0, _ = O.nearest_init(dims, [0.0] * nd, [1.0] * nd, dco, dv)
            cond = O.fftsim_condition(st, dims, [0.0] * nd, [1.0] * nd, dco, dv, knodes0, mf["mean"], md["maxneighbors"])
            z = O.fftsim_sample_conditional(F, w, mf["sill"], cond)
        else:
            z = O.fftsim_sample(F, w, mf["sill"], mf["mean"], inds0)
        assert relerr(z, fx.arr(case, "Z")) < TOL, case


@pytest.mark.gpu
@needs_fixtures
def test_cuda_path_against_reference(gpu_lib):
    """the product itself against the reference's vectors (1e-9, north_star), same cases"""
    import gsp_b200 as gsp

    fx = Fixtures()
    mf = fx.m["fftsim_meta"]
    for case in ("fftsim_2d", "fftsim_3d", "fftsim_view"):
        md = fx.m[case]
        dims, nd = md["dims"], len(md["dims"])
        plan = gsp.FFTPlan(gpu_lib, iso(KINDS[mf["kind"]], mf["sill"], mf["range"], nd), dims, [0.0] * nd, [1.0] * nd)
        inds1 = fx.arr(case, "inds") if "inds" in md else None
        z = plan.sample(1, fx.arr(case, "w").reshape(1, -1, order="F"), sill=mf["sill"], mu=mf["mean"], inds1=inds1)[0]
        assert relerr(z, fx.arr(case, "Z")) < 1e-9, case
        plan.close()
    mu = fx.m["lusim_uni_meta"]
    st = with_nugget(iso(KINDS[mu["kind"]], mu["sill"], mu["range"], 2), mu["nugget"])
    for case in ("lusim_uni", "lusim_cond"):
        md = fx.m[case]
        dims = md["dims"]
        if "dcoords" in md:
            dinds, z1 = O.nearest_init(dims, [0.0, 0.0], [1.0, 1.0], fx.arr(case, "dcoords").T, fx.arr(case, "dvals_Z"))
        else:
            dinds, z1 = None, None
        plan = gsp.LUPlan(gpu_lib, st, (gsp._lib.make_grid_domain(dims, [0.0, 0.0], [1.0, 1.0]), None),
                          None if dinds is None else dinds + 1, z1, mu["mean"])
        z = plan.sample(1, fx.arr(case, "W1").reshape(-1, 1))[:, 0]
        assert relerr(z, fx.arr(case, "Z1")) < 1e-9, case
        plan.close()
