"""Mirrors of the reference's own tests for this path, against the reference-facing Python layer
(test/field.jl:3-71,115-132; test/initialization.jl:5-59; test/ensembles.jl:1-60) on the emulated build."""
import numpy as np
import pytest

import gsp_b200 as gsp


def test_defaultsimulation():  # test/field.jl:3-15
    proc1 = gsp.GaussianProcess(gsp.GaussianVariogram())
    proc2 = gsp.GaussianProcess(gsp.GaussianCovariance())
    grid = gsp.CartesianGrid(100, 100)
    vgrid = grid.view(range(1, 1001))
    pset1 = gsp.PointSet(np.random.default_rng(0).random((1000, 2)))
    assert isinstance(gsp.defaultsimulation(proc1, grid), gsp.FFTSIM)
    assert isinstance(gsp.defaultsimulation(proc1, vgrid), gsp.FFTSIM)
    with pytest.raises(NotImplementedError):  # the reference picks SEQSIM (out of scope here)
        gsp.defaultsimulation(proc1, pset1)
    assert isinstance(gsp.defaultsimulation(proc2, pset1), gsp.LUSIM)


def test_lusim_api(emu_lib):  # test/field.jl:17-71 (smaller grids: the emulator is slow)
    method = gsp.LUSIM(library=emu_lib)
    rng = np.random.default_rng(123)
    # 1-D, unconditional and conditional
    proc = gsp.GaussianProcess(gsp.SphericalCovariance(range=10.0))
    grid = gsp.CartesianGrid(100)
    data = gsp.georef({"Z": [0.0, 1.0, 0.0, 1.0, 0.0]}, [(0.0,), (25.0,), (50.0,), (75.0,), (100.0,)])
    real = gsp.rand(proc, grid, rng=rng, method=method)
    assert real.field.dtype == np.float64 and real.nrow == 100
    real = gsp.rand(proc, grid, rng=rng, method=method, data=data)
    assert real.Z.dtype == np.float64
    assert real.Z[0] == 0.0 and real.Z[25] == 1.0 and real.Z[50] == 0.0 and real.Z[75] == 1.0 and real.Z[99] == 0.0
    # cosimulation
    func = [[1.0, 0.95], [0.95, 1.0]] * gsp.SphericalCovariance(range=10.0)
    proc = gsp.GaussianProcess(func, [0.0, 0.0])
    real = gsp.rand(proc, grid, rng=rng, method=method)
    assert real.field1.dtype == np.float64 and real.field2.dtype == np.float64
    assert np.corrcoef(real.field1, real.field2)[0, 1] > 0.5
    # 2-D
    proc = gsp.GaussianProcess(gsp.SphericalCovariance(range=10.0))
    real = gsp.rand(proc, gsp.CartesianGrid(12, 12), rng=rng, method=method)
    assert real.field.shape == (144,)
    # bivariate with data (named columns)
    func = [[1.0, 0.8], [0.8, 1.0]] * gsp.SphericalCovariance(range=35.0)
    proc = gsp.GaussianProcess(func, [0.1, 0.2])
    grid = gsp.CartesianGrid(10, 10)
    data = gsp.georef({"Cu": [0.0, 0.1, 0.0], "Zn": [0.1, 0.0, 0.1]}, [(2.5, 2.5), (5.0, 7.5), (7.5, 5.0)])
    ens = gsp.rand(proc, grid, 3, rng=rng, method=method, data=data)
    assert len(ens) == 3 and ens.variables() == ("Cu", "Zn")
    j = grid.nearest(np.array([2.5, 2.5]))
    assert all(r.Cu[j] == 0.0 and r.Zn[j] == 0.1 for r in ens)


def test_matern_through_rand(emu_lib):
    """MaternCovariance / MaternVariogram (order = nu) cross the boundary with their parameter: nu = 1/2 reproduces the
    exponential model of the same range draw for draw, through LUSIM and FFTSIM."""
    grid = gsp.CartesianGrid(12, 9)
    data = gsp.georef({"Z": [1.0, -0.5, 0.3]}, [(2.5, 2.5), (9.5, 4.5), (5.5, 7.5)])
    lu = gsp.LUSIM(library=emu_lib)
    a = gsp.rand(gsp.GaussianProcess(gsp.MaternCovariance(range=6.0, sill=1.4, order=0.5)), grid, 3, rng=np.random.default_rng(4), method=lu, data=data)
    b = gsp.rand(gsp.GaussianProcess(gsp.ExponentialCovariance(range=6.0, sill=1.4)), grid, 3, rng=np.random.default_rng(4), method=lu, data=data)
    for ra, rb in zip(a, b):
        assert np.abs(ra.Z - rb.Z).max() < 1e-10
    ff = gsp.FFTSIM(library=emu_lib)
    a = gsp.rand(gsp.GaussianProcess(gsp.MaternVariogram(range=6.0, order=0.5)), grid, 2, rng=np.random.default_rng(5), method=ff)
    b = gsp.rand(gsp.GaussianProcess(gsp.ExponentialVariogram(range=6.0)), grid, 2, rng=np.random.default_rng(5), method=ff)
    for ra, rb in zip(a, b):
        assert np.abs(ra.field - rb.field).max() < 1e-10
    assert gsp.MaternCovariance().structs[0].param == 1.0          # GeoStatsFunctions' default order
    with pytest.raises(ValueError):
        gsp.MaternCovariance(order=0.0)
    with pytest.raises(TypeError):
        gsp.SphericalCovariance(order=1.0)
    with pytest.raises(ValueError, match="Matern order"):            # the C ABI rejects a non-positive order itself
        emu_lib.pairwise([(8, 1.0, np.eye(3), -1.0)], np.zeros((2, 2)))


def test_lusim_rejects_variograms(emu_lib):  # lusim.jl:44-50
    proc = gsp.GaussianProcess(gsp.SphericalVariogram(range=10.0))
    with pytest.raises(ValueError, match="stationary, symmetric and banded"):
        gsp.rand(proc, gsp.CartesianGrid(10), method=gsp.LUSIM(library=emu_lib))


def test_gaussianprocess_mean_arity():  # gaussian.jl:26-31
    with pytest.raises(AssertionError):
        gsp.GaussianProcess(gsp.SphericalCovariance(), [0.0, 0.0])


def test_fftsim_api(emu_lib):  # test/field.jl:115-132
    method = gsp.FFTSIM(library=emu_lib)
    rng = np.random.default_rng(2019)
    proc = gsp.GaussianProcess(gsp.GaussianVariogram(range=3.0))
    grid = gsp.CartesianGrid(20, 20)
    real = gsp.rand(proc, grid, rng=rng, method=method)
    assert real.field.dtype == np.float64 and real.nrow == 400
    vgrid = grid.view(range(1, 201))
    real = gsp.rand(proc, vgrid, rng=rng, method=method)
    assert real.domain == vgrid and real.nrow == 200
    with pytest.raises(AssertionError):  # fftsim.jl:69
        gsp.rand(gsp.GaussianProcess([[1.0, 0.5], [0.5, 1.0]] * gsp.SphericalCovariance(range=3.0), [0.0, 0.0]), grid, method=method)
    ens = gsp.rand(proc, grid, 4, rng=7, method=method)  # device RNG
    assert len(ens) == 4 and abs(ens[0].field.mean()) < 1e-12


def test_initialization():  # test/initialization.jl:5-59
    proc = gsp.GaussianProcess(gsp.GaussianVariogram())
    grid = gsp.CartesianGrid((-0.5, -0.5), (99.5, 99.5), dims=(100, 100))
    data2D = gsp.georef({"value": [1.0, 2.0, 3.0]}, [(25.0, 25.0), (50.0, 75.0), (75.0, 50.0)])
    real, mask = gsp.initialize(proc, grid, data2D, gsp.NearestInit())
    for i, j in ((5076, 3), (2526, 1), (7551, 2)):
        assert real["value"][i - 1] == data2D.value[j - 1] and mask["value"][i - 1]
    grid3 = gsp.CartesianGrid(10, 10, 10)
    vals = np.random.default_rng(0).random(10)
    data = gsp.georef({"z": vals}, np.random.default_rng(1).random((10, 3)))
    real, mask = gsp.initialize(proc, grid3, data, gsp.ExplicitInit(range(1, 11)))
    assert np.array_equal(real["z"][:10], vals) and mask["z"][:10].all()
    real, mask = gsp.initialize(proc, grid3, data, gsp.ExplicitInit(range(991, 1001)))
    assert np.array_equal(real["z"][990:], vals)
    real, mask = gsp.initialize(proc, grid3, data, gsp.ExplicitInit(range(1, 4), range(998, 1001)))
    assert np.array_equal(real["z"][997:], vals[:3]) and mask["z"].sum() == 3


def test_ensemble():  # test/ensembles.jl
    grid = gsp.CartesianGrid(3, 3)
    ens = gsp.Ensemble(grid, {"z": np.tile(np.arange(1.0, 4.0), (9, 1))})
    assert len(ens) == 3 and ens[1].z.tolist() == [2.0] * 9
    assert ens.mean().z.tolist() == [2.0] * 9
    assert ens.var().z.tolist() == [1.0] * 9
    assert np.allclose(ens.cdf(2.0).z, 2 / 3) and np.allclose(ens.ccdf(2.0).z, 1 / 3)
    assert ens.quantile(0.5).z.tolist() == [2.0] * 9
    assert "2D Ensemble" in repr(ens) and "N° reals:  3" in repr(ens)
    assert [r.z[0] for r in ens] == [1.0, 2.0, 3.0]


def test_unseeded_rand_calls_are_independent(emu_lib):
    """rng=None draws a fresh seed per call (the reference uses Random.default_rng(), field.jl:47-48); an int seed reproduces."""
    proc = gsp.GaussianProcess(gsp.SphericalCovariance(range=5.0))
    grid = gsp.CartesianGrid(12, 10)
    for method in (gsp.LUSIM(library=emu_lib), gsp.FFTSIM(library=emu_lib)):
        a = gsp.rand(proc, grid, 2, method=method)
        b = gsp.rand(proc, grid, 2, method=method)
        assert not np.array_equal(a[0].field, b[0].field)
        c = gsp.rand(proc, grid, 2, rng=11, method=method)
        d = gsp.rand(proc, grid, 2, rng=11, method=method)
        assert np.array_equal(c[0].field, d[0].field) and np.array_equal(c[1].field, d[1].field)


def test_expectation(emu_lib):  # test/field.jl:156-176 ("Expectation") on a 40 x 40 grid (the emulator is slow); posterior mean = simple Kriging
    import gsp_oracle as O
    from helpers import ostructs, iso

    proc = gsp.GaussianProcess(gsp.SphericalVariogram(range=14.0))
    grid = gsp.CartesianGrid((0.5, 0.5), (40.5, 40.5), dims=(40, 40))
    mval = gsp.mean(proc, grid)
    assert np.all(mval.field == 0.0) and mval.nrow == 1600
    assert np.all(gsp.mean(gsp.GaussianProcess(gsp.SphericalCovariance(range=5.0), 2.5), grid).field == 2.5)
    pts = [(10.0, 10.0), (20.0, 30.0), (30.0, 20.0)]
    data = gsp.georef({"Z": [1.0, 0.0, 1.0]}, pts)
    mval = gsp.mean(proc, grid, data=data, library=emu_lib)
    Z = mval.Z.reshape(40, 40)                # [y][x]: element (i, j), 1-based, is Z[j-1, i-1]; the data sit on centroids
    assert abs(Z[9, 9] - 1.0) < 1e-12 and abs(Z[29, 19] - 0.0) < 1e-12 and abs(Z[19, 29] - 1.0) < 1e-12
    # against the oracle's Kriging (expectation/field/gaussian.jl:21-25; GeoStatsModels.fitpredict defaults: 10 nearest -> all 3 data)
    st = ostructs(iso(O.SPHERICAL, 1.0, 14.0, 2))
    cent = O.grid_centroids((40, 40), [0.5, 0.5], [1.0, 1.0])
    nbr, lam = O.krige_neighbors_weights(st, cent, np.array(pts), 10)
    zo = (lam * np.array([1.0, 0.0, 1.0])[nbr]).sum(axis=1)
    assert np.abs(mval.Z - zo).max() < 1e-10
    ev = gsp.expectedvalue(proc, grid, data=data, library=emu_lib)
    assert np.array_equal(ev.Z, mval.Z) and np.array_equal(gsp.expectedvalue(proc, grid).field, gsp.mean(proc, grid).field)
    # a view of the grid: only its elements are predicted
    vgrid = grid.view(range(1, 801))
    mv = gsp.mean(proc, vgrid, data=data, library=emu_lib)
    assert mv.nrow == 800 and np.abs(mv.Z - zo[:800]).max() < 1e-10
