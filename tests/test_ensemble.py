"""Device-resident ensembles and their statistics (src/ensembles.jl:42-52) - SURVEY §8f rank 1.

Pins: the reference's own value-level test (test/ensembles.jl:24-59: constant realizations 1, 2, 3 on a 3x3 grid) is
re-stated here against the oracle AND the kernels (emulated build; the GPU versions are in test_gpu_parity.py)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import gsp_b200 as gsp
import gsp_oracle as O
from helpers import iso, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_pins(mean, var, cdf, ccdf, quant):
    """test/ensembles.jl:24-59"""
    ones = np.ones(9)
    assert np.array_equal(mean(), 2.0 * ones)
    assert np.array_equal(var(), 1.0 * ones)
    assert np.array_equal(cdf(1), 1 / 3 * ones)
    assert np.array_equal(cdf(2), 2 / 3 * ones)
    assert np.array_equal(cdf(3), 3 / 3 * ones)
    for i in (1, 2, 3):
        assert np.allclose(ccdf(i), 1 - cdf(i), rtol=1e-15, atol=1e-16)
    q = quant([0.0, 0.5, 1.0])
    assert np.array_equal(q[1], 2.0 * ones) and np.array_equal(q[0], ones) and np.array_equal(q[2], 3.0 * ones)


def test_oracle_matches_reference_ensemble_test():
    Z = np.stack([i * np.ones(9) for i in (1.0, 2.0, 3.0)])
    st = lambda x=0.0, ps=(0.5,): O.ensemble_stats(Z, x, ps)
    _reference_pins(lambda: st()[0], lambda: st()[1], lambda x: st(x)[2], lambda x: st(x)[3], lambda ps: st(0.0, ps)[4])


def test_oracle_quantile_known_answers():
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10, 101):
        v = rng.standard_normal(n)
        for p in (0.0, 0.25, 0.5, 0.9, 1.0, 1 / 3):
            assert abs(O.julia_quantile(v, p) - np.quantile(v, p)) <= 1e-15 * max(1.0, np.abs(v).max())
    assert O.julia_quantile(np.array([1.0, 2.0, 3.0, 4.0]), 0.5) == 2.5  # Julia docs: quantile(1:4, 0.5)
    with pytest.raises(ValueError):
        O.julia_quantile(np.ones(3), 1.5)


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
def test_device_ensemble_reference_pins(emu_lib, devices):
    lib = gsp.Library(emu_lib.path, devices=devices)
    ens = gsp.DeviceEnsemble(lib, 9, 3)
    ens.put(np.stack([i * np.ones(9) for i in (1.0, 2.0, 3.0)]))
    _reference_pins(ens.mean, ens.var, ens.cdf, ens.ccdf, ens.quantile)
    assert np.array_equal(ens.fetch(1, 1)[0], 2.0 * np.ones(9))
    ens.close()
    lib.close()


@pytest.mark.parametrize("devices,n,R", [([0], 77, 1), ([0], 77, 2), ([0], 130, 5), ([0], 77, 64), ([0, 0], 77, 100), ([0, 0, 0], 40, 7),
                                         ([0], 19, 1000), ([0, 0], 3, 5000)])
def test_device_ensemble_statistics_vs_oracle(emu_lib, devices, n, R):
    rng = np.random.default_rng(n * 1000 + R)
    Z = rng.standard_normal((R, n)) * 1.7 + 40.0  # a mean far from zero: the shifted sums must not cancel
    Z[:, 0] = 3.25                                 # a constant node
    lib = gsp.Library(emu_lib.path, devices=devices)
    ens = gsp.DeviceEnsemble(lib, n, R)
    ens.put(Z[: R // 2], 0)
    ens.put(Z[R // 2:], R // 2)
    assert np.array_equal(ens.fetch(), Z)
    ps = [0.0, 0.1, 0.5, 0.93, 1.0]
    mean, var, cdf, ccdf, q = O.ensemble_stats(Z, 40.3, ps)
    assert relerr(ens.mean(), mean) < 1e-14
    if R > 1:
        assert np.abs(ens.var() - var).max() < 1e-12 * var.max()
    else:
        assert np.all(np.isnan(ens.var()))
    assert np.array_equal(ens.cdf(40.3), cdf) and np.array_equal(ens.ccdf(40.3), ccdf)
    assert relerr(ens.quantile(ps), q) < 1e-15
    m, m2 = ens.moments()
    assert relerr(m, mean) < 1e-14
    with pytest.raises(ValueError):  # Statistics.quantile: ArgumentError
        ens.quantile([1.2])
    ens.close()
    lib.close()


def test_fill_by_simulation_matches_host_path(emu_lib):
    lib2 = gsp.Library(emu_lib.path, devices=[0, 0])
    for lib in (emu_lib, lib2):
        st = iso(O.SPHERICAL, 1.3, 3.0, 2)
        plan = gsp.FFTPlan(lib, st, (12, 8), [0.0, 0.0], [1.0, 1.0])
        w = np.random.default_rng(1).random((5, 96))
        inds1 = np.array([1, 96, 17, 40])
        for kw in ({}, {"inds1": inds1}):
            e = plan.sample_ensemble(5, w, sill=1.3, mu=0.2, **kw)
            assert np.array_equal(e.fetch(), plan.sample(5, w, sill=1.3, mu=0.2, **kw))
            e.close()
            e = plan.sample_ensemble(5, None, seed=4, sill=1.3, mu=0.2, **kw)
            assert np.array_equal(e.fetch(), plan.sample(5, None, seed=4, sill=1.3, mu=0.2, **kw))
            e.close()
        plan.close()
        dom = (gsp._lib.make_grid_domain((8, 6), (0, 0), (1, 1)), None)
        q = gsp.LUPlan(lib, st, dom, np.array([3, 9]), np.array([0.5, -0.5]), 0.0)
        W = np.random.default_rng(3).standard_normal((46, 5))
        W2 = np.random.default_rng(4).standard_normal((46, 5))
        e = q.sample_ensemble(5, W)
        assert np.array_equal(e.fetch().T, q.sample(5, W))
        e.close()
        e = q.sample_ensemble(5, W2, rho=0.6, W1=W)
        assert np.array_equal(e.fetch().T, q.sample(5, W2, rho=0.6, W1=W))
        e.close()
        e = q.sample_ensemble(5, None, seed=8, stream=1, rho=0.6)
        assert np.array_equal(e.fetch().T, q.sample(5, None, seed=8, stream=1, rho=0.6))
        assert np.array_equal(e.mean()[[2, 8]], [0.5, -0.5])  # conditioning data honoured exactly in every realization
        assert np.array_equal(e.var()[[2, 8]], [0.0, 0.0])
        e.close()
        q.close()
    lib2.close()


def test_rand_resident_ensemble(emu_lib):
    """rand(..., resident=True): Ensemble with the fetch hook (ensembles.jl:16,27-31); statistics equal the host ensemble's."""
    grid = gsp.CartesianGrid(16, 8)
    proc = gsp.GaussianProcess(gsp.SphericalCovariance(range=3.0, sill=1.5), 0.4)
    host = gsp.rand(proc, grid, 6, rng=3, method=gsp.FFTSIM(library=emu_lib))
    dev = gsp.rand(proc, grid, 6, rng=3, method=gsp.FFTSIM(library=emu_lib), resident=True)
    assert len(dev) == 6 and dev.variables() == ("field",)
    assert np.array_equal(dev[2].field, host[2].field)
    assert relerr(dev.mean().field, host.mean().field) < 1e-14
    assert relerr(dev.var().field, host.var().field) < 1e-13
    assert np.array_equal(dev.cdf(0.4).field, host.cdf(0.4).field)
    assert relerr(dev.quantile(0.3).field, host.quantile(0.3).field) < 1e-14
    assert [relerr(a.field, b.field) < 1e-14 for a, b in zip(dev.quantile([0.1, 0.9]), host.quantile([0.1, 0.9]))] == [True, True]
    dev.close()
    # bivariate LUSIM with conditioning data (named columns, like test/field.jl:60-71)
    func = [[1.0, 0.7], [0.7, 1.0]] * gsp.SphericalCovariance(range=4.0)
    proc2 = gsp.GaussianProcess(func, [0.1, 0.2])
    g2 = gsp.CartesianGrid(8, 8)
    data = gsp.georef({"Cu": [0.0, 0.1], "Zn": [0.1, 0.0]}, [(2.5, 2.5), (5.5, 6.5)])
    host = gsp.rand(proc2, g2, 4, rng=np.random.default_rng(5), method=gsp.LUSIM(library=emu_lib), data=data)
    dev = gsp.rand(proc2, g2, 4, rng=np.random.default_rng(5), method=gsp.LUSIM(library=emu_lib), data=data, resident=True)
    assert dev.variables() == ("Cu", "Zn")
    for v in host.variables():
        assert np.array_equal(dev[3][v], host[3][v])
        assert relerr(dev.mean()[v], host.mean()[v]) < 1e-13
    j = g2.nearest(np.array([2.5, 2.5]))
    assert dev.mean()["Cu"][j] == 0.0 and dev.var()["Zn"][j] == 0.0
    dev.close()


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'oracle')); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
    import gsp_b200 as gsp, gsp_oracle as O
    from helpers import iso
    from bench import shard_range
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = gsp.Library(os.path.join(%(root)r, 'tests', 'emu', 'libgspb200_emu.so'))
    plan = gsp.FFTPlan(lib, iso(O.SPHERICAL, 1.0, 3.0, 2), (12, 10), [0.0, 0.0], [1.0, 1.0])
    R = 9
    r0, r1 = shard_range(R, rank, world)
    ens = plan.sample_ensemble(r1 - r0, None, seed=11, first_real=r0, mu=5.0)   # this rank's shard stays on its device
    mean, m2 = ens.moments()
    parts = [None] * world
    dist.all_gather_object(parts, (r1 - r0, mean, m2))                          # one exchange of 2 n doubles per rank
    cnt, gmean, gvar = gsp.merge_moments(parts)
    if rank == 0:
        full = plan.sample(R, None, seed=11, first_real=0, mu=5.0)
        assert cnt == R
        assert np.abs(gmean - full.mean(axis=0)).max() < 1e-13
        assert np.abs(gvar - full.var(axis=0, ddof=1)).max() < 1e-13
        print('OK')
    dist.destroy_process_group()
""")


def test_two_rank_ensemble_moments_gloo(emu_lib, tmp_path):
    """N>1 (one process per GPU): per-rank partial moments + one all-gather reproduce the single-rank mean / variance."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29547", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK" in out.stdout
