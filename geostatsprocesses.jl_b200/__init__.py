"""B200-native Gaussian-field simulation engine (LUSIM / FFTSIM hot path of GeoStatsProcesses.jl).

The directory name contains a dot, so import it through the `gsp_b200` shim at the repo root:

    import gsp_b200 as gsp
    ens = gsp.rand(gsp.GaussianProcess(gsp.SphericalCovariance(range=20.0)), gsp.CartesianGrid(50, 50), 100,
                   method=gsp.LUSIM())
"""
from ._lib import DEFAULT_LIB, DeviceEnsemble, FFTPlan, GspError, Library, LUPlan, PosDefException, SIGNATURES  # noqa: F401
from .domains import CartesianGrid, GeoTable, GridView, PointSet, georef  # noqa: F401
from .functions import (MaternCovariance, MaternVariogram, CircularCovariance, CircularVariogram, SineHoleCovariance, SineHoleVariogram, CubicCovariance, CubicVariogram, ExponentialCovariance, ExponentialVariogram,  # noqa: F401
                        GaussianCovariance, GaussianVariogram, GeoStatsFunction, NuggetEffect, PentasphericalCovariance,
                        PentasphericalVariogram, SphericalCovariance, SphericalVariogram, metric_matrix)
from .processes import (FFTSIM, LUSIM, Ensemble, ExplicitInit, FieldSimulationMethod, GaussianProcess, NearestInit,  # noqa: F401
                        default_library, defaultsimulation, expectedvalue, initialize, mean, merge_moments, preprocess_fftsim, preprocess_lusim, rand,
                        rand_fftsim, rand_lusim, set_devices)

__version__ = "0.1.0"
