"""Writes tests/golden/golden_small.npz: seeded inputs and ORACLE outputs (the reference itself cannot
run here - no Julia - so these vectors pin regressions of the restatement, not the reference)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import gsp_oracle as O  # noqa: E402
from helpers import iso, ostructs  # noqa: E402

rng = np.random.default_rng(20261017)
st = ostructs(iso(O.SPHERICAL, 1.0, 20.0, 2))
coords = O.grid_centroids((12, 10), (0, 0), (1, 1))
dinds = np.sort(rng.choice(120, 9, replace=False))
z1 = rng.standard_normal(9)
pre = O.lusim_preprocess(st, coords, dinds, z1, 0.0)
W = rng.standard_normal((111, 4))
Z = O.lusim_sample(pre, W)
st = ostructs(iso(O.EXPONENTIAL, 1.0, 5.0, 3))
F = O.fftsim_preprocess(st, (8, 6, 4), [0, 0, 0], [1, 1, 1])
w = rng.random(8 * 6 * 4)
Zf = O.fftsim_sample(F, w, 1.0, 0.5)
np.savez(os.path.join(HERE, "golden_small.npz"), lu_dinds=dinds, lu_z1=z1, lu_W=W, lu_Z=Z, fft_w=w, fft_Z=Zf, fft_F=F)
print("written")
