"""FFTSIM 256^3 schedule sweep (development tool): lanes x slab mode x slab size, device-resident noise, checks every
variant against the default schedule (bit-identical fields expected: same kernels, same arithmetic)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import aniso3
import torch
lib = gsp.Library()
dev = torch.device("cuda:0")
dims = (256, 256, 256); N = 256 ** 3; R = int(os.environ.get("SWEEP_R", "32"))
st = aniso3(O.SPHERICAL, 1.0, (40.0, 20.0, 10.0), 30.0)
w = torch.rand((R, N), dtype=torch.float64, device=dev)
z = torch.empty((R, N), dtype=torch.float64, device=dev)
zref = None
configs = [(0, 1, 0, 0)]
for lanes in (2, 4):
    configs.append((0, lanes, 0, 0))
for lanes in (1, 2, 3, 4):
    for planes in (8, 16, 32, 64):
        configs.append((1, lanes, planes, 0))
for lanes in (1, 2, 3):
    for bundles in (1, 2, 3, 5, 9):
        configs.append((2, lanes, 0, bundles))
sel = os.environ.get("SWEEP_ONLY")
for mode, lanes, planes, bundles in configs:
    if sel and str(mode) not in sel.split(","):
        continue
    os.environ["GSP_FFT_SLAB"] = str(mode); os.environ["GSP_FFT_LANES"] = str(lanes)
    os.environ["GSP_FFT_SLAB_PLANES"] = str(planes or 32); os.environ["GSP_FFT_SLAB_BUNDLES"] = str(bundles or 4)
    plan = gsp.FFTPlan(lib, st, dims, [0.0] * 3, [1.0] * 3)
    best = 1e9; wall = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        plan.sample_dev(R, w.data_ptr(), 0, 0, 1.0, 0.0, 0, None, z.data_ptr())
        wall = min(wall, (time.perf_counter() - t0) * 1e3)
        best = min(best, lib.last_sample_ms())
    if zref is None:
        zref = z.clone()
        same = True
    else:
        same = bool(torch.equal(z, zref))
    print(f"mode {mode} lanes {lanes} planes {planes} bundles {bundles}: {best / R * 1e3:7.1f} us/real {R / best * 1e3:7.0f} real/s wall {wall / R * 1e3:7.1f} us/real identical={same}", flush=True)
    plan.close()
