"""FFTSIM 256^3: separate passes vs fused x+y plane kernels (GSP_FFT_FUSE), lanes 1/2/4, injected noise and device RNG.
Development tool: every variant must reproduce the default schedule bit for bit."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import aniso3
import torch
lib = gsp.Library()
dev = torch.device("cuda:0")
n = int(os.environ.get("FUSE_N", "256"))
dims = (n, n, n); N = n ** 3; R = int(os.environ.get("SWEEP_R", "32"))
st = aniso3(O.SPHERICAL, 1.0, (40.0, 20.0, 10.0), 30.0)
w = torch.rand((R, N), dtype=torch.float64, device=dev)
z = torch.empty((R, N), dtype=torch.float64, device=dev)
zref = {}
lanes_list = [int(x) for x in os.environ.get("FUSE_LANES", "1,2,4").split(",")]
for fuse in (0, 1):
    for lanes in lanes_list:
        os.environ["GSP_FFT_FUSE"] = str(fuse); os.environ["GSP_FFT_LANES"] = str(lanes)
        plan = gsp.FFTPlan(lib, st, dims, [0.0] * 3, [1.0] * 3)
        for mode in ("inject", "rng") if os.environ.get("FUSE_RNG", "0") == "1" else ("inject",):
            best = 1e9
            for _ in range(4):
                torch.cuda.synchronize()
                plan.sample_dev(R, w.data_ptr() if mode == "inject" else 0, 7, 0, 1.0, 0.0, 0, None, z.data_ptr())
                best = min(best, lib.last_sample_ms())
            torch.cuda.synchronize()
            if mode not in zref:
                zref[mode] = z.clone(); same = True
            else:
                same = bool(torch.equal(z, zref[mode]))
            print(f"fuse {fuse} lanes {lanes} {mode:6s}: {best / R * 1e3:7.1f} us/real {R / best * 1e3:7.0f} real/s identical={same}", flush=True)
        plan.close()
# per-kernel times, one lane, serialised
for fuse in (0, 1):
    os.environ["GSP_FFT_FUSE"] = str(fuse); os.environ["GSP_FFT_LANES"] = "1"
    plan = gsp.FFTPlan(lib, st, dims, [0.0] * 3, [1.0] * 3)
    lib.profile_enable(True)
    plan.sample_dev(16, w.data_ptr(), 7, 0, 1.0, 0.0, 0, None, z.data_ptr())
    prof = lib.profile_read()
    lib.profile_enable(False)
    print("fuse", fuse, {k: round(v["ms"] / v["launches"] * 1e3, 1) for k, v in prof.items()}, flush=True)
    plan.close()
