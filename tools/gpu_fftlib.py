"""FFTSIM 256^3 timing for a given build of the library (argv[1] = path of the .so; A/B of compile-time variants)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import aniso3
import torch
path = sys.argv[1] if len(sys.argv) > 1 else None
lib = gsp.Library(path) if path else gsp.Library()
dev = torch.device("cuda:0")
dims = (256, 256, 256); N = 256 ** 3; R = 32
st = aniso3(O.SPHERICAL, 1.0, (40.0, 20.0, 10.0), 30.0)
w = torch.rand((R, N), dtype=torch.float64, device=dev)
z = torch.empty((R, N), dtype=torch.float64, device=dev)
ref = os.environ.get("FFTLIB_REF")
for lanes in [int(x) for x in os.environ.get("FFTLIB_LANES", "1,4").split(",")]:
    os.environ["GSP_FFT_LANES"] = str(lanes)
    plan = gsp.FFTPlan(lib, st, dims, [0.0] * 3, [1.0] * 3)
    for mode in ("inject", "rng"):
        best = 1e9
        for _ in range(4):
            plan.sample_dev(R, w.data_ptr() if mode == "inject" else 0, 7, 0, 1.0, 0.0, 0, None, z.data_ptr())
            best = min(best, lib.last_sample_ms())
        torch.cuda.synchronize()
        zc = z[:2].cpu().numpy()
        print(f"{os.path.basename(path or 'default')} lanes {lanes} {mode:6s}: {best / R * 1e3:7.1f} us/real {R / best * 1e3:7.0f} real/s mean {zc.mean():+.2e} var {zc.var():.6f} sum {float(np.abs(zc).sum()):.10e}", flush=True)
    if lanes == 1:
        lib.profile_enable(True)
        plan.sample_dev(8, w.data_ptr(), 7, 0, 1.0, 0.0, 0, None, z.data_ptr())
        prof = lib.profile_read()
        lib.profile_enable(False)
        print("   kernels us:", {k.replace("fft_", ""): round(v["ms"] / v["launches"] * 1e3, 1) for k, v in prof.items()}, flush=True)
    plan.close()
