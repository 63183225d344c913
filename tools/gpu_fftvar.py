"""A/B of compile-time variants of the FFT kernels (development tool): one library file per variant, FFTSIM 256^3."""
import os, sys, glob
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gsp_b200 as gsp, gsp_oracle as O
from helpers import aniso3
import torch
dev = torch.device("cuda:0")
dims = (256, 256, 256); N = 256 ** 3; R = 16
st = aniso3(O.SPHERICAL, 1.0, (40.0, 20.0, 10.0), 30.0)
w = torch.rand((R, N), dtype=torch.float64, device=dev)
z = torch.empty((R, N), dtype=torch.float64, device=dev)
# library baseline: cuFFT through torch.fft (rfftn + irfftn only, no spectral multiply)
x = w[0].view(256, 256, 256)
for _ in range(3):
    y = torch.fft.irfftn(torch.fft.rfftn(x), s=(256, 256, 256))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(R):
    y = torch.fft.irfftn(torch.fft.rfftn(w[r].view(256, 256, 256)), s=(256, 256, 256))
e1.record(); torch.cuda.synchronize()
print(f"cuFFT (torch.fft.rfftn + irfftn, f64, no multiply): {e0.elapsed_time(e1) / R * 1e3:.1f} us/real", flush=True)
libs = sorted(glob.glob(os.path.join(ROOT, "geostatsprocesses.jl_b200", "libgspb200*.so")))
for path in libs:
    for lanes in (1, 4):
        os.environ["GSP_FFT_LANES"] = str(lanes)
        lib = gsp.Library(path)
        plan = gsp.FFTPlan(lib, st, dims, [0.0] * 3, [1.0] * 3)
        best = 1e9
        for _ in range(4):
            plan.sample_dev(R, w.data_ptr(), 0, 0, 1.0, 0.0, 0, None, z.data_ptr())
            best = min(best, lib.last_sample_ms())
        line = f"{os.path.basename(path)} lanes {lanes}: {best / R * 1e3:7.1f} us/real"
        if lanes == 1:
            lib.profile_enable(True)
            plan.sample_dev(R, w.data_ptr(), 0, 0, 1.0, 0.0, 0, None, z.data_ptr())
            prof = lib.profile_read()
            lib.profile_enable(False)
            line += " | " + " ".join(f"{k.replace('fft_', '')} {v['ms'] / v['launches'] * 1e3:.1f}" for k, v in prof.items())
        print(line, flush=True)
        plan.close(); lib.close()
