"""bench.py - Gaussian realizations/s of the B200-native FFTSIM / LUSIM hot path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: the oracle restatement on the host cores)

Workload (config.workload): FFTSIM on a 3-D CartesianGrid 256^3, anisotropic SphericalCovariance
(ranges 40/20/10, 30 deg z-rotation), 64 realizations per GPU per step (BASELINE.json configs[3]: 512
realizations over 8 GPUs), injected uniform noise resident in HBM.  A "step" = one pass of the hot
path over that batch.  Realizations shard over ranks with no data-path collective ("scaling": "weak").
LUSIM 16k nodes (configs[2]) is measured in the same run and reported under "lusim".
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = (256, 256, 256)
RANGES = (40.0, 20.0, 10.0)
ANGLE = 30.0
REALS_PER_GPU = 64
E2E_REALS = 8          # realizations per e2e step (host pinned buffers, PCIe inside the timed region)
LUSIM_GRID = (128, 128)
LUSIM_ND = 1000
LUSIM_R = 1000
SPHERICAL, EXPONENTIAL = 1, 2


def shard_range(R: int, rank: int, world: int):
    """contiguous shard [r0, r1) of R realizations owned by `rank` (same rule as the library's multi-device split)."""
    base, rem = divmod(R, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def fft_structs():
    th = math.radians(ANGLE)
    R = np.array([[math.cos(th), -math.sin(th), 0.0], [math.sin(th), math.cos(th), 0.0], [0.0, 0.0, 1.0]])
    return [(SPHERICAL, 1.0, np.diag(1.0 / np.asarray(RANGES)) @ R.T)]


def lusim_structs():
    A = np.zeros((3, 3))
    A[0, 0] = A[1, 1] = 1.0 / 20.0
    return [(EXPONENTIAL, 1.0, A)]


def peaks():
    p = {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p = {"hbm_gbs": float(m["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    except Exception:
        pass
    return p


class ClockSampler:
    def __init__(self, dev):
        self.dev = dev
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.dev),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arms
_cpu_cache = {}


def cpu_fft_sampler():
    """oracle FFTSIM 256^3 on the host cores: returns fn(seed) -> seconds for ONE realization (sampling only;
    the once-per-ensemble spectrum build is excluded on both arms, the noise is pre-drawn like the injected GPU noise)."""
    if "F" not in _cpu_cache:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import gsp_oracle as O

        st = [O.Structure(*s) for s in fft_structs()]
        _cpu_cache["F"] = O.fftsim_preprocess(st, DIMS, [0.0] * 3, [1.0] * 3)
        _cpu_cache["O"] = O
    F, O = _cpu_cache["F"], _cpu_cache["O"]
    n = int(np.prod(DIMS))

    def one(seed):
        w = np.random.default_rng(seed).random(n)
        t0 = time.perf_counter()
        O.fftsim_sample(F, w, 1.0, 0.0)
        return time.perf_counter() - t0

    return one


def cpu_fft_realizations(nreal, seed=4):
    one = cpu_fft_sampler()
    one(seed)  # warm-up (thread pool, page faults)
    return sum(one(seed + 1 + k) for k in range(nreal)) / nreal


def run_reference(args):
    """--impl reference: the reference recipe on the host cores.  The reference itself (Julia) cannot run in this
    image, so the timed code is the oracle restatement (scipy.fft / pocketfft, all host threads): kind = "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = 1  # bounded sample: each step = 1 realization of the same 256^3 workload
    one = cpu_fft_sampler()
    for k in range(max(args.warmup, 1)):
        one(k)
    secs = [one(100 + k) for k in range(args.steps * per_step)]
    total = sum(secs)
    val = args.steps * per_step / total
    line = {
        "impl": "reference", "metric": "Gaussian realizations/s", "value": val, "unit": "realizations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "FFTSIM 3D CartesianGrid 256^3, anisotropic SphericalCovariance (40,20,10; 30deg z-rotation), "
                               "per-realization sampling (bounded sample: 1 realization per step)", "reals_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": "realizations/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps * per_step} realization(s) of the 256^3 workload, scipy.fft workers={cores}; "
                                   "Julia is absent from the image, so the oracle restatement stands in for the reference"},
        "e2e": {"value": val, "unit": "realizations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import gsp_b200 as gsp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything libraries print while we run (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = gsp.Library(devices=[local])  # raises without the CUDA extension / a GPU: no fallback

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N = int(np.prod(DIMS))
    Rg = args.reals_per_gpu
    r0 = rank * Rg  # weak scaling: every rank owns Rg realizations; global index offsets keep RNG streams distinct
    plan = gsp.FFTPlan(lib, fft_structs(), DIMS, [0.0] * 3, [1.0] * 3)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4 + rank)
    w = torch.rand((Rg, N), dtype=torch.float64, device=dev, generator=gen)   # injected noise, resident in HBM
    z = torch.empty((Rg, N), dtype=torch.float64, device=dev)

    def step():
        plan.sample_dev(Rg, w.data_ptr(), 0, r0, 1.0, 0.0, 0, None, z.data_ptr())
        return lib.last_sample_ms()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up so that nvidia-smi is already sampling when the timed region begins
    step()
    time.sleep(0.5 if rank == 0 else 0.0)
    for _ in range(max(args.warmup, 3)):
        step()
    launches0 = lib.kernel_launches()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += step()
    barrier()
    wall = time.perf_counter() - t0
    launches = lib.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([wall, dev_ms / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    wall_max, dev_max = float(tmax[0]), float(tmax[1])
    value = Rg * world * args.steps / wall_max

    # sanity on the produced fields (not timed): exact mean mu and variance sill per realization
    z0 = z[0]
    mean0 = float(z0.mean())
    var0 = float((z0 * z0).sum() / (N - 1))
    ok_invariants = abs(mean0) < 1e-10 and abs(var0 - 1.0) < 1e-10

    # ---- per-kernel device times (instrumented pass right after the timed region, CUDA events per launch)
    # The timed region runs 4 realizations concurrently (lanes), which makes per-launch event times overlap; the per-kernel
    # numbers therefore come from a second plan restricted to ONE lane (kernels back to back on one stream).
    prof = None
    if rank == 0:
        os.environ["GSP_FFT_LANES"] = "1"
        plan1 = gsp.FFTPlan(lib, fft_structs(), DIMS, [0.0] * 3, [1.0] * 3)
        os.environ.pop("GSP_FFT_LANES")
        Rp = min(Rg, 16)
        plan1.sample_dev(Rp, w.data_ptr(), 0, r0, 1.0, 0.0, 0, None, z.data_ptr())
        lib.profile_enable(True)
        plan1.sample_dev(Rp, w.data_ptr(), 0, r0, 1.0, 0.0, 0, None, z.data_ptr())
        prof = lib.profile_read()
        lib.profile_enable(False)
        plan1.close()

    # ---- e2e: same metric through the host-pointer C-ABI call with pinned host buffers (H2D + D2H inside)
    Re = args.e2e_reals
    hw = torch.empty((Re, N), dtype=torch.float64).pin_memory()
    hz = torch.empty((Re, N), dtype=torch.float64).pin_memory()
    hw.copy_(w[:Re].cpu())

    def e2e_step():
        rc = lib.lib.gsp_fft_sample(plan.h, Re, hw.data_ptr(), 0, r0, 1.0, 0.0, 0, None, hz.data_ptr())
        lib.check(rc)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t0
    te = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = Re * world * e2e_steps / float(te[0])
    e2e_match = bool(torch.equal(hz[0], z[0].cpu()))

    # the same call with w = NULL: noise from the on-device counter RNG, as `rand(process, domain, n)` without an injected
    # array would run (fields still go back to pinned host memory); reported beside the injected-noise figure, not instead of it
    def e2e_rng_step():
        lib.check(lib.lib.gsp_fft_sample(plan.h, Re, None, 4, r0, 1.0, 0.0, 0, None, hz.data_ptr()))

    e2e_rng_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_rng_step()
    barrier()
    tr = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
    e2e_rng_value = Re * world * e2e_steps / float(tr[0])

    # ---- resident ensemble + statistics in HBM: what a user who wants mean / variance / quantile maps pays
    ens_line = None
    if rank == 0 and not args.skip_ensemble:
        ens_line = bench_ensemble(lib, plan, Rg, N, r0)

    # ---- LUSIM 16k nodes (configs[2]) on rank 0's GPU, reported beside the headline
    lus = None
    if rank == 0 and not args.skip_lusim:
        lus = bench_lusim(lib, gsp, torch, dev)

    if rank == 0:
        pk = peaks()
        alg_bytes = 20.0 * N  # per realization: noise in 8N + half-spectrum F 4N + field out 8N
        roof = {"bound": "hbm", "peak": pk["hbm_gbs"], "unit": "GB/s", "peak_source": pk["source"], "traffic": None}
        if prof:
            dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
            name, rec = dom
            total = sum(v["ms"] for v in prof.values())
            # algorithmic bytes one launch of that kernel must move (one realization per launch)
            per_launch = {"fft_xpass_fwd": 8.0 * N + 16.0 * plan_nh(DIMS), "fft_xpass_inv": 8.0 * N + 16.0 * plan_nh(DIMS),
                          "fft_strided_fwd": 32.0 * plan_nh(DIMS), "fft_strided_inv": 32.0 * plan_nh(DIMS),
                          "fft_strided_fwd_mul_inv": 40.0 * plan_nh(DIMS)}.get(name, alg_bytes)
            avg_ms = rec["ms"] / rec["launches"]
            ach = per_launch / avg_ms / 1e6
            try:  # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/)
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                    tj = json.load(f)
                if DIMS == (256, 256, 256) and name in tj["kernels"]:
                    roof["traffic"] = tj["kernels"][name]["traffic_bytes"]
                    roof["traffic_source"] = tj["source"] + "; L2 write-back leaves part of the stores in L2 at kernel end"
            except Exception:
                pass
            roof.update({"kernel": name, "achieved": ach, "frac": ach / pk["hbm_gbs"], "kernel_avg_ms": avg_ms,
                         "kernel_share_of_step": rec["ms"] / total, "alg_bytes_per_launch": per_launch,
                         "kernel_ms": {k: round(v["ms"], 4) for k, v in prof.items()}})
        pipe = alg_bytes * Rg * args.steps / dev_max / 1e9
        # pass-model bytes: 5 passes, each reads and writes a half spectrum (or the real field), + F in the z pass
        pass_bytes = 8.0 * N * 2 + 16.0 * plan_nh(DIMS) * 8 + 8.0 * plan_nh(DIMS)
        dram = pass_bytes * Rg * args.steps / dev_max / 1e9
        roof.update({"pipeline_alg_bytes_per_realization": alg_bytes, "pipeline_achieved": pipe, "pipeline_frac": pipe / pk["hbm_gbs"],
                     "pipeline_pass_model_bytes_per_realization": pass_bytes, "pipeline_dram_achieved": dram,
                     "pipeline_dram_frac": dram / pk["hbm_gbs"],
                     "note": "kernel_* figures: one lane (kernels serialised); pipeline_* figures: the timed region (4 concurrent lanes)"})
        cpu_sec = cpu_fft_realizations(args.cpu_reals) if world == 1 and not args.skip_cpu else None
        line = {
            "metric": "Gaussian realizations/s", "value": value, "unit": "realizations/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": wall_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "FFTSIM 3D CartesianGrid 256^3, anisotropic SphericalCovariance (40,20,10; 30deg z-rotation), "
                                   f"{Rg} realizations per GPU per step, injected U[0,1) noise resident in HBM",
                       "reals_per_gpu_per_step": Rg, "global_reals_per_step": Rg * world, "l2": "inputs larger than L2 (8.6 GB noise per GPU per step)",
                       "parallelism": f"realizations sharded over {world} GPU(s), no collective on the data path; "
                                      "4 realizations in flight per GPU on concurrent streams"},
            "device_ms_per_step": dev_max / args.steps * 1e3,
            "e2e": {"value": e2e_value, "unit": "realizations/s", "h2d_bytes_per_step": 8 * N * Re, "d2h_bytes_per_step": 8 * N * Re,
                    "reals_per_step": Re, "steps": e2e_steps, "matches_device_path": e2e_match,
                    "api": "gsp_fft_sample (host pointers, pinned), 3-stream H2D/compute/D2H pipeline",
                    "device_rng_variant": {"value": e2e_rng_value, "unit": "realizations/s", "h2d_bytes_per_step": 0,
                                           "d2h_bytes_per_step": 8 * N * Re,
                                           "note": "same call with w = NULL (on-device Philox noise); PCIe carries only the fields"}},
            "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
            "invariants_ok": ok_invariants, "field_mean": mean0, "field_var": var0,
        }
        if cpu_sec is not None:
            line["cpu_baseline"] = {"value": 1.0 / cpu_sec, "unit": "realizations/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{args.cpu_reals} realization(s) of the same 256^3 workload with the oracle restatement "
                                              f"(scipy.fft workers={os.cpu_count()}); Julia is absent so the restatement stands in for the reference CPU path"}
        if lus is not None:
            line["lusim"] = lus
        if ens_line is not None:
            line["ensemble_statistics"] = ens_line
        emit(line)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def bench_ensemble(lib, plan, R, N, r0):
    """R realizations (device RNG) simulated into a device-resident ensemble, then mean, variance, cdf and three quantile maps
    written into PINNED host buffers: only the n-vectors of results cross PCIe (ensembles.jl:42-52 run in HBM instead of O(n R)
    host loops).  Timed through the C ABI (gsp_fft_sample_ensemble, gsp_ensemble_*)."""
    import torch

    L = lib.lib
    hbuf = torch.empty((3, N), dtype=torch.float64).pin_memory()
    ps = np.array([0.1, 0.5, 0.9])

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, time.perf_counter() - t0

    ens = plan.sample_ensemble(R, None, seed=3, first_real=r0)  # allocation (8 N R bytes) + warm-up, not timed
    _, t_sim = timed(lambda: plan.sample_ensemble(R, None, seed=4, first_real=r0, ens=ens))
    lib.check(L.gsp_ensemble_mean(ens.h, hbuf[0].data_ptr()))  # warm-up of the statistics kernels
    lib.profile_enable(True)
    _, t_mean = timed(lambda: lib.check(L.gsp_ensemble_mean(ens.h, hbuf[0].data_ptr())))
    mean_abs_max = float(hbuf[0].abs().max())
    _, t_var = timed(lambda: lib.check(L.gsp_ensemble_var(ens.h, hbuf[1].data_ptr())))
    var_mean = float(hbuf[1].mean())
    _, t_cdf = timed(lambda: lib.check(L.gsp_ensemble_cdf(ens.h, 0.0, hbuf[2].data_ptr())))
    cdf_mean = float(hbuf[2].mean())
    _, t_q = timed(lambda: lib.check(L.gsp_ensemble_quantile(ens.h, 3, ps.ctypes.data, hbuf.data_ptr())))
    med_abs_mean = float(hbuf[1].abs().mean())
    prof = lib.profile_read()
    lib.profile_enable(False)
    ens.close()
    pk = peaks()
    bytes_pass = 8.0 * N * R
    out = {"workload": f"{R} FFTSIM 256^3 realizations (on-device Philox noise) resident in HBM; mean, var, cdf(0), quantile([.1,.5,.9]) "
                       "per node into pinned host buffers",
           "simulate_s": t_sim, "realizations_per_s_simulate_resident": R / t_sim,
           "mean_s": t_mean, "var_s": t_var, "cdf_s": t_cdf, "quantile3_s": t_q,
           "realizations_per_s_simulate_plus_mean_var": R / (t_sim + t_mean + t_var),
           "realizations_per_s_simulate_plus_all_statistics": R / (t_sim + t_mean + t_var + t_cdf + t_q),
           "d2h_bytes_per_statistic": 8 * N, "d2h_bytes_if_realizations_were_downloaded": 8 * N * R,
           "kernel_ms": {k: round(v["ms"] / v["launches"], 3) for k, v in prof.items()},
           "checks": {"mean_abs_max": mean_abs_max, "var_mean": var_mean, "cdf0_mean": cdf_mean, "median_abs_mean": med_abs_mean}}
    for k in ("ens_moments", "ens_count"):
        if k in prof:
            gbs = bytes_pass / (prof[k]["ms"] / prof[k]["launches"]) / 1e6
            out[k + "_GBps"] = gbs
            out[k + "_frac_of_hbm_peak"] = gbs / pk["hbm_gbs"]
    return out


def plan_nh(dims):
    return (dims[0] // 2 + 1) * dims[1] * dims[2]


def bench_lusim(lib, gsp, torch, dev):
    """LUSIM conditional, 128x128 grid (16,384 nodes) + 1,000 hard data, ExponentialCovariance range 20, 1,000 realizations."""
    N = LUSIM_GRID[0] * LUSIM_GRID[1]
    rng = np.random.default_rng(3)
    dinds = np.sort(rng.choice(N, LUSIM_ND, replace=False))
    z1 = rng.standard_normal(LUSIM_ND) * 0.5
    dom = (gsp._lib.make_grid_domain(LUSIM_GRID, [0.0, 0.0], [1.0, 1.0]), None)
    gsp.LUPlan(lib, lusim_structs(), dom, dinds + 1, z1, 0.0).close()  # warm-up (allocator, module load)
    t0 = time.perf_counter()
    plan = gsp.LUPlan(lib, lusim_structs(), dom, dinds + 1, z1, 0.0)
    plan_s = time.perf_counter() - t0
    asm_ms, fac_ms, solve_ms = plan.times()  # CUDA-event stage timers of the un-profiled plan (look-ahead streams concurrent)
    lib.profile_enable(True)
    p2 = gsp.LUPlan(lib, lusim_structs(), dom, dinds + 1, z1, 0.0)
    prof_plan = lib.profile_read()
    p2.close()
    R = LUSIM_R
    W = torch.randn((R, plan.Ns), dtype=torch.float64, device=dev)
    Z = torch.empty((R, N), dtype=torch.float64, device=dev)
    for _ in range(2):
        plan.sample_dev(R, W.data_ptr(), plan.Ns, 0, 0, 0, math.nan, None, Z.data_ptr(), N)
    lib.profile_enable(True)
    ms = []
    for _ in range(3):
        plan.sample_dev(R, W.data_ptr(), plan.Ns, 0, 0, 0, math.nan, None, Z.data_ptr(), N)
        ms.append(lib.last_sample_ms())
    prof_samp = lib.profile_read()
    lib.profile_enable(False)
    sample_ms = min(ms)
    Np = (LUSIM_ND + 127) // 128 * 128 + (plan.Ns + 127) // 128 * 128
    f_chol = Np ** 3 / 3.0
    f_lz = float(plan.Ns) ** 2 * R
    chol_ms = sum(v["ms"] for k, v in prof_plan.items() if k.startswith("gemm_dmma") or k == "potrf_diag")
    # end to end through the host API with pinned buffers
    hW = torch.randn((R, plan.Ns), dtype=torch.float64).pin_memory()
    hZ = torch.empty((R, N), dtype=torch.float64).pin_memory()
    t0 = time.perf_counter()
    lib.check(lib.lib.gsp_lu_sample(plan.h, R, hW.data_ptr(), 0, 0, 0, math.nan, None, hZ.data_ptr()))
    e2e_s = time.perf_counter() - t0
    exact = bool(np.array_equal(hZ.numpy()[:, dinds], np.repeat(z1[None, :], R, 0)))
    plan.close()
    return {"workload": "LUSIM conditional 128x128 grid (16,384 nodes) + 1,000 data, ExponentialCovariance(range=20), 1,000 realizations",
            "plan_wall_s": plan_s, "assemble_device_ms": asm_ms, "factor_device_ms": fac_ms, "solve_d2_device_ms": solve_ms,
            "factor_tflops": f_chol / fac_ms / 1e9 if fac_ms else None,
            "factor_plus_sample_tflops": (f_chol + f_lz) / (fac_ms + sample_ms) / 1e9 if fac_ms else None,
            "fp64_peak_tflops": 35.5,
            "factor_plus_sample_frac_of_peak": (f_chol + f_lz) / (fac_ms + sample_ms) / 1e9 / 35.5 if fac_ms else None,
            "factor_kernel_ms_sum_serialised": chol_ms,
            "sample_device_ms": sample_ms, "sample_tflops": f_lz / sample_ms / 1e9,
            "realizations_per_s_sampling": R / sample_ms * 1e3, "realizations_per_s_end_to_end": R / (plan_s + e2e_s),
            "e2e_sample_wall_s": e2e_s, "data_honoured_exactly": exact,
            "kernel_ms_plan": {k: round(v["ms"], 3) for k, v in prof_plan.items()},
            "kernel_ms_sample_x3": {k: round(v["ms"], 3) for k, v in prof_samp.items()},
            "fp64_peak_note": "cuBLAS Dgemm 8192^3 measured 35.5 TF/s on this pool (tools/gpu_check.py); fractions are of that"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--reals-per-gpu", type=int, default=REALS_PER_GPU)
    ap.add_argument("--e2e-reals", type=int, default=E2E_REALS)
    ap.add_argument("--cpu-reals", type=int, default=4)
    ap.add_argument("--skip-lusim", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-ensemble", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
