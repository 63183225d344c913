"""bench.py - Gaussian realizations/s of the B200-native FFTSIM / LUSIM hot path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: the oracle restatement on the host cores)

Headline workload (config.workload): FFTSIM on a 3-D CartesianGrid 256^3, anisotropic SphericalCovariance
(ranges 40/20/10, 30 deg z-rotation), 64 realizations per GPU per step (BASELINE.json configs[3]: 512
realizations over 8 GPUs), injected uniform noise resident in HBM.  A "step" = one pass of the hot
path over that batch.  Realizations shard over ranks with no data-path collective ("scaling": "weak").
`e2e`: the same metric through gsp_fft_sample with pinned HOST buffers (H2D + D2H inside the timed region), the measured
pinned-copy rate of the box beside it, and the resident variant (ensemble stays in HBM, mean + variance maps return).

LUSIM half of the metric (`lusim`, every N): rank 0 opens ONE context over all N GPUs of the job - the library's multi-device model -
and times configs[2] (C3: 16,384 nodes + 1,000 data, 1,000 realizations) and configs[4] (C5: bivariate, 32,768 nodes, 4,096
realizations): plan (assembly + distributed Cholesky + d2), resident sampling, end to end with host noise / host fields; `roofline`
per configuration against the FP64 tensor peak measured IN THIS RUN (cuBLAS Dgemm through torch.matmul), `speedup_vs_n1` against a
1-device context in the same run, `cpu_baseline` (oracle on the host cores) at N = 1.  `lusim.c1` and `fftsim_c2` are one-line
results of configs[0] and configs[1].
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

if "reference" in sys.argv:
    # the CPU arm uses every host core whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1, which silently turned
    # OpenBLAS single-threaded in round 1's N>1 reference runs); must happen before numpy / scipy load their BLAS
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = (256, 256, 256)
RANGES = (40.0, 20.0, 10.0)
ANGLE = 30.0
REALS_PER_GPU = 64
E2E_REALS = 8          # realizations per e2e step (host pinned buffers, PCIe inside the timed region)
SPHERICAL, EXPONENTIAL = 1, 2
# LUSIM halves of the metric (BASELINE.json configs[2] and configs[4]); C1 / C2 are one-line extras
LUSIM_C3 = {"name": "c3", "dims": (128, 128), "nd": 1000, "kind": EXPONENTIAL, "range": 20.0, "R": 1000, "nvars": 1, "rho": None, "seed": 3,
            "workload": "LUSIM conditional 128x128 grid (16,384 nodes) + 1,000 data, ExponentialCovariance(range=20), 1,000 realizations"}
LUSIM_C5 = {"name": "c5", "dims": (256, 128), "nd": 500, "kind": SPHERICAL, "range": 20.0, "R": 4096, "nvars": 2, "rho": 0.7, "seed": 5,
            "workload": "bivariate LUSIM (rho = 0.7) 256x128 grid (32,768 nodes) + 500 shared data nodes, SphericalCovariance(range=20), "
                        "4,096 realizations of both variables"}
LUSIM_C1 = {"name": "c1", "dims": (50, 50), "nd": 0, "kind": SPHERICAL, "range": 20.0, "R": 100, "nvars": 1, "rho": None, "seed": 1,
            "workload": "LUSIM unconditional 50x50 grid, SphericalCovariance(range=20), 100 realizations"}


def shard_range(R: int, rank: int, world: int):
    """contiguous shard [r0, r1) of R realizations owned by `rank` (same rule as the library's multi-device split)."""
    base, rem = divmod(R, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def fft_structs():
    th = math.radians(ANGLE)
    R = np.array([[math.cos(th), -math.sin(th), 0.0], [math.sin(th), math.cos(th), 0.0], [0.0, 0.0, 1.0]])
    return [(SPHERICAL, 1.0, np.diag(1.0 / np.asarray(RANGES)) @ R.T)]


def lusim_structs(cfg):
    A = np.zeros((3, 3))
    A[0, 0] = A[1, 1] = 1.0 / cfg["range"]
    return [(cfg["kind"], 1.0, A)]   # C5: both marginals of [1 .7; .7 1] * Spherical are this structure (lusim.jl:132-137)


def lusim_data(cfg):
    N = cfg["dims"][0] * cfg["dims"][1]
    rng = np.random.default_rng(cfg["seed"])
    dinds = np.sort(rng.choice(N, cfg["nd"], replace=False)) if cfg["nd"] else np.zeros(0, dtype=np.int64)
    z = [rng.standard_normal(cfg["nd"]) * 0.5 for _ in range(cfg["nvars"])]
    return N, dinds, z


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
    threads = {"scipy_fft_workers": cores, "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS")}
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(cores)
        threads["blas"] = [f"{i.get('internal_api')} {i.get('num_threads')}" for i in threadpoolctl.threadpool_info() if i.get("user_api") == "blas"]
    except Exception:
        pass
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
        "threads": threads,
    }
    if not args.skip_lusim:
        # LUSIM half of the metric on the same host cores (outside the timed FFTSIM steps): C3 in full, C5 reduced + extrapolated
        line["lusim"] = {"cpu_baseline": cpu_lusim_baseline()}
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
    del z0

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

    # ---- what bounds e2e: pinned host<->device copy bandwidth of this box, all ranks copying at once (both directions concurrently,
    # the same 2 x 8N bytes per realization the e2e step moves).  e2e cannot exceed aggregate / (16 N) whatever the kernels do.
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    dwb = w[:Re]
    dzb = z[:Re]
    barrier()
    t0 = time.perf_counter()
    creps = 2
    for _ in range(creps):
        with torch.cuda.stream(s_in):
            dwb.copy_(hw, non_blocking=True)
        with torch.cuda.stream(s_out):
            hz.copy_(dzb, non_blocking=True)
    barrier()
    tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    host_copy_gbs = 2.0 * 8 * N * Re * creps * world / float(tc[0]) / 1e9
    hw.copy_(w[:Re].cpu())  # restore (dwb aliases w)
    del dwb, dzb

    # ---- the resident path of the same metric on every rank: Rg realizations simulated into a device-resident ensemble (on-device
    # noise), only the per-node mean and variance maps (2 x 8N bytes) return to pinned host memory - Ensemble's `fetch` hook design
    res_value = None
    if not args.skip_ensemble:
        hst = torch.empty((2, N), dtype=torch.float64).pin_memory()
        ens_r = plan.sample_ensemble(Rg, None, seed=11, first_real=r0)

        def resident_step():
            plan.sample_ensemble(Rg, None, seed=12, first_real=r0, ens=ens_r)
            lib.check(lib.lib.gsp_ensemble_mean(ens_r.h, hst[0].data_ptr()))
            lib.check(lib.lib.gsp_ensemble_var(ens_r.h, hst[1].data_ptr()))

        resident_step()
        barrier()
        t0 = time.perf_counter()
        resident_step()
        barrier()
        trs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(trs, op=dist.ReduceOp.MAX)
        res_value = Rg * world / float(trs[0])
        ens_r.close()
        del hst

    # ---- resident ensemble + statistics in HBM: what a user who wants mean / variance / quantile maps pays
    ens_line = None
    if rank == 0 and not args.skip_ensemble:
        ens_line = bench_ensemble(lib, plan, Rg, N, r0)

    # ---- LUSIM half of the metric (configs[2] and configs[4]): rank 0 drives ONE context over all `world` GPUs (the library's
    # multi-device model: one process, peer access over NVLink); the other ranks release their buffers and wait on the CPU
    # (gloo barrier - an NCCL barrier would spin a kernel on the GPUs being measured)
    lus = None
    if not args.skip_lusim:
        plan.close()
        plan = None
        del w, z, hw, hz
        torch.cuda.empty_cache()
        import datetime
        cpu_group = dist.new_group(backend="gloo", timeout=datetime.timedelta(minutes=30)) if world > 1 else None
        barrier()
        if rank == 0:
            try:
                for d_ in range(world):
                    fr, tot = torch.cuda.mem_get_info(d_)
                    print(f"[bench] before LUSIM: device {d_}: {fr >> 20} of {tot >> 20} MiB free", file=sys.stderr, flush=True)
                lus = bench_lusim_all(gsp, torch, world, args.skip_cpu)
            except Exception as ex:  # the headline line is still emitted; the other ranks must not hang in the barrier
                import traceback
                traceback.print_exc()
                lus = {"error": f"{type(ex).__name__}: {ex}"}
        if cpu_group is not None:
            dist.barrier(group=cpu_group)

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
        try:  # the same with the DRAM bytes ncu counted for the five kernels (per-kernel windows: stores still in L2 at a kernel's end are
            # not counted anywhere, so this is a LOWER bound of the traffic and the pass model an upper one)
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tj = json.load(f)
            ncu_bytes = sum(k["traffic_bytes"] for k in tj["kernels"].values()) if DIMS == (256, 256, 256) else None
        except Exception:
            ncu_bytes = None
        if ncu_bytes:
            roof.update({"pipeline_ncu_dram_bytes_per_realization": ncu_bytes,
                         "pipeline_dram_frac_ncu_bytes": ncu_bytes * Rg * args.steps / dev_max / 1e9 / pk["hbm_gbs"]})
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
                                           "note": "same call with w = NULL (on-device Philox noise); PCIe carries only the fields"},
                    "host_copy_ceiling": {"pinned_h2d_plus_d2h_GBps_all_ranks": host_copy_gbs,
                                          "realizations_per_s_at_that_bandwidth": host_copy_gbs * 1e9 / (16.0 * N),
                                          "note": "plain pinned copies with NO compute, all ranks at once, 8N bytes in and 8N bytes out per realization "
                                                  "(torch copy_ on two streams): the PCIe / host-memory rate of this box.  The e2e value sits at this "
                                                  "rate (within the few % the two measurements differ by): e2e is copy-bound, not kernel-bound"},
                    "resident_statistics_variant": None if res_value is None else {
                        "value": res_value, "unit": "realizations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16 * N,
                        "note": "realizations stay in HBM (gsp_fft_sample_ensemble, on-device noise); only the mean and variance maps "
                                "return to pinned host memory (gsp_ensemble_mean / _var) - the Ensemble `fetch` design of ensembles.jl:16"}},
            "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
            "invariants_ok": ok_invariants, "field_mean": mean0, "field_var": var0,
        }
        if cpu_sec is not None:
            line["cpu_baseline"] = {"value": 1.0 / cpu_sec, "unit": "realizations/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{args.cpu_reals} realization(s) of the same 256^3 workload with the oracle restatement "
                                              f"(scipy.fft workers={os.cpu_count()}); Julia is absent so the restatement stands in for the reference CPU path"}
        if lus is not None:
            if "fftsim_c2" in lus:
                line["fftsim_c2"] = lus.pop("fftsim_c2")
            line["lusim"] = lus
        if ens_line is not None:
            line["ensemble_statistics"] = ens_line
        emit(line)
    if plan is not None:
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


def measure_fp64_peak(torch, dev):
    """FP64 tensor-core denominator measured IN THIS RUN: cuBLAS Dgemm 8192^3 through torch.matmul (library call, not ours):
    best single call (burst) and ~2 s back to back (sustained).  Nominal B200 FP64 tensor peak: 40 TF/s."""
    n = 8192
    a = torch.randn((n, n), dtype=torch.float64, device=dev)
    b = torch.randn((n, n), dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    fl = 2.0 * n ** 3
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize(dev)
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(4, int(2000.0 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    e1.synchronize()
    sus = e0.elapsed_time(e1) / reps
    del a, b, c
    return {"burst_tflops": fl / best / 1e9, "sustained_tflops": fl / sus / 1e9, "nominal_tflops": 40.0,
            "how": f"torch.matmul f64 {n}^3 (cuBLAS Dgemm): best of 5 single calls; {reps} calls back to back"}


def bench_lusim_config(gsp, torch, lib, cfg, ndev, peak_tf, detailed=False, e2e=True):
    """one LUSIM configuration through the C ABI on the context `lib` (ndev devices driven by this process):
    plan (assembly + Cholesky + d2; the factorization is distributed over the devices), resident sampling with the on-device
    RNG (realizations sharded over the devices), and the end-to-end run with host noise / host fields in pinned memory."""
    N, dinds, z = lusim_data(cfg)
    nv, R, rho = cfg["nvars"], cfg["R"], cfg["rho"]
    dom = (gsp._lib.make_grid_domain(cfg["dims"], [0.0, 0.0], [1.0, 1.0]), None)
    st = lusim_structs(cfg)
    nd = len(dinds)

    def make_plans(share=True):
        # both marginals of C5 are the same covariance over the same data nodes: like `rand(...; method=LUSIM())` of the host layer the
        # second variable shares the first one's factor (gsp_lu_plan_create_like) and only gets its own d2
        plans = [gsp.LUPlan(lib, st, dom, dinds + 1 if nd else None, z[0] if nd else None, 0.0)]
        for j in range(1, nv):
            plans.append(gsp.LUPlan(lib, st, dom, dinds + 1 if nd else None, z[j] if nd else None, 0.0, like=plans[0] if share else None))
        return plans

    def sync():
        for d in range(ndev):
            torch.cuda.synchronize(d)

    for p in make_plans():  # warm-up (allocator pools, module load on every device)
        p.close()
    sync()
    # the plan is built once per ensemble: best of two timed builds (the first one is closed before the second starts: at most
    # one set of plans is alive, 8.7 GB per variable and device at C5)
    best = None
    for rep in range(2):
        t0 = time.perf_counter()
        plans = make_plans()
        plan_s = time.perf_counter() - t0
        tm = [p.times() for p in plans]
        if best is None or plan_s < best[0]:
            best = (plan_s, tm)
        if rep == 0:
            for p in plans:
                p.close()
    plan_s, tm = best
    asm_ms, fac_ms, solve_ms = (sum(t[k] for t in tm) for k in range(3))
    Ns = plans[0].Ns
    Np = (nd + 127) // 128 * 128 + (Ns + 127) // 128 * 128
    nfact = 1  # factorizations actually run (the variables share one factor)
    f_chol = nfact * Np ** 3 / 3.0
    f_lz = nv * float(Ns) ** 2 * R

    # ---- resident sampling (device RNG, fields stay in HBM, sharded over the devices)
    ens = [gsp.DeviceEnsemble(lib, N, R) for _ in range(nv)]

    def sample_resident():
        plans[0].sample_ensemble(R, None, seed=7, stream=0, ens=ens[0])
        if nv == 2:
            plans[1].sample_ensemble(R, None, seed=7, stream=1, rho=rho, ens=ens[1])

    sample_resident()
    sync()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        sample_resident()
        ts.append(time.perf_counter() - t0)
    sample_s = min(ts)
    mean_err = None
    if nd:
        m = ens[0].mean()
        mean_err = float(np.abs(m[dinds] - z[0]).max())   # data nodes: every realization carries z1
    for e in ens:
        e.close()
    out = {"workload": cfg["workload"], "n_devices": ndev, "plan_wall_s": plan_s, "assemble_device_ms": asm_ms, "factor_device_ms": fac_ms,
           "solve_d2_device_ms": solve_ms, "factor_tflops": f_chol / fac_ms / 1e9,
           "sample_resident_wall_ms": sample_s * 1e3, "sample_tflops": f_lz / sample_s / 1e12,
           "factor_plus_sample_tflops": (f_chol + f_lz) / (fac_ms / 1e3 + sample_s) / 1e12,
           "realizations_per_s_sampling_resident": R / sample_s,
           "realizations_per_s_plan_plus_resident_sampling": R / (plan_s + sample_s),
           "factorizations": nfact, "shared_factor": nv > 1,
           "flops": {"cholesky": f_chol, "L_times_W": f_lz,
                     "note": "Np^3/3 per FACTORIZATION RUN (padded joint matrix; variables with the same marginal covariance and data nodes share "
                             "one factor) and Ns^2 R per variable (SURVEY 8d)"},
           "data_mean_abs_err_resident": mean_err}
    if peak_tf:
        out["factor_plus_sample_frac_of_peak"] = out["factor_plus_sample_tflops"] / (peak_tf * ndev)
        out["factor_frac_of_peak"] = out["factor_tflops"] / (peak_tf * ndev)
        out["peak_tflops_used"] = peak_tf * ndev

    if nv > 1:
        # for comparison: every variable factored on its own, as the reference's `map` over the variables does (lusim.jl:66-107)
        t0 = time.perf_counter()
        unshared = make_plans(share=False)
        out["plan_wall_s_unshared"] = time.perf_counter() - t0
        out["factor_device_ms_unshared"] = sum(p.times()[1] for p in unshared)
        for p in unshared:
            p.close()

    # ---- end to end: plan + host-pointer sampling, injected host noise and host fields in pinned memory (H2D + D2H inside)
    if e2e:
        hW = [torch.empty((R, Ns), dtype=torch.float64, pin_memory=True).normal_() for _ in range(nv)]
        hZ = [torch.empty((R, N), dtype=torch.float64, pin_memory=True) for _ in range(nv)]

        def sample_host():
            lib.check(lib.lib.gsp_lu_sample(plans[0].h, R, hW[0].data_ptr(), 0, 0, 0, math.nan, None, hZ[0].data_ptr()))
            if nv == 2:
                lib.check(lib.lib.gsp_lu_sample(plans[1].h, R, hW[1].data_ptr(), 0, 1, 0, rho, hW[0].data_ptr(), hZ[1].data_ptr()))

        sample_host()
        te = []
        for _ in range(2):
            t0 = time.perf_counter()
            sample_host()
            te.append(time.perf_counter() - t0)
        e2e_s = min(te)
        exact = all(bool(np.array_equal(hZ[j].numpy()[:, dinds], np.repeat(z[j][None, :], R, 0))) for j in range(nv)) if nd else None
        out.update({"e2e_sample_wall_s": e2e_s, "realizations_per_s_end_to_end": R / (plan_s + e2e_s),
                    "e2e_h2d_bytes": 8 * Ns * R * nv, "e2e_d2h_bytes": 8 * N * R * nv, "data_honoured_exactly": exact})
        del hW, hZ

    # ---- per-kernel device times of one profiled plan + sampling (1-device contexts only: events serialise the launches)
    if detailed and ndev == 1:
        lib.profile_enable(True)
        for p in make_plans():
            p.close()
        prof_plan = lib.profile_read()
        dev = torch.device("cuda", lib.devices[0])
        W = torch.randn((R, Ns), dtype=torch.float64, device=dev)
        Z = torch.empty((R, N), dtype=torch.float64, device=dev)
        plans[0].sample_dev(R, W.data_ptr(), Ns, 0, 0, 0, math.nan, None, Z.data_ptr(), N)
        lib.profile_enable(True)
        ms = []
        for _ in range(3):
            plans[0].sample_dev(R, W.data_ptr(), Ns, 0, 0, 0, math.nan, None, Z.data_ptr(), N)
            ms.append(lib.last_sample_ms())
        prof_samp = lib.profile_read()
        lib.profile_enable(False)
        del W, Z
        out.update({"sample_device_ms_injected_noise": min(ms), "sample_device_tflops": float(Ns) ** 2 * R / min(ms) / 1e9,
                    "factor_kernel_ms_sum_serialised": sum(v["ms"] for k, v in prof_plan.items() if k.startswith("gemm_dmma") or k.startswith("potrf")),
                    "kernel_ms_plan": {k: round(v["ms"], 3) for k, v in prof_plan.items()},
                    "kernel_launches_plan": {k: v["launches"] for k, v in prof_plan.items()},
                    "kernel_ms_sample_x3": {k: round(v["ms"], 3) for k, v in prof_samp.items()}})
    for p in plans:
        p.close()
    return out


def bench_fftsim_c2(gsp, torch, lib):
    """BASELINE configs[1] in one line: FFTSIM 2-D 1024x1024, GaussianCovariance(range=50), 64 realizations, noise and fields in HBM"""
    A = np.zeros((3, 3))
    A[0, 0] = A[1, 1] = 1.0 / 50.0
    dims = (1024, 1024)
    N, R = dims[0] * dims[1], 64
    t0 = time.perf_counter()
    plan = gsp.FFTPlan(lib, [(3, 1.0, A)], dims, [0.0, 0.0], [1.0, 1.0])
    plan_s = time.perf_counter() - t0
    dev = torch.device("cuda", lib.devices[0])
    w = torch.rand((R, N), dtype=torch.float64, device=dev)
    z = torch.empty((R, N), dtype=torch.float64, device=dev)
    ms = []
    for _ in range(4):
        plan.sample_dev(R, w.data_ptr(), 0, 0, 1.0, 0.0, 0, None, z.data_ptr())
        ms.append(lib.last_sample_ms())
    plan.close()
    best = min(ms[1:])
    return {"workload": "FFTSIM 2D CartesianGrid 1024x1024, GaussianCovariance(range=50), 64 realizations, injected noise resident in HBM",
            "plan_wall_s": plan_s, "sample_device_ms": best, "realizations_per_s": R / best * 1e3,
            "alg_GBps": 20.0 * N * R / best / 1e6, "frac_of_hbm_peak_alg_bytes": 20.0 * N * R / best / 1e6 / peaks()["hbm_gbs"]}


def bench_lusim_all(gsp, torch, world, skip_cpu):
    """LUSIM half of the metric on rank 0: C3 and C5 (+ C1) on a context over ALL `world` GPUs of the job (the other ranks idle in
    a CPU-side barrier meanwhile), the same on one GPU when world > 1 (speed-up inside one line), the in-run FP64 peak and, at
    N = 1, the oracle's CPU time for the same configurations."""
    dev0 = torch.device("cuda", 0)
    peak = measure_fp64_peak(torch, dev0)
    ptf = peak["burst_tflops"]
    out = {"fp64_peak": peak, "peak_note": "fractions are of the in-run cuBLAS Dgemm burst figure x number of devices"}
    libN = gsp.Library(devices=list(range(world)))
    out["c3"] = bench_lusim_config(gsp, torch, libN, LUSIM_C3, world, ptf, detailed=True)
    out["c5"] = bench_lusim_config(gsp, torch, libN, LUSIM_C5, world, ptf)
    out["c1"] = bench_lusim_config(gsp, torch, libN, LUSIM_C1, world, ptf, e2e=False)
    for key in ("c3", "c5"):  # the roofline object of the LUSIM half: FP64 tensor (DMMA) flops of Cholesky + L*W over the in-run peak
        c = out[key]
        c["roofline"] = {"bound": "tensor", "achieved": c["factor_plus_sample_tflops"], "peak": ptf * world, "unit": "TFLOP/s",
                         "frac": c["factor_plus_sample_tflops"] / (ptf * world), "traffic": None,
                         "peak_source": f"in-run cuBLAS Dgemm burst x {world} device(s)",
                         "flops_counted": "Np^3/3 per factorization run + Ns^2 R per variable (SURVEY 8d), time = factor (CUDA events) + resident sampling (wall)"}
    libN.close()
    if world > 1:
        lib1 = gsp.Library(devices=[0])
        for key, cfg in (("c3", LUSIM_C3), ("c5", LUSIM_C5)):
            one = bench_lusim_config(gsp, torch, lib1, cfg, 1, ptf)
            out[key]["one_gpu_same_run"] = {k: one[k] for k in ("plan_wall_s", "factor_device_ms", "sample_resident_wall_ms",
                                                                "realizations_per_s_end_to_end", "realizations_per_s_plan_plus_resident_sampling")}
            out[key]["speedup_vs_n1"] = {"factor": one["factor_device_ms"] / out[key]["factor_device_ms"],
                                         "sample_resident": one["sample_resident_wall_ms"] / out[key]["sample_resident_wall_ms"],
                                         "realizations_per_s_end_to_end": out[key]["realizations_per_s_end_to_end"] / one["realizations_per_s_end_to_end"],
                                         "realizations_per_s_plan_plus_resident_sampling":
                                             out[key]["realizations_per_s_plan_plus_resident_sampling"] / one["realizations_per_s_plan_plus_resident_sampling"]}
        lib1.close()
    try:
        lib1 = gsp.Library(devices=[0])
        out["fftsim_c2"] = bench_fftsim_c2(gsp, torch, lib1)
        lib1.close()
    except Exception as ex:
        out["fftsim_c2"] = {"error": f"{type(ex).__name__}: {ex}"}
    if world == 1 and not skip_cpu:
        out["cpu_baseline"] = cpu_lusim_baseline()
    return out


def cpu_lusim_baseline(full_c3=True):
    """the oracle restatement of lusim.jl:38-175 on the host cores (SciPy / OpenBLAS potrf, trsm, gemm), timed beside the GPU numbers:
    C3 in full; C5 at 1/4 of the nodes (8,192 + 125 data, 256 realizations) with the stated N^3 / N^2 R extrapolation."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gsp_oracle as O
    try:
        import threadpoolctl
        blas = [f"{i.get('internal_api')} {i.get('num_threads')} threads" for i in threadpoolctl.threadpool_info() if i.get("user_api") == "blas"]
    except Exception:
        blas = ["unknown"]

    def run(cfg, dims, nd, R):
        N = dims[0] * dims[1]
        rng = np.random.default_rng(cfg["seed"])
        dinds = np.sort(rng.choice(N, nd, replace=False))
        z1 = rng.standard_normal(nd) * 0.5
        A = np.zeros((3, 3))
        A[0, 0] = A[1, 1] = 1.0 / cfg["range"]
        st = [O.Structure(cfg["kind"], 1.0, A)]
        coords = O.grid_centroids(dims, [0, 0], [1, 1])
        t0 = time.perf_counter()
        pre = O.lusim_preprocess(st, coords, dinds, z1, 0.0)
        t_pre = time.perf_counter() - t0
        W = rng.standard_normal((N - nd, R))
        t0 = time.perf_counter()
        O.lusim_sample(pre, W)
        t_s = time.perf_counter() - t0
        # the O(N^2) covariance assembly inside the preprocess, timed on its own so that the extrapolation scales it by N^2, not N^3
        t0 = time.perf_counter()
        O.pairwise(st, coords[pre.sinds])
        O.pairwise(st, coords[dinds], coords[pre.sinds])
        O.pairwise(st, coords[dinds])
        t_asm = min(time.perf_counter() - t0, t_pre)
        return t_pre, t_s, t_asm

    out = {"kind": "port", "cores": os.cpu_count(), "blas": blas,
           "note": "Julia is absent from the image: the oracle restatement (SciPy/OpenBLAS) stands in for the reference's LAPACK path"}
    if full_c3:
        tp, ts, ta = run(LUSIM_C3, LUSIM_C3["dims"], LUSIM_C3["nd"], LUSIM_C3["R"])
        out["c3"] = {"preprocess_s": tp, "of_which_assembly_s": ta, "sample_s": ts, "realizations_per_s_end_to_end": LUSIM_C3["R"] / (tp + ts),
                     "realizations_per_s_sampling": LUSIM_C3["R"] / ts, "sample": "the full configuration (one variable, 1,000 realizations)"}
    tp, ts, ta = run(LUSIM_C5, (128, 64), 125, 256)
    pre_x = 2 * ((tp - ta) * 4.0 ** 3 + ta * 4.0 ** 2)
    s_x = ts * 2 * 4.0 ** 2 * (LUSIM_C5["R"] / 256.0)
    out["c5"] = {"preprocess_s_measured_reduced": tp, "of_which_assembly_s": ta, "sample_s_measured_reduced": ts,
                 "preprocess_s_extrapolated": pre_x, "sample_s_extrapolated": s_x,
                 "realizations_per_s_end_to_end": LUSIM_C5["R"] / (pre_x + s_x),
                 "sample": "8,192 nodes + 125 data, 256 realizations, one variable; extrapolated x2 variables, factorization part x4^3 (N^3), "
                           "assembly part x4^2 (N^2), L*W x4^2 x16 (N^2 R) to the full configuration"}
    return out


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
