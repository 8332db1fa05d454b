#!/usr/bin/env python
"""bench.py - triangle-steps/s of the DE (FP64) shallow-water timestep on B200.

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's own C/OpenMP
                                                            code on the host cores)

Workloads (`--config`, BASELINE.json `configs`):
  sweep      (default) configs[2]: synthetic rectangular_cross 2000x2000 = 16,000,000 triangles per
             GPU, smooth everywhere-wet fields, DE1 (rk2), Reflective boundaries, Manning 0.03, scalar
             rain Rate_operator.  Weak scaling: every rank owns a 16M-triangle strip, so --gpus 8 is the
             8000x4000 = 128M-triangle mesh of configs[3].
  tsunami    configs[1]: 1000x1000 cells (4M triangles), sloping beach + island, DE1, Manning 0.025,
             time-dependent set-stage boundary + Transmissive + Reflective.
  structures configs[4]: 4000x2000 cells (32M triangles), DE1 with rk3, Inlet_operator + Boyd box culvert.
A "step" is one full timestep (DE1: two flux evaluations).

One JSON line on stdout (rank 0).  `value` is device-timed (CUDA events on the library's stream) with
the state resident in HBM; `e2e` is the same metric through the public API - host numpy arrays in,
`for t in domain.evolve(yieldstep, duration)`, host numpy arrays out - wall-clock timed, with the
host->device and device->host copies inside the timed region.

The reference arm times the reference's own sw_domain_openmp.c + quantity.c (oracle/_ref, compiled
from /root/reference where it lies) under the numpy restatement of its Python time loop
(oracle/driver.py) on ALL host cores, on the same mesh and fields as one GPU's share of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


IS_REFERENCE_ARM = ("reference" in sys.argv[1:] and "--impl" in sys.argv[1:]) or "--impl=reference" in sys.argv[1:]
# The CPU legs use every host core.  libgomp reads OMP_NUM_THREADS when it is loaded, and torchrun
# exports OMP_NUM_THREADS=1 to its workers, so the reference arm overrides it before anything loads.
if IS_REFERENCE_ARM:
    os.environ["OMP_NUM_THREADS"] = str(host_cores())
    os.environ["SWK_NO_NATIVE_SETUP"] = "1"      # mesh set-up in numpy only: the arm never maps libswk.so
elif int(os.environ.get("WORLD_SIZE", "1")) == 1:
    os.environ["OMP_NUM_THREADS"] = str(host_cores())
os.environ.setdefault("OMP_PROC_BIND", "close")

import numpy as np  # noqa: E402

ALG_BYTES = {"extrapolate": 204.0, "flux": 260.0, "update": 92.0, "flux_update": 268.0}
STEP_BYTES = {"DE0": 556.0, "DE1": 1076.0, "DE2": 1572.0}
METRIC = "triangle-steps/sec (DE1, FP64)"


def log(*a):
    print("[bench r%s]" % os.environ.get("RANK", "0"), *a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as fh:
                return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of one GPU sampled DURING the timed region through NVML inside this
    process (a few microseconds per query).  No nvidia-smi child: starting one initialises NVML for
    every GPU of the box, which stalls kernel launches of all ranks for tens of milliseconds."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index=0, period=0.02):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.period = period
        self.sm, self.reasons, self.power = [], set(), []
        self._stop_evt = threading.Event()
        self.h = None
        self.max_mhz = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; CUDA_VISIBLE_DEVICES may renumber them
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:      # no NVML: one nvidia-smi query before / after instead
            self.h = None
            self.err = repr(e)

    def sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in self.REASONS:
            if r & bit:
                self.reasons.add(name)
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass

    def run(self):
        if self.h is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.sample()
            except Exception:
                break
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2.0)
        if self.h is None or not self.sm:
            # no NVML binding: one nvidia-smi query right AFTER the timed region (never next to it, see above)
            out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0,
                   "note": "NVML unavailable (%s): nvidia-smi queried once right after the timed region"
                           % getattr(self, "err", "no samples")}
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                line = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                       "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                      timeout=20).stdout.strip().splitlines()[0]
                f = [x.strip() for x in line.split(",")]
                out.update(sm_mhz=float(f[0]), sm_max_mhz=float(f[1]), samples=1,
                           reasons=[n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                       "sw_power_cap"), f[2:6]) if v.lower().startswith("active")])
            except Exception:
                pass
            return out
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "sm_mhz_min": float(min(self.sm)),
                "power_w_max": (max(self.power) if self.power else None),
                "how": "NVML in-process, every %.0f ms during the timed region" % (self.period * 1e3)}


# ------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------
def workload_config(a, world=1):
    tri = 4 * a.size * a.size
    if a.config == "sweep":
        w = ("rectangular_cross %dx%d per GPU (%d triangles/GPU), DE1 rk2 FP64, Reflective, Manning 0.03, "
             "rain Rate_operator 1e-4 (BASELINE.json configs[2]; --gpus 8 = configs[3])" % (a.size, a.size, tri))
    elif a.config == "tsunami":
        w = ("rectangular_cross %dx%d (%d triangles), sloping beach + island, DE1 rk2 FP64, Manning 0.025, "
             "left: Transmissive_n_momentum_zero_t_momentum_set_stage(0.5 sin(2 pi t/60)), right: Transmissive, "
             "top/bottom: Reflective (BASELINE.json configs[1])" % (a.size, a.size, tri))
    else:
        tri = 2 * a.size * a.size * 4
        w = ("rectangular_cross %dx%d (%d triangles), DE1 with rk3 FP64, Inlet_operator Q=100 + Boyd_box_operator "
             "across an embankment (BASELINE.json configs[4])" % (2 * a.size, a.size, tri))
    if a.config != "sweep":
        tri = tri // world            # one global mesh shared out over the ranks (strong scaling)
    # bytes a step streams from resident state (cq, eq, xg, fg, conn, eu, bk, eta ~ 450 B/triangle)
    state_gb = 450.0 * tri / 1e9
    if state_gb > 4 * 0.126:
        l2 = "inputs larger than L2 (%.1f GB of state per GPU vs 126 MB L2), no flush" % state_gb
    else:
        l2 = ("NOT VALID AS A BENCH LINE: %.2f GB of state per GPU is within reach of the 126 MB L2 and no flush is "
              "done (size chosen for debugging)" % state_gb)
    return {"workload": w, "triangles_per_gpu": tri, "l2_policy": l2}


def build_domain(a, rank=0, nranks=1, device=0, comm=None):
    from anuga_core_b200 import workloads
    if a.config == "sweep":
        if nranks == 1:
            return workloads.roofline_sweep_domain(a.size, a.size, alg="DE1", rain=1.0e-4, device=device)
        from anuga_core_b200 import parallel
        m, n = parallel.weak_scaling_shape(a.size, nranks)
        return parallel.strip_partitioned_sweep_domain(m, n, rank, nranks, device=device)
    if a.config == "structures":
        # fixed total size shared out over the ranks (strong scaling): the rank's strip of the mesh, the
        # structures created on the distributed domain with their global geometry (as the reference's parallel
        # scripts do after distribute)
        return workloads.structures_domain(2 * a.size, a.size, rank=rank, nranks=nranks, comm=comm, device=device)
    if nranks == 1:
        return workloads.tsunami_domain(a.size, a.size, device=device)
    # tsunami at N > 1: every rank builds the sequential domain and cuts out its own part (replicated build)
    from anuga_core_b200 import parallel
    g = workloads.tsunami_domain(a.size, a.size)
    N = g.number_of_triangles
    return parallel.distribute(g, nranks, epart=(np.arange(N) * nranks) // N, ranks=[rank],
                               domain_kw=dict(device=device))[rank]


def alg_of(a):
    return "DE2" if a.config == "structures" else "DE1"


# ------------------------------------------------------------------------------------------
# the reference's CPU implementation (reference arm and cpu_baseline)
# ------------------------------------------------------------------------------------------
def cpu_reference_run(a, size, steps, warmup):
    """The reference's own C/OpenMP kernels (oracle/_ref, compiled from /root/reference) under the
    numpy restatement of its Python time loop, on the host cores.  Falls back to the C port."""
    from anuga_core_b200 import workloads
    from oracle.driver import OracleDomain, LIBS
    if os.path.exists(LIBS["ref_fma"]):
        backend, kind = "ref_fma", "reference"
    elif os.path.exists(LIBS["ref"]):
        backend, kind = "ref", "reference"
    else:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
        backend, kind = "port", "port"
    t_setup = time.time()
    if a.config == "tsunami":
        d = workloads.tsunami_domain(size, size)
    elif a.config == "structures":
        d = workloads.structures_domain(2 * size, size, with_structures=False)
    else:
        d = workloads.roofline_sweep_domain(size, size, alg="DE1", rain=1.0e-4)
    o = OracleDomain(workloads.domain_to_scenario(d), backend=backend)
    N = d.number_of_triangles
    del d
    o.relative_finaltime = None
    o.relative_yieldtime = 1.0e300
    o.distribute_to_vertices_and_edges()
    step = {"rk2": o.evolve_one_rk2_step, "rk3": o.evolve_one_rk3_step, "euler": o.evolve_one_euler_step}[
        o.timestepping_method]

    def run(k):
        for _ in range(k):
            t0 = o.relative_time
            step(None, None)
            o.apply_fractional_steps()
            o.relative_time = t0 + o.timestep
    log("reference arm: %d triangles, set-up %.1f s, backend %s, %s threads" % (
        N, time.time() - t_setup, backend, os.environ.get("OMP_NUM_THREADS")))
    run(warmup)
    t0 = time.perf_counter()
    run(steps)
    dt = time.perf_counter() - t0
    cores = 1 if backend == "port" else int(os.environ.get("OMP_NUM_THREADS", "1"))
    what = ("reference sw_domain_openmp.c + quantity.c, -O3 -march=x86-64-v3 -fopenmp (multiprocessor_mode 2)"
            if kind == "reference" else "serial C port")
    return {"value": N * steps / dt, "unit": "triangle-steps/s", "cores": cores, "kind": kind,
            "sample": "%s: %d triangles (%s cells per side), %d timed steps after %d warm-up steps, %d OpenMP threads; %s"
                      % (a.config, N, size, steps, warmup, cores, what),
            "ms_per_step": dt / steps * 1e3, "triangles": N}


def reference_arm(a, config):
    size = a.cpu_size or a.size
    N = 4 * size * size * (2 if a.config == "structures" else 1)
    # bounded: keep the timed region within ~150 s of host time (about 1e7 triangle-steps/s on 16 cores)
    est_step = N / 8.0e6
    steps = max(1, min(a.steps, int(150.0 / est_step) or 1))
    warmup = max(1, min(a.warmup, int(40.0 / est_step) or 1))
    r = cpu_reference_run(a, size, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "triangle-steps/s",
            "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "triangle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "one GPU's share of the workload (%d triangles) on the host cores; the CPU rate per triangle "
                    "does not depend on how many such shares a multi-GPU job holds" % r["triangles"]}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# multi-GPU parity check (runs before the timed region at N > 1)
# ------------------------------------------------------------------------------------------
def multi_gpu_parity_check(comm, device):
    """A small wet/dry dam break (DE1, then DE2) distributed over all ranks must reproduce the
    single-GPU run of the same domain bit for bit (every full triangle, the time and the timestep)."""
    from anuga_core_b200 import workloads, parallel
    out = {"ranks": comm.size, "cases": []}
    bad_total = 0
    for alg, m, n in (("DE1", 48, 40), ("DE2", 40, 32)):
        g = workloads.dam_break_domain(m, n, alg=alg, device=device)
        N = g.number_of_triangles
        epart = (np.arange(N) * comm.size) // N
        d = parallel.distribute(g, comm.size, epart=epart, ranks=[comm.rank], domain_kw=dict(device=device))[comm.rank]
        d.attach_communicator(comm)
        for _ in d.evolve(yieldstep=0.5, finaltime=1.0):
            pass
        for _ in g.evolve(yieldstep=0.5, finaltime=1.0):
            pass
        nf = d.number_of_full_triangles
        ids = d.tri_l2s[:nf]
        bad = 0
        for name in ("stage", "xmomentum", "ymomentum"):
            bad += int(np.count_nonzero(d.quantities[name].centroid_values[:nf] != g.quantities[name].centroid_values[ids]))
        bad += int(d.total_steps != g.total_steps) + int(d.timestep != g.timestep) + int(d.get_time() != g.get_time())
        bad = int(comm.allreduce_sum(float(bad)))
        bad_total += bad
        out["cases"].append({"case": "dam_break %dx%d %s" % (m, n, alg), "steps": int(g.total_steps),
                             "values_differing_from_1gpu": bad})
        d._release_device()
        g._release_device()
    out["bit_identical"] = bad_total == 0
    return out


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="sweep", choices=["sweep", "tsunami", "structures"])
    ap.add_argument("--size", type=int, default=None, help="cells per side per GPU (sweep: 2000 -> 16M triangles)")
    ap.add_argument("--cpu-size", type=int, default=None,
                    help="cells per side of the CPU runs (default: the reference arm uses --size; the cpu_baseline "
                         "of the GPU arm a bounded 1000-cell sample)")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.size is None:
        a.size = {"sweep": 2000, "tsunami": 1000, "structures": 2000}[a.config]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = workload_config(a, world)

    if a.impl == "reference":
        if rank == 0:
            reference_arm(a, config)
        return 0

    import anuga_core_b200 as ab
    if ab.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 device (there is no CPU fallback)")
    sampler = ClockSampler(local_rank) if rank == 0 else None      # NVML initialised before anything is timed
    comm = None
    t_all = time.time()
    if world > 1:
        from anuga_core_b200 import parallel
        comm = parallel.init_process_group(backend="nccl", device=local_rank)
    parity = None
    if comm is not None and not a.no_parity_check:
        parity = multi_gpu_parity_check(comm, local_rank)
        if rank == 0:
            log("parity check:", parity)
        if not parity["bit_identical"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "multi-GPU run differs from single-GPU run",
                                  "parity_check": parity}))
            return 3
    t_setup = time.time()
    d = build_domain(a, rank, world, device=local_rank, comm=comm)
    t_built = time.time()
    if comm is not None:
        d.attach_communicator(comm)
    # first yield: upload + distribute; leaves the state resident
    it = d.evolve(yieldstep=1.0e9, finaltime=None)
    next(it)
    dev = d._dev
    N_local = d.number_of_full_triangles
    t_up = time.time()
    setup = {"parity_check_s": t_setup - t_all, "build_domain_s": t_built - t_setup, "upload_first_yield_s": t_up - t_built}
    log("set-up", setup)

    def barrier():
        dev.synchronize()
        if comm is not None:
            comm.barrier()

    host_ops = [op for op in d.fractional_step_operators if getattr(op, "host_side", False)]

    def run_steps(k, per_kernel):
        """k timesteps of the device-resident loop; CUDA-event milliseconds"""
        if not host_ops:
            return dev.run_steps(k, per_kernel=per_kernel)
        return d.run_steps_with_host_operators(k)

    run_steps(a.warmup, False)
    barrier()
    per_kernel_in_region = (world == 1) and not host_ops
    launches0 = dev.kernel_launch_count()
    if rank == 0:
        sampler.start()
    barrier()
    ms = run_steps(a.steps, per_kernel_in_region)
    barrier()
    launches = dev.kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if not per_kernel_in_region:
        dev.run_steps(min(a.steps, 50), per_kernel=True)         # same steps again, bracketed kernel by kernel
        barrier()
    ms_local = ms
    if comm is not None:
        ms = comm.allreduce_max(ms)
        N_total = comm.allreduce_sum(N_local)
        ms_min = comm.allreduce_min(ms_local)
    else:
        N_total, ms_min = N_local, ms
    value = N_total * a.steps / (ms * 1e-3)
    ktime = dev.kernel_timing()
    log("timed region: %.3f ms/step (min over ranks %.3f)" % (ms / a.steps, ms_min / a.steps))

    # ---- end to end through the public API: host arrays -> evolve -> host arrays ------------------
    K2 = a.e2e_steps or max(10, min(a.steps, 100))
    q = d.quantities
    dt_now = dev.get_statistics().timestep
    barrier()
    steps_before = dev.get_statistics().total_steps
    t0 = time.perf_counter()
    d.sync_from_host(d.conserved_quantities)       # the host numpy arrays are the input: H2D of stage, x/ymomentum
    for _t in d.evolve(yieldstep=K2 * dt_now, duration=K2 * dt_now):
        pass                                       # the yield leaves the conserved centroid arrays on the host (D2H)
    checksum = float(q["stage"].centroid_values[0] + q["xmomentum"].centroid_values[-1])
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_steps = dev.get_statistics().total_steps - steps_before
    if comm is not None:
        e2e_s = comm.allreduce_max(e2e_s)
    h2d = 3 * 8 * d.number_of_triangles / max(e2e_steps, 1)
    d2h = 3 * 8 * d.number_of_triangles / max(e2e_steps, 1)
    e2e = {"value": N_total * e2e_steps / e2e_s, "unit": "triangle-steps/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps_per_call": int(e2e_steps), "seconds": e2e_s, "checksum": checksum,
           "note": "Domain.sync_from_host (page-locked numpy arrays -> HBM) + `for t in domain.evolve(yieldstep=T, "
                   "duration=T)` (%d timesteps in the device loop, then the yield's download of stage / xmomentum / "
                   "ymomentum into the numpy arrays), wall clock, max over ranks" % e2e_steps}
    if comm is not None:
        comm.barrier()
    if rank != 0:
        return finish(comm)

    peak, peak_src = measured_peak()
    alg = alg_of(a)
    kernels = {}
    k_steps = a.steps if per_kernel_in_region else min(a.steps, 50)        # steps the per-kernel events cover
    nsub = {"DE0": 1, "DE1": 2, "DE2": 3}[alg]
    passes = {"extrapolate": nsub, "flux": 1, "update": 1, "flux_update": nsub - 1}
    for name, (tot, n) in ktime.items():
        if n > 0 and passes[name] > 0:
            # one PASS over the triangles (with the halo overlap a pass is two launches: halo sources, rest)
            pass_ms = tot / k_steps / passes[name]
            kernels[name] = {"avg_ms": pass_ms, "launches": int(n), "launches_per_pass": n / (k_steps * passes[name]),
                             "share_of_step": (tot / k_steps) / (ms / a.steps),
                             "achieved_gbs": ALG_BYTES[name] * n_active_triangles(d, name) / (pass_ms * 1e-3) / 1e9}
    roofline = None
    if kernels:
        dom = max(kernels, key=lambda k: kernels[k]["share_of_step"])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and a.config == "sweep" and a.size == 2000:
            try:
                with open(tpath) as fh:
                    traffic = json.load(fh).get(dom)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[dom]["achieved_gbs"] / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_triangle": ALG_BYTES[dom],
                    "measured_in": ("the timed region (CUDA events around every launch)" if per_kernel_in_region else
                                    "a second pass of the same steps right after the timed region (the timed region "
                                    "replays CUDA graphs - NCCL calls included - or runs host-side operators between "
                                    "steps and cannot carry per-kernel events)"),
                    "whole_step": {"algorithmic_bytes_per_triangle_step": STEP_BYTES[alg],
                                   "achieved_gbs": value / world * STEP_BYTES[alg] / 1e9,
                                   "frac_of_peak": value / world * STEP_BYTES[alg] / 1e9 / peak,
                                   "frac_of_8TBs": value / world * STEP_BYTES[alg] / 8.0e12},
                    # halo pack / unpack, NCCL send / recv / allreduce waits, clock kernels and launch gaps: what the
                    # step spends outside the four hot kernels (at N > 1 mostly waiting for the slowest rank)
                    "outside_hot_kernels_ms_per_step": ms / a.steps - sum(t for t, n in ktime.values() if n > 0) / k_steps,
                    "kernels": kernels}
    line = {"metric": METRIC if alg == "DE1" else "triangle-steps/sec (%s, FP64)" % alg, "value": value,
            "unit": "triangle-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak" if a.config == "sweep" else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "parallelism": ("1 process/GPU, strip partition, NCCL halo send/recv + uint64 min-allreduce captured in "
                            "the step's CUDA graph" if world > 1 else "single GPU"),
            "setup_seconds": setup, "ms_per_step_min_over_ranks": ms_min / a.steps,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
    if parity is not None:
        line["parity_check"] = parity
    if world == 1 and not a.no_cpu_baseline:
        cs = a.cpu_size or min(a.size, 1000)
        r = cpu_reference_run(a, cs, a.cpu_steps, 2)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    return finish(comm)


def finish(comm):
    """End of a GPU-arm process: last barrier, flush, and leave without running destructors (device handles
    and the NCCL communicator go with the process; tearing them down rank by rank can block on the others)."""
    if comm is not None:
        comm.finalize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def n_active_triangles(d, kernel):
    """triangles one launch of `kernel` processes: pass A covers ghosts too, the flux / update kernels
    only the full triangles of a sub-domain"""
    if kernel == "extrapolate" or d.numproc == 1:
        return d.number_of_triangles
    return d.number_of_full_triangles


if __name__ == "__main__":
    sys.exit(main())
