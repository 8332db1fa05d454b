#!/usr/bin/env python
"""bench.py - triangle-steps/s of the DE1 (FP64) shallow-water timestep on B200.

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's own C/OpenMP
                                                            code on the host cores)

Workload (BASELINE.json configs[2]): synthetic
rectangular_cross 2000x2000 = 16,000,000 triangles per GPU, smooth everywhere-wet
fields, DE1 (rk2), Reflective boundaries, Manning 0.03, scalar rain Rate_operator.
A "step" is one full DE1 timestep (two flux evaluations).  Weak scaling: every rank owns
a 16M-triangle strip, so --gpus 8 is the 8000x4000... = 128M-triangle mesh of configs[3].

One JSON line on stdout (rank 0).  `value` is device-timed (CUDA events on the library's
stream) with the state resident in HBM; `e2e` goes through Domain.evolve with host numpy
arrays (upload, K steps, download) and is wall-clock timed.
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

# the reference arm / cpu baseline use every host core (read by libgomp at load time)
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
os.environ.setdefault("OMP_PROC_BIND", "close")

import numpy as np  # noqa: E402

ALG_BYTES = {"extrapolate": 204.0, "flux": 260.0, "update": 92.0, "flux_update": 268.0}
STEP_BYTES = {"DE0": 556.0, "DE1": 1076.0, "DE2": 1572.0}


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
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append(line.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2.0)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_domain(size, rank=0, nranks=1, device=0):
    from anuga_core_b200 import workloads
    if nranks == 1:
        return workloads.roofline_sweep_domain(size, size, alg="DE1", rain=1.0e-4, device=device)
    from anuga_core_b200 import parallel
    m, n = parallel.weak_scaling_shape(size, nranks)
    return parallel.strip_partitioned_sweep_domain(m, n, rank, nranks, device=device)


def cpu_reference_run(size, steps, warmup, kind_pref="reference"):
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
    d = workloads.roofline_sweep_domain(size, size, alg="DE1", rain=1.0e-4)
    o = OracleDomain(workloads.domain_to_scenario(d), backend=backend)
    o.relative_finaltime = None
    o.relative_yieldtime = 1.0e300
    o.distribute_to_vertices_and_edges()

    def run(k):
        for _ in range(k):
            t0 = o.relative_time
            o.evolve_one_rk2_step(None, None)
            o.apply_fractional_steps()
            o.relative_time = t0 + o.timestep
    run(warmup)
    t0 = time.perf_counter()
    run(steps)
    dt = time.perf_counter() - t0
    N = d.number_of_triangles
    cores = 1 if backend == "port" else int(os.environ.get("OMP_NUM_THREADS", "1"))
    return {"value": N * steps / dt, "unit": "triangle-steps/s", "cores": cores, "kind": kind,
            "sample": "rectangular_cross %dx%d (%d triangles), DE1, %d steps after %d warm-up, %s"
                      % (size, size, N, steps, warmup,
                         "reference sw_domain_openmp.c -O3 -march=x86-64-v3 -fopenmp + quantity.c"
                         if kind == "reference" else "serial C port"),
            "ms_per_step": dt / steps * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=2000, help="cells per side per GPU (2000 -> 16M triangles)")
    ap.add_argument("--cpu-size", type=int, default=500, help="cells per side of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=100)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "rectangular_cross %dx%d per GPU (%d triangles/GPU), DE1 rk2 FP64, Reflective, "
                          "Manning 0.03, rain Rate_operator 1e-4 (BASELINE.json configs[2]; --gpus 8 = configs[3])"
                          % (a.size, a.size, 4 * a.size * a.size),
              "triangles_per_gpu": 4 * a.size * a.size,
              "l2_policy": "inputs larger than L2 (>= 7 GB of state per GPU vs 126 MB L2), no flush"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(a.steps, a.cpu_steps))
        r = cpu_reference_run(a.cpu_size, steps, max(1, min(a.warmup, 2)))
        line = {"impl": "reference", "metric": "triangle-steps/sec (DE1, FP64)", "value": r["value"],
                "unit": "triangle-steps/s", "n_gpus": a.gpus, "steps": steps, "warmup": max(1, min(a.warmup, 2)),
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "triangle-steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import anuga_core_b200 as ab
    if ab.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 device (there is no CPU fallback)")
    comm = None
    if world > 1:
        from anuga_core_b200 import parallel
        comm = parallel.init_process_group()
    t_setup = time.time()
    d = build_domain(a.size, rank, world, device=local_rank)
    if comm is not None:
        d.attach_communicator(comm)
    # first yield: upload + distribute; leaves the state resident
    it = d.evolve(yieldstep=1.0e9, finaltime=None)
    next(it)
    dev = d._dev
    N_local = d.number_of_full_triangles
    setup_s = time.time() - t_setup

    def barrier():
        dev.synchronize()
        if comm is not None:
            comm.barrier()

    dev.run_steps(a.warmup, per_kernel=False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = dev.kernel_launch_count()
    barrier()
    ms = dev.run_steps(a.steps, per_kernel=True)
    barrier()
    launches = dev.kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if comm is not None:
        ms = comm.allreduce_max(ms)
        N_total = comm.allreduce_sum(N_local)
    else:
        N_total = N_local
    value = N_total * a.steps / (ms * 1e-3)
    ktime = dev.kernel_timing()

    # ---- end to end through the public API with host arrays ------------------------------
    q = d.quantities
    K2 = max(1, a.e2e_steps)
    d.sync_to_host()
    barrier()
    t0 = time.perf_counter()
    d.sync_from_host(d.conserved_quantities) # H2D of stage, xmomentum, ymomentum from page-locked numpy arrays
    dev.evolve(1.0e300, None, K2)            # K2 timesteps, clock scalars read back per batch
    d._mark_device_newer()
    d.sync_to_host()                         # D2H of the conserved centroid arrays
    barrier()
    e2e_s = time.perf_counter() - t0
    if comm is not None:
        e2e_s = comm.allreduce_max(e2e_s)
    h2d = 3 * 8 * d.number_of_triangles / K2
    d2h = 3 * 8 * d.number_of_triangles / K2
    e2e = {"value": N_total * K2 / e2e_s, "unit": "triangle-steps/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps_per_call": K2,
           "note": "Domain.sync_from_host + swk_evolve(%d steps, clock scalars read back per batch) + "
                   "sync_to_host, wall clock; numpy arrays page-locked with cudaHostRegister" % K2}

    if rank != 0:
        return 0
    peak, peak_src = measured_peak()
    dom = max(ktime, key=lambda k: ktime[k][0])
    kernels = {}
    for name, (tot, n) in ktime.items():
        if n > 0:
            avg_ms = tot / n
            kernels[name] = {"avg_ms": avg_ms, "launches": int(n), "share_of_step": tot / ms,
                             "achieved_gbs": ALG_BYTES[name] * d.number_of_triangles / (avg_ms * 1e-3) / 1e9}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as fh:
                traffic = json.load(fh).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kernels[dom]["achieved_gbs"] / peak, "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_triangle": ALG_BYTES[dom],
                "whole_step": {"algorithmic_bytes_per_triangle_step": STEP_BYTES["DE1"],
                               "achieved_gbs": value / world * STEP_BYTES["DE1"] / 1e9,
                               "frac_of_peak": value / world * STEP_BYTES["DE1"] / 1e9 / peak,
                               "frac_of_8TBs": value / world * STEP_BYTES["DE1"] / 8.0e12},
                "kernels": kernels}
    line = {"metric": "triangle-steps/sec (DE1, FP64)", "value": value, "unit": "triangle-steps/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(config, setup_seconds=setup_s,
                                                 parallelism="1 process/GPU, strip partition, NCCL halo + min-allreduce"
                                                 if world > 1 else "single GPU"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
    if world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_run(a.cpu_size, a.cpu_steps, 2)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
