"""First-contact probe for the GPU box: prints parity numbers instead of asserting.
Run:  python tests/gpu_probe.py   (under gpurun)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import anuga_core_b200 as ab
from oracle.driver import OracleDomain
import scenarios as S


def compare(tag, d, o):
    d.sync_to_host()
    q = d.quantities
    print("%-34s stage %.3e xmom %.3e ymom %.3e | t=%.6f/%.6f" % (
        tag, S.rel_err(q["stage"].centroid_values, o.stage_c), S.rel_err(q["xmomentum"].centroid_values, o.xmom_c),
        S.rel_err(q["ymomentum"].centroid_values, o.ymom_c), d.get_time(), o.get_time()), flush=True)


def passes(maker, name, **kw):
    d = maker(**kw)
    o = OracleDomain(S.domain_to_scenario(d), backend="ref")
    # pass A
    d.distribute_to_vertices_and_edges()
    o.distribute_to_vertices_and_edges()
    q = d.quantities
    print("[%s] extrapolate: stage_e %.3e height_e %.3e xmom_e %.3e ymom_e %.3e bed_e %.3e stage_v %.3e" % (
        name, S.rel_err(q["stage"].edge_values, o.stage_e), S.rel_err(q["height"].edge_values, o.height_e),
        S.rel_err(q["xmomentum"].edge_values, o.xmom_e), S.rel_err(q["ymomentum"].edge_values, o.ymom_e),
        S.rel_err(q["elevation"].edge_values, o.bed_e), S.rel_err(q["stage"].vertex_values, o.stage_v)))
    d.update_boundary()
    o.update_boundary()
    dev = d._dev
    print("[%s] boundary: %.3e %.3e %.3e" % (name, S.rel_err(dev.get_quantity("STAGE_B"), o.stage_b),
                                           S.rel_err(dev.get_quantity("XMOM_B"), o.xmom_b),
                                           S.rel_err(dev.get_quantity("YMOM_B"), o.ymom_b)))
    ft = d.compute_fluxes(0)
    o.compute_fluxes(0)
    print("[%s] flux: dt %.17g vs %.17g | eu %.3e %.3e %.3e | max_speed %.3e | bfs %.3e vs %.3e" % (
        name, ft, o.flux_timestep, S.rel_err(q["stage"].explicit_update, o.stage_eu),
        S.rel_err(q["xmomentum"].explicit_update, o.xmom_eu), S.rel_err(q["ymomentum"].explicit_update, o.ymom_eu),
        S.rel_err(d.get_max_speed(), o.max_speed), dev.get_statistics().boundary_flux_sum[0], o.boundary_flux_sum[0]))
    o.compute_forcing_terms()
    o.timestep = d.timestep = min(d.CFL * ft, 1000.0)
    d.update_conserved_quantities()
    o.update_conserved_quantities()
    compare("[%s] after update" % name, d, o)


def evolve(maker, name, yieldstep, finaltime, backend="ref", **kw):
    d = maker(**kw)
    o = OracleDomain(S.domain_to_scenario(d), backend=backend)
    t0 = time.time()
    steps = 0
    for t in d.evolve(yieldstep=yieldstep, finaltime=finaltime):
        steps += d.number_of_steps
    t1 = time.time()
    for t in o.evolve(yieldstep=yieldstep, finaltime=finaltime):
        pass
    compare("[%s] evolve %d steps (oracle %d) %.2fs" % (name, steps, len(o.timestep_history), t1 - t0), d, o)
    print("      bfi %.6e vs %.6e  fsvi %.6e vs %.6e  launches %d" % (
        d.boundary_flux_integral, o.boundary_flux_integral, d.fractional_step_volume_integral,
        o.fractional_step_volume_integral, d.kernel_launches))


if __name__ == "__main__":
    print("devices:", ab.device_count())
    for alg in ("DE0", "DE1"):
        passes(S.dam_break, "dam %s" % alg, n=10, alg=alg)
    passes(S.wet_dry_beach, "beach DE1", n=12, alg="DE1")
    passes(S.tsunami, "tsunami DE1", n=12, alg="DE1")
    for alg in ("DE0", "DE1", "DE2"):
        evolve(S.dam_break, "dam %s n=20" % alg, 0.5, 2.0, n=20, alg=alg)
    evolve(S.wet_dry_beach, "beach DE1 n=24", 0.5, 3.0, n=24, alg="DE1")
    evolve(S.wet_dry_beach, "beach DE0 n=24", 0.5, 3.0, n=24, alg="DE0")
    evolve(S.smooth_wet, "smooth DE1 rain n=30", 1.0, 4.0, n=30, alg="DE1", rain=1e-4)
    evolve(S.tsunami, "tsunami dirichlet n=24", 1.0, 4.0, n=24, alg="DE1")
    evolve(S.tsunami, "tsunami set_stage n=24", 1.0, 4.0, n=24, alg="DE1", left="set_stage")
    evolve(S.dam_break, "dam DE1 n=100 noreorder", 0.5, 1.0, n=100, alg="DE1", reorder=False)
    evolve(S.dam_break, "dam DE1 n=100", 0.5, 1.0, n=100, alg="DE1")
    # quick throughput look
    for n in (500, 1000):
        d = S.smooth_wet(n=n, alg="DE1", rain=1e-4)
        it = d.evolve(yieldstep=1000.0, finaltime=None)
        next(it)
        dev = d._dev
        r = dev.evolve(1e9, None, 3)
        dev.synchronize()
        t0 = time.time()
        K = 20
        r = dev.evolve(1e9, None, K)
        dev.synchronize()
        dt = time.time() - t0
        N = d.number_of_triangles
        print("n=%d N=%d: %.3f ms/step, %.3e tri-steps/s, roofline frac (1076B @6534.8GB/s) %.3f" % (
            n, N, dt / K * 1e3, N * K / dt, N * K / dt * 1076 / 6534.8e9), flush=True)
