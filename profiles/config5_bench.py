"""BASELINE.json configs[4]: rk3 (DE2) on rectangular_cross 4000x2000 = 32M
triangles with an Inlet_operator (Q = 100 m^3/s over a line) and one Boyd_box_operator through an
embankment.  Prints triangle-steps/s of the evolve loop with the structures (host-stepped: their scalar
hydraulics run on the host every step, on gathered inlet cells) and without them (device-resident loop).
usage: python profiles/config5_bench.py [m=4000] [n=2000] [steps=30]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import anuga_core_b200 as A


def build(m, n, structures):
    d = A.rectangular_cross_domain(m, n, len1=float(m), len2=float(n))
    d.set_flow_algorithm("DE2")
    d.set_store(False)
    L = float(m)
    xe = 0.5 * L                                     # embankment across the domain at x = L/2
    d.set_quantity("elevation", lambda x, y: 2.0 * np.exp(-((x - xe) / 6.0) ** 2) + 0.001 * (L - x) / L)
    d.set_quantity("stage", lambda x, y: np.where(x < xe, 1.2, 0.4), location="centroids")
    d.set_quantity("friction", 0.03)
    Br = A.Reflective_boundary(d)
    d.set_boundary({t: Br for t in d.get_boundary_tags()})
    if structures:
        y0 = 0.5 * n
        A.Inlet_operator(d, A.Region(d, line=[[0.1 * L + 0.3, y0 - 20.2], [0.1 * L + 0.3, y0 + 20.3]]), Q=100.0)
        A.Boyd_box_operator(d, losses=1.5, width=3.0, height=1.5,
                            end_points=[[xe - 15.1, y0 + 0.3], [xe + 15.1, y0 + 0.3]],
                            apron=2.55, enquiry_gap=1.4, manning=0.013)
    return d


def run(m, n, steps, structures, dt_estimate=None):
    t0 = time.time()
    d = build(m, n, structures)
    N = d.number_of_triangles
    if structures:
        # the public evolve loop: one device-resident step, then the host-side operators, per timestep;
        # the timed call includes the download of the centroid arrays the generator protocol requires
        for t in d.evolve(yieldstep=1.0e9, duration=3.5 * dt_estimate):
            pass
        setup = time.time() - t0
        before = d.total_steps
        d._dev.synchronize()
        t = time.time()
        for tt in d.evolve(yieldstep=1.0e9, duration=(steps - 0.5) * d.timestep):
            pass
        d._dev.synchronize()
        sec = time.time() - t
        steps = d.total_steps - before
        ops = d.fractional_step_operators
        extra = dict(inlet_triangles=int(len(ops[0].inlet.triangle_indices)),
                     culvert_Q=float(ops[1].discharge), culvert_case=str(ops[1].case),
                     culvert_accumulated_flow=float(ops[1].accumulated_flow))
    else:
        it = d.evolve(yieldstep=1.0e9, finaltime=None)
        next(it)
        setup = time.time() - t0
        d._dev.run_steps(3)
        ms = d._dev.run_steps(steps)
        sec = ms * 1e-3
        r = d._dev.get_statistics()
        extra = dict(timestep=float(r.timestep))
    return dict(structures=structures, triangles=N, steps=steps, ms_per_step=sec / steps * 1e3,
                triangle_steps_per_s=N * steps / sec, setup_seconds=setup, **extra)


if __name__ == "__main__":
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    plain = run(m, n, steps, False)
    print(json.dumps(plain), flush=True)
    print(json.dumps(run(m, n, steps, True, dt_estimate=plain["timestep"])), flush=True)
