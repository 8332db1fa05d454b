#!/usr/bin/env python
"""Generate the golden fixtures from the UNMODIFIED Python reference.

Run in the build container only (needs the scratch build, oracle/build_pyref.py):
    python oracle/build_pyref.py && python tests/golden/make_golden.py

For every case of cases.py the reference Domain (multiprocessor_mode 2, the OpenMP C backend,
parity build -ffp-contract=off) is evolved and the inputs/outputs are stored in
tests/golden/<case>.npz:
    init_*      centroid arrays after set_quantity (checks this repo's host-side Quantity/Mesh)
    step1_*     conserved centroid values after the first timestep
    final_*     ... at finaltime
    dts         the sequence of timesteps, yields the yield times
    bfi, fsvi   boundary_flux_integral / fractional_step_volume_integral at the end
plus mesh_cross_3x4.npz (mesh arrays) and the embedded 8-digit KAT of the reference's own test.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import pyref  # noqa: E402
import cases  # noqa: E402

os.environ.setdefault("OMP_NUM_THREADS", "4")
anuga = pyref.import_anuga()


def run_case(name):
    builder, ev = cases.CASES[name]
    d = builder(anuga)
    d.set_multiprocessor_mode(2)
    q = d.quantities
    out = {
        "init_stage": q["stage"].centroid_values.copy(), "init_xmom": q["xmomentum"].centroid_values.copy(),
        "init_ymom": q["ymomentum"].centroid_values.copy(), "init_elev": q["elevation"].centroid_values.copy(),
        "init_friction": q["friction"].centroid_values.copy(),
    }
    dts = []
    orig = d.apply_fractional_steps

    def hook():
        orig()
        dts.append(d.timestep)
        if len(dts) == 1:
            out["step1_stage"] = q["stage"].centroid_values.copy()
            out["step1_xmom"] = q["xmomentum"].centroid_values.copy()
            out["step1_ymom"] = q["ymomentum"].centroid_values.copy()
    d.apply_fractional_steps = hook
    yields = []
    for t in d.evolve(**ev):
        yields.append(t)
    out["final_stage"] = q["stage"].centroid_values.copy()
    out["final_xmom"] = q["xmomentum"].centroid_values.copy()
    out["final_ymom"] = q["ymomentum"].centroid_values.copy()
    out["final_stage_edge"] = q["stage"].edge_values.copy()
    out["final_xmom_vertex"] = q["xmomentum"].vertex_values.copy()
    out["dts"] = np.array(dts)
    out["yields"] = np.array(yields)
    out["bfi"] = np.array([d.get_boundary_flux_integral()])
    out["fsvi"] = np.array([d.get_fractional_step_volume_integral()])
    k = 0
    for op in d.fractional_step_operators:      # culverts: resolved geometry + accumulated transfer
        if hasattr(op, "inlets"):
            for j, inlet in enumerate(op.inlets):
                out["struct%d_inlet%d_ids" % (k, j)] = np.asarray(inlet.triangle_indices, dtype=np.int64)
            out["struct%d_enquiry" % k] = np.array([i.enquiry_index for i in op.inlets], dtype=np.int64)
            out["struct%d_length" % k] = np.array([op.culvert_length])
            out["struct%d_accumulated_flow" % k] = np.array([op.accumulated_flow])
            k += 1
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-28s N=%6d steps=%5d t=%.3f  max|stage|=%.6f" % (name, len(d), len(dts), yields[-1],
                                                            np.abs(out["final_stage"]).max()))
    return d, out


def main():
    only = sys.argv[1:]
    if only:                                    # regenerate just the named cases
        for name in only:
            run_case(name)
        return
    for name in cases.CASES:
        d, out = run_case(name)
        if name == "kat_bedslope_more_steps":
            # the reference's own expected values (8 digits) must hold for the build we generate from
            assert np.allclose(out["final_stage"][:6], cases.KAT_BEDSLOPE_W_EX_HEAD), out["final_stage"][:6]
    # mesh fixture
    dm = anuga.rectangular_cross_domain(3, 4, len1=3.0, len2=5.0)
    m = dm.mesh
    np.savez_compressed(os.path.join(HERE, "mesh_cross_3x4.npz"), **{
        k: np.array(getattr(m, k)) for k in ("nodes", "triangles", "areas", "normals", "edgelengths", "radii",
                                             "centroid_coordinates", "vertex_coordinates",
                                             "edge_midpoint_coordinates", "neighbours", "neighbour_edges",
                                             "surrogate_neighbours", "number_of_boundaries", "boundary_cells",
                                             "boundary_edges")})
    from anuga.abstract_2d_finite_volumes.mesh_factory import rectangular
    p, e, b = rectangular(4, 3, 2.0, 1.5)
    dr = anuga.Domain(p, e, b)
    np.savez_compressed(os.path.join(HERE, "mesh_rect_4x3.npz"), **{
        k: np.array(getattr(dr.mesh, k)) for k in ("nodes", "triangles", "areas", "normals", "edgelengths", "radii",
                                                  "centroid_coordinates", "neighbours", "neighbour_edges",
                                                  "surrogate_neighbours", "number_of_boundaries",
                                                  "boundary_cells", "boundary_edges")})
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
