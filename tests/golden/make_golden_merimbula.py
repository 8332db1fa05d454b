#!/usr/bin/env python
"""Golden run of the reference on a real unstructured mesh: the Merimbula lake model of the reference's
examples/parallel/run_parallel_merimbula.py (mesh examples/parallel/data/merimbula_10785_1.tsh, 10 785
triangles).  Stores the mesh arrays (so that the GPU box, which has no reference tree, can rebuild the
domain), the initial quantities and the reference's results in tests/golden/merimbula_de1.npz.
usage: python oracle/build_pyref.py && python tests/golden/make_golden_merimbula.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyref  # noqa: E402

os.environ.setdefault("OMP_NUM_THREADS", "4")
anuga = pyref.import_anuga()
import merimbula_case  # noqa: E402


def main():
    d = merimbula_case.build(anuga, anuga.create_domain_from_file(merimbula_case.TSH))
    d.set_multiprocessor_mode(2)
    q = d.quantities
    keys = sorted(d.boundary.keys())
    out = {"nodes": np.asarray(d.get_nodes()), "triangles": np.asarray(d.get_triangles(), dtype=np.int64),
           "boundary_keys": np.array(keys, dtype=np.int64).reshape(-1, 2),
           "boundary_tags": np.array([d.boundary[k] for k in keys]),
           "elevation_vertex": q["elevation"].vertex_values.copy()}
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
    yields = [t for t in d.evolve(**merimbula_case.EVOLVE)]
    out.update(final_stage=q["stage"].centroid_values.copy(), final_xmom=q["xmomentum"].centroid_values.copy(),
               final_ymom=q["ymomentum"].centroid_values.copy(), dts=np.array(dts), yields=np.array(yields))
    np.savez_compressed(os.path.join(HERE, "merimbula_de1.npz"), **out)
    print("merimbula_de1: N=%d steps=%d" % (len(d), len(dts)))


if __name__ == "__main__":
    main()
