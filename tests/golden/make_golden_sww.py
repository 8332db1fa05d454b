#!/usr/bin/env python
"""SWW golden fixtures written by the REFERENCE's own writer (anuga/file/sww.py), run in the build
container through the scipy-backed netCDF4 stand-in of oracle/pyref.py (NetCDF-3 64-bit offset is
the format the reference asks for).  Stores every variable, dimension and attribute of the files in
tests/golden/sww_*.npz:
    sww_static   a domain that is stored without evolving (two frames; host arrays only)
    sww_evolve   cases.beach_de1 (n=10) evolved with set_store(True): the yield-time output path
    sww_merged   the reference's sww_merge_parallel applied to three per-rank files (written by this
                 repository's writer for a distributed copy of sww_static)
usage: python oracle/build_pyref.py && python tests/golden/make_golden_sww.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import pyref  # noqa: E402
import cases  # noqa: E402
import sww_cases  # noqa: E402

os.environ.setdefault("OMP_NUM_THREADS", "4")
anuga = pyref.import_anuga()


def dump(path, out_name):
    from scipy.io import netcdf_file
    f = netcdf_file(path, "r", mmap=False)
    out = {}
    for k, v in f.variables.items():
        out["var_" + k] = np.array(v[:])
        out["dims_" + k] = np.array(list(v.dimensions))
    for k, n in f.dimensions.items():
        out["dim_" + k] = np.array([-1 if n is None else n])
    for k in f._attributes:
        a = getattr(f, k)
        out["att_" + k] = np.array(a.decode() if isinstance(a, bytes) else a)
    f.close()
    np.savez_compressed(os.path.join(HERE, out_name + ".npz"), **out)
    print(out_name, sorted(k for k in out if k.startswith("var_")))


def main():
    tmp = tempfile.mkdtemp()
    d = sww_cases.static_domain(anuga, tmp, "ref_static")
    sww_cases.store_two_frames(d)
    dump(os.path.join(tmp, "ref_static.sww"), "sww_static")

    # merge: per-rank files written by THIS repository's writer, merged by the reference's sww_merge
    import anuga_core_b200 as ab
    from anuga_core_b200 import parallel as P
    from anuga.utilities.sww_merge import sww_merge_parallel
    sww_cases.distributed_static_files(ab, P, tmp, "dist_static", nparts=3)
    sww_merge_parallel(os.path.join(tmp, "dist_static"), 3, verbose=False, delete_old=False)
    dump(os.path.join(tmp, "dist_static.sww"), "sww_merged")

    d = sww_cases.evolve_domain(anuga, cases, tmp, "ref_evolve")
    for t in d.evolve(**sww_cases.EVOLVE):
        pass
    dump(os.path.join(tmp, "ref_evolve.sww"), "sww_evolve")


if __name__ == "__main__":
    main()
