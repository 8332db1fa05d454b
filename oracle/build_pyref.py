#!/usr/bin/env python
"""Build the UNMODIFIED Python reference (ANUGA) in a scratch directory.

TEST INFRASTRUCTURE ONLY.  Nothing in the product imports this.

The reference is a Python package with Cython/C extensions and a meson build
that is not available in this image.  This script copies the package from the
read-only checkout to a scratch directory OUTSIDE the repository (default
/tmp/anuga_pyref), cythonizes the ``*_ext.pyx`` files that ``import anuga``
needs and compiles them in place with setuptools + /usr/bin/gcc.

The resulting tree is used in THIS container only, by
``tests/golden/make_golden.py``, to produce the committed golden fixtures.  It
never travels to the GPU box and is never copied into the repository.

Two builds of ``sw_domain_openmp_ext`` matter:
  * parity build  : -O3 -ffp-contract=off -fopenmp  (no FMA contraction)
  * timing build  : -O3 -march=native -fopenmp      (what meson.build:47 asks)
Select with --native.

Usage:  python oracle/build_pyref.py [--dest DIR] [--native] [--src DIR]
"""
import argparse
import os
import shutil
import subprocess
import sys
import sysconfig

SKIP = ("sparse_matrix_ext", "openacc", "cuda")

EXTRA_SOURCES = {
    "fitsmooth_ext": ["utilities/quad_tree.c", "utilities/sparse_dok.c",
                      "utilities/sparse_csr.c"],
}


def build(src, dest, native):
    import numpy
    from Cython.Build import cythonize  # noqa: F401  (checked early)

    pkg = os.path.join(dest, "anuga")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    os.makedirs(dest, exist_ok=True)
    shutil.copytree(os.path.join(src, "anuga"), pkg,
                    ignore=shutil.ignore_patterns("*.pyc", "__pycache__"))
    # copytree keeps the read-only bits of the checkout
    for root, dirs, files in os.walk(pkg):
        for d in dirs:
            os.chmod(os.path.join(root, d), 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)

    pyx = []
    for root, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith("_ext.pyx") and not any(s in f for s in SKIP):
                pyx.append(os.path.join(root, f))
    pyx.sort()

    setup_py = os.path.join(dest, "_build_setup.py")
    lines = [
        "import numpy, os",
        "from setuptools import setup, Extension",
        "from Cython.Build import cythonize",
        "exts = []",
    ]
    for p in pyx:
        rel = os.path.relpath(p, dest)
        mod = rel[:-4].replace(os.sep, ".")
        name = os.path.basename(p)[:-4]
        d = os.path.dirname(rel)
        srcs = [rel] + [os.path.join("anuga", s) for s in EXTRA_SOURCES.get(name, [])]
        cargs = ["-O3", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", "-w"]
        largs = []
        if "openmp" in name or name == "fitsmooth_ext":
            cargs.append("-fopenmp")
            largs.append("-fopenmp")
        if native and name == "sw_domain_openmp_ext":
            cargs.append("-march=native")
        else:
            cargs.append("-ffp-contract=off")
        lang = "c++" if name == "neighbour_table_ext" else "c"
        lines.append(
            "exts.append(Extension(%r, %r, include_dirs=[numpy.get_include(), "
            "'anuga/utilities', %r], extra_compile_args=%r, extra_link_args=%r, language=%r))"
            % (mod, srcs, d, cargs, largs, lang))
    lines.append("setup(name='anuga_pyref', ext_modules=cythonize(exts, "
                 "compiler_directives={'language_level': 3}, quiet=True), script_args=['build_ext', '--inplace', '-j', '8'])")
    with open(setup_py, "w") as fh:
        fh.write("\n".join(lines) + "\n")

    env = dict(os.environ, CC="/usr/bin/gcc", CXX="/usr/bin/g++",
               LDSHARED="/usr/bin/gcc -shared")
    subprocess.check_call([sys.executable, setup_py], cwd=dest, env=env)
    with open(os.path.join(dest, "BUILD_INFO"), "w") as fh:
        fh.write("native=%s\nsrc=%s\npython=%s\n" % (native, src, sysconfig.get_python_version()))
    print("built python reference in", dest)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dest", default=os.environ.get("ANUGA_PYREF", "/tmp/anuga_pyref"))
    ap.add_argument("--native", action="store_true")
    a = ap.parse_args()
    build(a.src, os.path.abspath(a.dest), a.native)
