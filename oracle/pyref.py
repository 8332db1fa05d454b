"""Import shim for the scratch build of the Python reference (see build_pyref.py).

TEST INFRASTRUCTURE ONLY - used in the build container to generate golden
fixtures and by CPU tests that are skipped when the scratch build is absent.
The product never imports this module.

``import_anuga()`` registers stub modules for the third-party packages that
the reference imports at module scope but that are not installed in this image
 and returns the ``anuga`` package from the scratch
tree.
"""
import os
import sys
import types

_REPO_COPY = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
# baseline/_ref (git-ignored, built by __graft_entry__.build() from /root/reference) travels to the GPU box
PYREF_DIR = os.environ.get("ANUGA_PYREF") or (_REPO_COPY if os.path.isdir(os.path.join(_REPO_COPY, "anuga"))
                                               else "/tmp/anuga_pyref")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.tri", "matplotlib.cm",
    "matplotlib.colors", "matplotlib.animation", "matplotlib.figure",
    "matplotlib.backends", "matplotlib.backends.backend_agg",
    "netCDF4", "utm", "pyproj", "affine", "meshpy", "meshpy.triangle",
    "pymetis", "osgeo", "osgeo.gdal", "osgeo.osr", "openpyxl", "xarray", "git",
    "tomli",
]


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed third-party module %s was called" % self.__name__)


def _netcdf3_dataset():
    """netCDF4.Dataset stand-in for the reference's SWW writer (anuga/file/netcdf.py:41-47 asks for
    format NETCDF3_64BIT, which scipy.io.netcdf_file reads and writes): enough of the netCDF4 API for
    anuga/file/sww.py to run unmodified in this container, so that SWW files written by the reference
    can be compared with the ones this repository writes."""
    from scipy.io import netcdf_file
    from scipy.io import _netcdf
    if not hasattr(_netcdf.netcdf_variable, "__len__"):     # netCDF4 variables have a length
        _netcdf.netcdf_variable.__len__ = lambda self: int(self.shape[0])

    class Dataset(netcdf_file):
        def __init__(self, filename, mode="r", format=None, **kw):
            netcdf_file.__init__(self, filename, mode if mode != "wl" else "w", mmap=False, version=2)
            for k, v in list(self._attributes.items()):     # netCDF4 hands out str, scipy bytes
                if isinstance(v, bytes):
                    self.__dict__[k] = v.decode()

        def createDimension(self, name, length):
            if length in (0, None):                 # netCDF4: size 0 / None = the record dimension
                self.dimensions[name] = None
                self._dims.insert(0, name)          # scipy wants it first in the header
            else:
                netcdf_file.createDimension(self, name, int(length))
    return Dataset


def available():
    return os.path.isdir(os.path.join(PYREF_DIR, "anuga"))


def import_anuga(epart_fn=None):
    """Return the reference ``anuga`` package.

    epart_fn(nparts, adjacency) -> list  is installed as ``pymetis.part_graph``
    so that partition tests can inject a deterministic element partition
    (pymetis itself is absent).
    """
    if not available():
        raise ImportError("python reference not built: run oracle/build_pyref.py")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    if isinstance(sys.modules.get("netCDF4"), _Stub):
        sys.modules["netCDF4"].Dataset = _netcdf3_dataset()
    if epart_fn is not None:
        def part_graph(nparts, adjacency=None, **kw):
            return 0, list(epart_fn(nparts, adjacency))
        sys.modules["pymetis"].part_graph = part_graph
    if PYREF_DIR not in sys.path:
        sys.path.insert(0, PYREF_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import anuga
    return anuga
