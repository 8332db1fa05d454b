"""Running scripts written for the reference package.

``install_as_anuga()`` registers this package in ``sys.modules`` under the reference's name together
with alias modules for the sub-module paths that example and validation scripts import from
(``from anuga.structures.boyd_box_operator import Boyd_box_operator``,
``from anuga.abstract_2d_finite_volumes.mesh_factory import rectangular_cross`` ...), so that a
script only needs two extra lines in front of its ``import anuga``.  Names outside the hot-path
scope (mesh generation from regions, plotting and validation utilities, file conversion) are not
provided; importing them fails as it would without the reference installed.

Also here: the few script-level helpers those scripts use next to the classes
(``Polygon_function``, ``read_polygon``, ``inside_polygon``, ``g``, ``indent``).
"""
import sys
import types

import numpy as np

g = 9.8                 # anuga/config.py:45
indent = "    "         # anuga/config.py


def inside_polygon(points, polygon, closed=True, verbose=False):
    """indices of the points inside the polygon (anuga/geometry/polygon.py:560-600)"""
    from .structures import _inside_polygon
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    return np.flatnonzero(_inside_polygon(pts, np.asarray(polygon, dtype=np.float64)))


def read_polygon(filename, delimiter=","):
    """list of [x, y] vertices from a csv file (anuga/geometry/polygon.py:1000-1060)"""
    poly = []
    with open(filename) as f:
        for line in f:
            fields = line.strip().split(delimiter)
            if len(fields) >= 2 and fields[0] != "":
                poly.append([float(fields[0]), float(fields[1])])
    return poly


class Polygon_function:
    """f(x, y) that takes a different value (constant or function of x, y) inside each polygon of
    `regions` = [(polygon, value), ...]; later regions win where they overlap; `default` elsewhere
    (anuga/geometry/polygon_function.py:11-140)"""

    def __init__(self, regions, default=0.0, geo_reference=None):
        assert len(regions) > 0 and len(regions[0]) == 2, \
            "Polygon_function takes a list of pairs (polygon, value)"
        self.default = default
        self.regions = [(np.asarray(p, dtype=np.float64), v) for p, v in regions]

    def __call__(self, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        pts = np.stack([x.reshape(-1), y.reshape(-1)], axis=1)
        z = self.default(x, y) if callable(self.default) else self.default * np.ones(pts.shape[0])
        z = np.array(z, dtype=np.float64).reshape(-1)
        for polygon, value in self.regions:
            ids = inside_polygon(pts, polygon)
            if callable(value):
                z[ids] = value(pts[ids, 0], pts[ids, 1])
            else:
                z[ids] = value
        return z.reshape(x.shape)


def get_pathname_from_package(package):
    """directory of an importable package (anuga/utilities/system_tools.py)"""
    import importlib
    import os
    return os.path.dirname(importlib.import_module(package).__file__)


def ensure_numeric(A, typecode=None):
    """anuga/utilities/numerical_tools.py: a numpy array of the given type (float by default)"""
    if A is None:
        return None
    return np.array(A, dtype=np.float64 if typecode is None else typecode)


def _alias(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def install_as_anuga():
    """make ``import anuga`` (and the common ``from anuga.<module> import <name>`` lines) resolve to
    this package; returns the package"""
    import anuga_core_b200 as pkg
    from . import boundaries, operators, parallel, structures
    from .mesh import rectangular, rectangular_cross
    sys.modules["anuga"] = pkg
    for name in ("anuga.abstract_2d_finite_volumes", "anuga.shallow_water", "anuga.geometry", "anuga.utilities"):
        _alias(name)
    _alias("anuga.utilities.system_tools", get_pathname_from_package=get_pathname_from_package)
    _alias("anuga.utilities.numerical_tools", ensure_numeric=ensure_numeric)
    _alias("anuga.structures", **{k: getattr(structures, k) for k in dir(structures) if not k.startswith("_")})
    _alias("anuga.operators", **{k: getattr(operators, k) for k in dir(operators) if not k.startswith("_")})
    for sub in ("inlet_operator", "inlet", "inlet_enquiry", "structure_operator", "boyd_box_operator",
                "boyd_pipe_operator", "weir_orifice_trapezoid_operator"):
        _alias("anuga.structures." + sub, **sys.modules["anuga.structures"].__dict__)
    for sub in ("rate_operators", "set_stage", "set_quantity", "set_elevation", "set_stage_operator",
                "set_quantity_operator", "set_elevation_operator", "base_operator"):
        _alias("anuga.operators." + sub, **sys.modules["anuga.operators"].__dict__)
    _alias("anuga.abstract_2d_finite_volumes.quantity", Quantity=pkg.Quantity)
    _alias("anuga.abstract_2d_finite_volumes.mesh_factory", rectangular=rectangular,
           rectangular_cross=rectangular_cross)
    _alias("anuga.abstract_2d_finite_volumes.generic_boundary_conditions",
           Dirichlet_boundary=pkg.Dirichlet_boundary, Transmissive_boundary=pkg.Transmissive_boundary,
           Time_boundary=pkg.Time_boundary)
    _alias("anuga.abstract_2d_finite_volumes.region", Region=pkg.Region)
    _alias("anuga.shallow_water.shallow_water_domain", Domain=pkg.Domain)
    _alias("anuga.shallow_water.boundaries",
           **{k: getattr(boundaries, k) for k in dir(boundaries) if k.endswith("_boundary")})
    _alias("anuga.geometry.polygon", inside_polygon=inside_polygon, read_polygon=read_polygon)
    _alias("anuga.geometry.polygon_function", Polygon_function=Polygon_function)
    par = dict(myid=parallel.myid, numprocs=parallel.numprocs, barrier=parallel.barrier,
               finalize=parallel.finalize, distribute=parallel.distribute_collective, pypar_available=True)
    _alias("anuga.parallel", **par)
    _alias("anuga.parallel.parallel_api", **par)
    return pkg
