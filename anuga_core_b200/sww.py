"""Yield-time output: the SWW file (NetCDF-3, 64-bit offset) the reference writes at every yield.

Mirrors anuga/file/sww.py: SWW_file.__init__ / store_connectivity / store_timestep (:71-434) and
Write_sww.store_header / store_triangulation / write_dynamic_quantities / store_static_quantities /
store_quantities (:563-1100), with the georeference attributes of
anuga/coordinate_transforms/geo_reference.py:169-185.  Same dimension, variable and attribute names,
types and values, so files are interchangeable with the reference's (tests compare them with files
the reference's own writer produced in the build container).

The reference writes through netCDF4 with format NETCDF3_64BIT (anuga/file/netcdf.py:41-47); that
package is not in this image, scipy.io.netcdf_file writes the same on-disk format.  SURVEY.md
section 8(f) row 3.  A distributed run writes one file per rank (name_P<n>_<rank>.sww, with the
local-to-global maps) and merges them afterwards: anuga_core_b200/sww_merge.py.
"""
import os

import numpy as np

max_float = 1.0e36                      # anuga/config.py:13
default_minimum_storable_height = 1.0e-3    # anuga/config.py:189
default_institution = "Geosciences Australia"
RANGE = "_range"


def _open(filename, mode):
    try:
        from scipy.io import netcdf_file
    except ImportError as e:                                    # pragma: no cover
        raise RuntimeError("set_store(True) needs scipy (scipy.io.netcdf_file) to write SWW files") from e
    return netcdf_file(filename, mode, mmap=False, version=2)


def write_header(fid, starttime, number_of_volumes, number_of_nodes, smoothing, order, static_quantities,
                 dynamic_quantities, static_c_quantities, dynamic_c_quantities,
                 description="Output from anuga.file.sww suitable for plotting",
                 institution=default_institution, timezone="UTC", precision="f"):
    """Write_sww.store_header + write_dynamic_quantities (sww.py:563-693, 776-803)"""
    npoints = number_of_nodes if smoothing else 3 * number_of_volumes
    fid.institution = institution
    fid.description = description
    fid.smoothing = "Yes" if smoothing else "No"
    fid.vertices_are_stored_uniquely = "False" if smoothing else "True"
    fid.order = np.int32(order)
    fid.revision_number = "anuga_core_b200"
    fid.revision_date = "None"
    fid.anuga_version = "anuga_core_b200"
    fid.starttime = starttime
    fid.timezone = timezone
    fid.createDimension("number_of_timesteps", None)         # the record dimension (first for scipy)
    fid.createDimension("number_of_volumes", number_of_volumes)
    fid.createDimension("number_of_triangle_vertices", number_of_nodes)
    fid.createDimension("number_of_vertices", 3)
    fid.createDimension("numbers_in_range", 2)
    fid.createDimension("number_of_points", npoints)
    fid.createVariable("x", precision, ("number_of_points",))
    fid.createVariable("y", precision, ("number_of_points",))
    fid.createVariable("volumes", "i", ("number_of_volumes", "number_of_vertices"))
    for q in static_quantities:
        fid.createVariable(q, precision, ("number_of_points",))
        r = fid.createVariable(q + RANGE, precision, ("numbers_in_range",))
        r[0] = max_float
        r[1] = -max_float
    for q in static_c_quantities:
        fid.createVariable(q, precision, ("number_of_volumes",))
    for q in dynamic_quantities:
        fid.createVariable(q, precision, ("number_of_timesteps", "number_of_points"))
        r = fid.createVariable(q + RANGE, precision, ("numbers_in_range",))
        r[0] = max_float
        r[1] = -max_float
    for q in dynamic_c_quantities:
        fid.createVariable(q, precision, ("number_of_timesteps", "number_of_volumes"))
    fid.createVariable("time", "d", ("number_of_timesteps",))


def write_georeference(fid, geo=None):
    """Geo_reference.write_NetCDF (geo_reference.py:169-185) with the defaults of Geo_reference()"""
    fid.xllcorner = float(getattr(geo, "xllcorner", 0.0))
    fid.yllcorner = float(getattr(geo, "yllcorner", 0.0))
    fid.zone = np.int32(getattr(geo, "zone", -1))
    fid.hemisphere = str(getattr(geo, "hemisphere", "undefined"))
    fid.false_easting = np.int32(getattr(geo, "false_easting", 500000))
    fid.false_northing = np.int32(getattr(geo, "false_northing", 10000000))
    fid.datum = str(getattr(geo, "datum", "wgs84"))
    fid.projection = str(getattr(geo, "projection", "UTM"))
    fid.units = str(getattr(geo, "units", "m"))


class SWW_file:
    def __init__(self, domain):
        self.domain = domain
        self.precision = "f"                                     # netcdf_float32
        self.filename = os.path.join(domain.get_datadir(), domain.get_name() + ".sww")
        self.store_centroids = bool(getattr(domain, "store_centroids", False))
        self.minimum_storable_height = getattr(domain, "minimum_storable_height", default_minimum_storable_height)
        self.static_quantities, self.dynamic_quantities = [], []
        self.static_c_quantities, self.dynamic_c_quantities = [], []
        for q, flag in domain.quantities_to_be_stored.items():
            assert q in domain.quantities, "Quantity %s is requested to be stored but it does not exist" % q
            assert flag in (1, 2)
            (self.static_quantities if flag == 1 else self.dynamic_quantities).append(q)
            if self.store_centroids:
                (self.static_c_quantities if flag == 1 else self.dynamic_c_quantities).append(q + "_c")
        os.makedirs(domain.get_datadir() or ".", exist_ok=True)
        fid = _open(self.filename, "w")
        self._header(fid)
        fid.close()

    # -- Write_sww.store_header ------------------------------------------------------------
    def _header(self, fid):
        d = self.domain
        write_header(fid, d.starttime, d.number_of_triangles, d.number_of_nodes, bool(d.smooth), d.default_order,
                     self.static_quantities, self.dynamic_quantities, self.static_c_quantities,
                     self.dynamic_c_quantities, institution=getattr(d, "institution", default_institution),
                     timezone=str(getattr(d, "timezone", "UTC")))

    # -- SWW_file.store_connectivity -------------------------------------------------------
    def store_connectivity(self):
        d = self.domain
        fid = _open(self.filename, "a")
        Q = d.quantities["stage"]
        X, Y, _, V = Q.get_vertex_values(xy=True, precision=np.float32)
        write_georeference(fid, getattr(d, "geo_reference", None))
        fid.variables["x"][:] = X
        fid.variables["y"][:] = Y
        fid.variables["volumes"][:] = np.asarray(V, dtype=np.int32).reshape(-1, 3)
        if d.numproc > 1:            # Write_sww.store_parallel_data (:831-873): what sww_merge needs
            fid.number_of_global_triangles = np.int32(d.number_of_global_triangles)
            fid.number_of_global_nodes = np.int32(d.number_of_global_nodes)
            for vname, dim, values in (("tri_l2g", "number_of_volumes", d.tri_l2g),
                                       ("node_l2g", "number_of_triangle_vertices", d.node_l2g),
                                       ("tri_full_flag", "number_of_volumes", d.tri_full_flag)):
                fid.createVariable(vname, "i", (dim,))
                fid.variables[vname][:] = np.asarray(values).astype(np.int32)
        for name in self.static_quantities:
            A, _ = d.quantities[name].get_vertex_values(xy=False, precision=np.float32)
            x = A.astype(np.float32)
            fid.variables[name][:] = x
            fid.variables[name + RANGE][0] = np.min(x)
            fid.variables[name + RANGE][1] = np.max(x)
        for name in self.static_c_quantities:
            fid.variables[name][:] = d.quantities[name[:-2]].centroid_values.astype(np.float32)
        fid.close()

    # -- SWW_file.store_timestep -----------------------------------------------------------
    def store_timestep(self):
        d = self.domain
        fid = _open(self.filename, "a")
        if "stage" in self.dynamic_quantities:
            w, _ = d.quantities["stage"].get_vertex_values(xy=False)
            z, _ = d.quantities["elevation"].get_vertex_values(xy=False)
            storable = np.array(w - z >= self.minimum_storable_height)
        else:
            storable = None
        values = {}
        for name in self.dynamic_quantities:
            A, _ = d.quantities[name].get_vertex_values(xy=False, precision=np.float32)
            if storable is not None:
                if name == "stage":                # dry points show the bed
                    A = np.where(storable, A, z)
                if name in ("xmomentum", "ymomentum"):
                    A = np.where(storable, A, np.zeros(A.size, A.dtype))
            values[name] = A
        time = d.relative_time
        tvar = fid.variables["time"]
        slice_index = int(tvar.shape[0])
        if slice_index > 0 and time <= tvar[slice_index - 1]:
            slice_index = int(np.where(np.abs(tvar[:] - time) < 1.0e-14)[0][0])
        tvar[slice_index] = time
        for name in self.dynamic_quantities:
            q_values = np.asarray(values[name])
            fid.variables[name][slice_index] = q_values.astype(np.float32)
            rng = fid.variables[name + RANGE]
            lo, hi = np.min(q_values), np.max(q_values)
            if lo < rng[0]:
                rng[0] = lo
            if hi > rng[1]:
                rng[1] = hi
        for name in self.dynamic_c_quantities:
            fid.variables[name][slice_index] = d.quantities[name[:-2]].centroid_values.astype(np.float32)
        fid.close()
        return slice_index
