"""Boundaries driven by a time-space field: File_boundary, Field_boundary, Time_space_boundary.

Reference: anuga/abstract_2d_finite_volumes/generic_boundary_conditions.py:419-700 (Time_space_boundary,
File_boundary), anuga/shallow_water/boundaries.py:993-1090 (Field_boundary), with the machinery underneath
them: abstract_2d_finite_volumes/file_function.py:226-494 (time axis, thinning, time limit, start-time
alignment) and fit_interpolate/interpolate.py:710-1110 (Interpolation_function: spatial interpolation of
every stored frame to the boundary-edge midpoints once, linear interpolation in time at every call).

B200 design: the frames interpolated to the segment's edge midpoints are uploaded ONCE ([frame][edge][3],
resident in HBM); at every flux evaluation the host only sends the time slot {ratio, frame index} through
the boundary value table and the boundary kernel does q = Q0 + ratio*(Q1 - Q0) per edge (kinds
SWK_BC_TIME_SPACE_TABLE / _MEAN_STAGE), inside the same two-half fused step as every other time-dependent
boundary.  A Time_space_boundary (a Python function of t, x, y) uses the same kind with one frame per RK
substep that the host rewrites every step.
"""
import numpy as np

from . import backend as _b
from .boundaries import Boundary, Modeltime_too_early, Modeltime_too_late

NAN = np.inf          # anuga/utilities/numerical_tools.py:15 (the value the reference stores outside the mesh)


# ----------------------------------------------------------------------------------------
# source data
# ----------------------------------------------------------------------------------------
def read_sww_series(filename, quantity_names):
    """x, y, triangles, time, starttime, georeference and the (frames, nodes) arrays of an SWW file"""
    from scipy.io import netcdf_file
    fid = netcdf_file(filename, "r", mmap=False)
    try:
        out = dict(
            starttime=float(np.asarray(fid.starttime).reshape(-1)[0]),
            xllcorner=float(np.asarray(getattr(fid, "xllcorner", 0.0)).reshape(-1)[0]),
            yllcorner=float(np.asarray(getattr(fid, "yllcorner", 0.0)).reshape(-1)[0]),
            time=np.array(fid.variables["time"][:], dtype=np.float64),
            x=np.array(fid.variables["x"][:], dtype=np.float64),
            y=np.array(fid.variables["y"][:], dtype=np.float64),
            triangles=(np.array(fid.variables["volumes"][:], dtype=np.int64) if "volumes" in fid.variables
                       else None),          # STS files (time series at the gauges of a boundary) have no mesh
        )
        for name in quantity_names:
            if name not in fid.variables:
                raise Exception("Quantity %s is missing from file %s" % (name, filename))
            out[name] = np.array(fid.variables[name][:])       # float32 in the file; promoted when used
    finally:
        fid.close()
    return out


def barycentric_weights(vertex_coordinates, triangles, points):
    """For every point the triangle that holds it and its three weights, with the arithmetic of
    calculate_sigma / new_triangle (anuga/utilities/quad_tree.c:26-94): sigma_i = distance from the opposite
    edge along that edge's unit normal, over the same for vertex i.  Returns (triangle or -1, (n, 3))."""
    V = np.asarray(vertex_coordinates, dtype=np.float64)
    T = np.asarray(triangles, dtype=np.int64)
    x1, y1 = V[T[:, 0], 0], V[T[:, 0], 1]
    x2, y2 = V[T[:, 1], 0], V[T[:, 1], 1]
    x3, y3 = V[T[:, 2], 0], V[T[:, 2], 1]

    def unit_normal(xa, ya, xb, yb, xo, yo):
        # normal of the edge a->b, flipped so that it points away from ... the way new_triangle flips it
        ny = xb - xa
        nx = -(yb - ya)
        ln = np.sqrt(nx * nx + ny * ny)
        nx, ny = nx / ln, ny / ln
        flip = (nx * xo + ny * yo) > 0
        return np.where(flip, -nx, nx), np.where(flip, -ny, ny)
    nx3, ny3 = unit_normal(x1, y1, x2, y2, x3 - x2, y3 - y2)
    nx1, ny1 = unit_normal(x2, y2, x3, y3, x1 - x3, y1 - y3)
    nx2, ny2 = unit_normal(x3, y3, x1, y1, x2 - x1, y2 - y1)
    d0 = (x1 - x2) * nx1 + (y1 - y2) * ny1
    d1 = (x2 - x3) * nx2 + (y2 - y3) * ny2
    d2 = (x3 - x1) * nx3 + (y3 - y1) * ny3
    P = np.asarray(points, dtype=np.float64)
    found = np.full(len(P), -1, dtype=np.int64)
    sig = np.zeros((len(P), 3))
    xmin, xmax = np.minimum(np.minimum(x1, x2), x3), np.maximum(np.maximum(x1, x2), x3)
    ymin, ymax = np.minimum(np.minimum(y1, y2), y3), np.maximum(np.maximum(y1, y2), y3)
    eps = 1.0e-12
    for i, (x, y) in enumerate(P):
        cand = np.flatnonzero((x >= xmin - eps) & (x <= xmax + eps) & (y >= ymin - eps) & (y <= ymax + eps))
        if len(cand) == 0:
            continue
        s0 = ((x - x2[cand]) * nx1[cand] + (y - y2[cand]) * ny1[cand]) / d0[cand]
        s1 = ((x - x3[cand]) * nx2[cand] + (y - y3[cand]) * ny2[cand]) / d1[cand]
        s2 = ((x - x1[cand]) * nx3[cand] + (y - y1[cand]) * ny3[cand]) / d2[cand]
        inside = np.flatnonzero((s0 >= -1.0e-10) & (s1 >= -1.0e-10) & (s2 >= -1.0e-10))
        if len(inside) == 0:
            continue
        # the triangle that holds the point most comfortably (a point on a shared edge gets the same
        # value from either side up to rounding)
        j = inside[np.argmax(np.minimum(np.minimum(s0[inside], s1[inside]), s2[inside]))]
        found[i] = cand[j]
        sig[i] = (s0[j], s1[j], s2[j])
    return found, sig


def interpolate_polyline(data, polyline_nodes, gauge_neighbour_id, interpolation_points, rtol=1.0e-6, atol=1.0e-8):
    """Values at points that lie on the polyline through the gauges: linear between a gauge and its neighbour,
    0 for points on no segment (geometry/polygon.py:1069-1130, polygon.c:269-317)."""
    from .cross_section import point_on_line
    data = np.asarray(data, dtype=np.float64)
    nodes = np.asarray(polyline_nodes, dtype=np.float64)
    pts = np.asarray(interpolation_points, dtype=np.float64)
    nb = np.asarray(gauge_neighbour_id, dtype=np.int64)
    assert data.shape[0] == nodes.shape[0], "function value must be specified at every interpolation node"
    assert data.shape[0] > 0, "Must define function value at one or more nodes"
    if nodes.shape[0] == 1:
        raise Exception("Polyline contained only one point. I need more. " + str(data))
    out = np.zeros(len(pts), dtype=np.float64)
    for j in range(nodes.shape[0]):
        k = int(nb[j])
        if k < 0:
            continue
        x0, y0 = float(nodes[j, 0]), float(nodes[j, 1])
        x1, y1 = float(nodes[k, 0]), float(nodes[k, 1])
        segment_len = np.sqrt((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0))
        slope = (data[k] - data[j]) / segment_len
        for i in range(len(pts)):
            x, y = float(pts[i, 0]), float(pts[i, 1])
            if point_on_line(x, y, x0, y0, x1, y1, rtol, atol):
                alpha = np.sqrt((x - x0) * (x - x0) + (y - y0) * (y - y0))
                out[i] = slope * alpha + data[j]
    return out


def gauges_on_boundary(vertex_coordinates, boundary_polygon):
    """Which STS gauges are vertices of the boundary polygon, in the polygon's order, and for each the next gauge
    along the boundary or -1 (file_function.py:378-418).  -> gauge ids, their coordinates (the polygon's),
    neighbour ids."""
    V = np.asarray(vertex_coordinates, dtype=np.float64)
    P = np.asarray(boundary_polygon, dtype=np.float64)
    temp, boundary_id, gauge_id = [], [], []
    for i in range(len(P)):
        close = np.all(np.abs(V - P[i]) <= 1e-4 + 1e-4 * np.abs(P[i]), axis=1)      # numpy.allclose(V[j], P[i])
        hit = np.flatnonzero(close)
        if len(hit):
            temp.append(P[i])
            gauge_id.append(int(hit[0]))
            boundary_id.append(i)
    if len(temp) == 0:
        raise Exception("None of the sts gauges fall on the boundary")
    neighbour = []
    for i in range(len(boundary_id) - 1):
        neighbour.append(i + 1 if boundary_id[i] + 1 == boundary_id[i + 1] else -1)
    neighbour.append(0 if boundary_id[-1] == len(P) - 1 and boundary_id[0] == 0 else -1)
    neighbour = np.asarray(neighbour, dtype=np.int64)
    if int(np.sum(neighbour >= 0)) != len(temp) - 1:
        raise Exception("incorrect number of segments")
    return np.asarray(gauge_id, dtype=np.int64), np.asarray(temp, dtype=np.float64), neighbour


class Interpolation_function:
    """f(t, point_id): frames of nodal values -> values at fixed points, linear in time
    (fit_interpolate/interpolate.py:710-1110, the spatial-with-interpolation-points use), or - without
    vertex_coordinates - f(t): a plain time series with several attributes (TMS files)."""

    def __init__(self, time, quantities, quantity_names, vertex_coordinates=None, triangles=None,
                 interpolation_points=None, time_thinning=1, gauge_neighbour_id=None):
        time = np.asarray(time, dtype=np.float64)
        if not np.all(time[1:] - time[:-1] >= 0):
            raise Exception("Time must be a monotonuosly increasing sequence %s" % time)
        self.time = np.array(time[::time_thinning])
        self.quantity_names = list(quantity_names)
        self.index = 0
        if vertex_coordinates is None:
            self.spatial = False
            self.interpolation_points = None
            self.precomputed_values = {}
            for name in self.quantity_names:
                Q = np.asarray(quantities[name], dtype=np.float64)
                assert Q.ndim == 1, "a time series without spatial information has one value per time"
                self.precomputed_values[name] = np.array(Q[::time_thinning])
            return
        self.interpolation_points = np.asarray(interpolation_points, dtype=np.float64)
        self.spatial = True
        if triangles is None:
            # STS: values along the polyline of gauges (interpolate.py:985-994)
            self.indices_outside_mesh = np.zeros(0, dtype=np.int64)
            self.precomputed_values = {}
            for name in self.quantity_names:
                Q = np.asarray(quantities[name])
                if Q.ndim == 2:
                    Q = np.array(Q[::time_thinning, :])
                frames = Q if Q.ndim == 2 else np.broadcast_to(Q, (len(self.time),) + Q.shape)
                out = np.zeros((len(self.time), len(self.interpolation_points)))
                for i in range(len(self.time)):
                    out[i] = interpolate_polyline(np.asarray(frames[i], dtype=np.float64), vertex_coordinates,
                                                  gauge_neighbour_id, self.interpolation_points)
                self.precomputed_values[name] = out
            return
        tri, sig = barycentric_weights(vertex_coordinates, triangles, self.interpolation_points)
        self.indices_outside_mesh = np.flatnonzero(tri < 0)
        T = np.asarray(triangles, dtype=np.int64)
        ok = tri >= 0
        nodes = T[np.where(ok, tri, 0)]                     # (m, 3) node ids
        # one row of the interpolation matrix per point; the reference's dictionary-of-keys product adds the
        # three terms in the triangle's vertex order, starting from zero (anuga/utilities/sparse.py:125-131)
        nodes_s, sig_s = nodes, sig
        self.precomputed_values = {}
        for name in self.quantity_names:
            Q = np.asarray(quantities[name])
            if Q.ndim == 2:
                Q = np.array(Q[::time_thinning, :])
            frames = Q if Q.ndim == 2 else np.broadcast_to(Q, (len(self.time),) + Q.shape)
            out = np.zeros((len(self.time), len(self.interpolation_points)))
            for i in range(len(self.time)):
                q = np.asarray(frames[i], dtype=np.float64)
                r = 0.0 + sig_s[:, 0] * q[nodes_s[:, 0]]
                r = r + sig_s[:, 1] * q[nodes_s[:, 1]]
                r = r + sig_s[:, 2] * q[nodes_s[:, 2]]
                out[i] = np.where(ok, r, NAN)
            self.precomputed_values[name] = out

    def __call__(self, t, point_id=None, x=None, y=None):
        if self.spatial and point_id is None:
            raise Exception("Either point_id or x and y must be specified")
        index, ratio = time_slot(self, t)
        q = np.zeros(len(self.quantity_names))
        for i, name in enumerate(self.quantity_names):
            Q = self.precomputed_values[name]
            Q0 = Q[index, point_id] if self.spatial else Q[index]
            if ratio > 0:
                Q1 = Q[index + 1, point_id] if self.spatial else Q[index + 1]
                q[i] = Q0 if (Q0 == NAN and Q1 == NAN) else Q0 + ratio * (Q1 - Q0)
            else:
                q[i] = Q0
        if self.spatial or x is None or y is None:
            return q
        try:                                     # a time series asked at many points (e.g. by Wind_stress):
            N = len(x)                           # one constant column per attribute (interpolate.py:1099-1116)
        except TypeError:
            return q
        assert len(y) == N, "x and y must have same length"
        return [col * np.ones(N, dtype=np.float64) for col in q]

    def get_time(self):
        return self.time


def time_slot(F, t):
    """(frame index, ratio) of model time t in the time axis of an interpolation function - this module's or
    the reference's own (an attached reference File_boundary): interpolate.py:1045-1062"""
    msg = "Model time %.16f is not contained in function domain [%.16f:%.16f].\n" % (t, F.time[0], F.time[-1])
    if t < F.time[0]:
        raise Modeltime_too_early(msg)
    if t > F.time[-1]:
        raise Modeltime_too_late(msg)
    while t > F.time[F.index]:
        F.index += 1
    while t < F.time[F.index]:
        F.index -= 1
    if t == F.time[F.index]:
        return F.index, 0.0
    return F.index, (t - F.time[F.index]) / (F.time[F.index + 1] - F.time[F.index])


def file_function(filename, domain=None, quantities=None, interpolation_points=None, time_thinning=1,
                  time_limit=None, verbose=False, use_cache=False, boundary_polygon=None):
    """file_function.py:29-168 + get_netcdf_file_function :226-494 for SWW files: returns the
    Interpolation_function with attribute starttime; moves domain.starttime forward to the file's if the
    file starts later."""
    if filename.endswith(".tms"):
        return _tms_function(filename, domain, quantities, time_thinning, time_limit)
    sts = filename.endswith(".sts")
    if not (filename.endswith(".sww") or sts):
        raise NotImplementedError("file_function: SWW, STS and TMS files are supported (got %s)" % filename)
    if boundary_polygon is not None and not sts:
        raise NotImplementedError("boundary_polygon applies to STS files only")
    if sts and boundary_polygon is None:
        raise Exception("Files of type sts require boundary polygon")
    if interpolation_points is None:
        raise NotImplementedError("file_function on an SWW / STS file needs interpolation_points")
    names = list(quantities) if quantities is not None else ["stage", "xmomentum", "ymomentum"]
    src = read_sww_series(filename, names)
    if sts and src["triangles"] is not None:
        raise Exception("Files of type STS must not carry a mesh")
    if not sts and src["triangles"] is None:
        raise Exception("Files of type SWW must contain spatial information")
    starttime = src["starttime"]
    time = src["time"]
    upper = len(time)
    assert upper > 0, "Time vector obtained from file %s has length 0" % filename
    if time_limit is not None:
        limit = time_limit - starttime
        for i, t in enumerate(time):
            if t > limit:
                upper = i
                break
        assert upper > 0, "Time vector is zero. Requested time limit is %f" % limit
    time = time[:upper]
    pts = np.array(interpolation_points, dtype=np.float64)
    pts[:, 0] -= src["xllcorner"]
    pts[:, 1] -= src["yllcorner"]
    domain_starttime = None if domain is None else domain.starttime
    if domain_starttime is not None and domain_starttime > starttime:
        time = time - domain_starttime + starttime
    vertex_coordinates = np.stack([src["x"], src["y"]], axis=1)
    if sts:
        poly = np.array(boundary_polygon, dtype=np.float64)
        poly[:, 0] -= src["xllcorner"]
        poly[:, 1] -= src["yllcorner"]
        gauge_id, vertex_coordinates, neighbour = gauges_on_boundary(vertex_coordinates, poly)
        F = Interpolation_function(time, {n: np.take(src[n][:upper], gauge_id, axis=1) for n in names}, names,
                                   vertex_coordinates, None, pts, time_thinning=time_thinning,
                                   gauge_neighbour_id=neighbour)
    else:
        F = Interpolation_function(time, {n: src[n][:upper] for n in names}, names, vertex_coordinates,
                                   src["triangles"], pts, time_thinning=time_thinning)
    F.starttime = starttime
    if domain is not None and starttime > domain.starttime:
        domain.set_starttime(starttime)
    return F


def _tms_function(filename, domain, quantities, time_thinning, time_limit):
    """a TMS file (NetCDF time series without spatial information, e.g. a tide gauge or a hydrograph) as f(t):
    file_function.py:226-494 for spatial = False"""
    from scipy.io import netcdf_file
    if isinstance(quantities, str):
        quantities = [quantities]
    names = list(quantities) if quantities is not None else ["stage", "xmomentum", "ymomentum"]
    fid = netcdf_file(filename, "r", mmap=False)
    try:
        missing = [q for q in ["time"] + names if q not in fid.variables]
        if missing:
            raise Exception("Quantities %s could not be found in file %s" % (str(missing), filename))
        if "x" in fid.variables and "y" in fid.variables:
            raise Exception("Files of type TMS must not contain spatial information")
        starttime = float(np.asarray(fid.starttime).reshape(-1)[0])
        time = np.array(fid.variables["time"][:], dtype=np.float64)
        data = {n: np.array(fid.variables[n][:], dtype=np.float64) for n in names}
    finally:
        fid.close()
    upper = len(time)
    assert upper > 0, "Time vector obtained from file %s has length 0" % filename
    if time_limit is not None:
        limit = time_limit - starttime
        for i, t in enumerate(time):
            if t > limit:
                upper = i
                break
        assert upper > 0, "Time vector is zero. Requested time limit is %f" % limit
    time = time[:upper]
    domain_starttime = None if domain is None else domain.starttime
    if domain_starttime is not None and domain_starttime > starttime:
        time = time - domain_starttime + starttime
    F = Interpolation_function(time, {n: data[n][:upper] for n in names}, names, time_thinning=time_thinning)
    F.starttime = starttime
    F.filename = filename
    if domain is not None and starttime > domain.starttime:
        domain.set_starttime(starttime)
    return F


TIME_FORMAT = "%d/%m/%y %H:%M:%S"          # anuga/config.py: time_format


def timefile2netcdf(file_text, file_out=None, quantity_names=None, time_as_seconds=False):
    """text time series 'time, value0 value1 ...' (time as DD/MM/YY hh:mm:ss or, with time_as_seconds, as seconds)
    -> NetCDF TMS file for file_function (file_conversion/file_conversion.py:89-217)"""
    import calendar
    import time as _time
    from scipy.io import netcdf_file
    if file_text[-4:] != ".txt":
        raise IOError("Input file %s should be of type .txt." % file_text)
    if file_out is None:
        file_out = file_text[:-4] + ".tms"
    with open(file_text) as fh:
        lines = [ln for ln in fh.readlines()]
    assert len(lines[0].split(",")) == 2, \
        "File %s must have the format 'datetime, value0 value1 value2 ...'" % file_text

    def seconds(field):
        if time_as_seconds:
            return float(field)
        try:
            return calendar.timegm(_time.strptime(field, TIME_FORMAT))
        except ValueError:
            raise Exception("First field in file %s must be date-time with format %s.\nI got %s instead."
                            % (file_text, TIME_FORMAT, field))
    starttime = seconds(lines[0].split(",")[0])
    d = len(lines[0].split(",")[1].split())
    T = np.zeros(len(lines))
    Q = np.zeros((len(lines), d))
    for i, line in enumerate(lines):
        fields = line.split(",")
        T[i] = seconds(fields[0]) - starttime
        for j, value in enumerate(fields[1].split()):
            Q[i, j] = float(value)
    assert np.all(T[1:] - T[:-1] > 0), "File %s must list time as a monotonuosly increasing sequence" % file_text
    fid = netcdf_file(file_out, "w", version=2)
    fid.institution = "Geoscience Australia"
    fid.description = "Time series"
    fid.starttime = starttime
    fid.createDimension("number_of_timesteps", len(T))
    fid.createVariable("time", "d", ("number_of_timesteps",))[:] = T
    for i in range(d):
        try:
            name = quantity_names[i]
        except Exception:
            name = "Attribute%d" % i
        fid.createVariable(name, "d", ("number_of_timesteps",))[:] = Q[:, i]
    fid.close()


# ----------------------------------------------------------------------------------------
# boundary objects
# ----------------------------------------------------------------------------------------
class _Table_boundary(Boundary):
    """common part: frames at the midpoints of the segment's boundary edges, resident on the device"""
    device_kind = _b.BC_TIME_SPACE_TABLE
    time_dependent = True
    mean_stage = 0.0

    def boundary_point_ids(self, ids):
        """rows of the frames for the boundary edges `ids` (boundary indices)"""
        raise NotImplementedError

    def frames_for(self, ids):
        """(frames, len(ids), 3) array for the device table"""
        raise NotImplementedError


class File_boundary(_Table_boundary):
    """generic_boundary_conditions.py:518-700"""

    def __init__(self, filename, domain, time_thinning=1, time_limit=None, boundary_polygon=None,
                 default_boundary=None, use_cache=False, verbose=False):
        self.domain = domain
        self.verbose = verbose
        # one interpolation point per boundary edge, in the order of the sorted (triangle, edge) keys
        keys = sorted(domain.boundary.keys())
        self.boundary_indices = {key: i for i, key in enumerate(keys)}
        vol = np.array([k[0] for k in keys], dtype=np.int64)
        edge = np.array([k[1] for k in keys], dtype=np.int64)
        geo = getattr(domain.mesh, "geo_reference", None)
        xll = geo.get_xllcorner() if geo is not None else 0.0
        yll = geo.get_yllcorner() if geo is not None else 0.0
        self.midpoint_coordinates = domain.edge_midpoint_coordinates[3 * vol + edge] + np.array([xll, yll])
        self.F = file_function(filename, domain, quantities=domain.conserved_quantities,
                               interpolation_points=self.midpoint_coordinates, time_thinning=time_thinning,
                               time_limit=time_limit, verbose=verbose, boundary_polygon=boundary_polygon)
        assert default_boundary is None or isinstance(default_boundary, Boundary), \
            "Keyword argument default_boundary must be either None or a boundary object.\n I got %s" % default_boundary
        self.default_boundary = default_boundary
        self.default_boundary_invoked = False
        q = self.F(self.F.time[0], point_id=0)
        assert len(q) == len(domain.conserved_quantities)
        self._point_of_boundary_index = None

    def __repr__(self):
        return "File boundary"

    def _rows(self, ids):
        d = self.domain
        return np.array([self.boundary_indices[(int(d.boundary_cells[m]), int(d.boundary_edges[m]))] for m in ids],
                        dtype=np.int64)

    def frames_for(self, ids):
        rows = self._rows(ids)
        names = self.F.quantity_names
        fr = np.stack([self.F.precomputed_values[n][:, rows] for n in names], axis=2)     # (T, P, 3)
        if not np.all(np.isfinite(fr)):
            bad = rows[np.flatnonzero(~np.all(np.isfinite(fr), axis=(0, 2)))[0]]
            x, y = self.midpoint_coordinates[bad]
            raise Exception("NAN value found in file_boundary at point id #%d: (%.2f, %.2f).\n"
                            "The point lies outside the mesh stored in the file." % (bad, x, y))
        return np.ascontiguousarray(fr)

    def device_values(self, t):
        """{ratio, frame index, mean stage}: the time slot the boundary kernel interpolates in"""
        index, ratio = time_slot(self.F, t)
        return (float(ratio), float(index), float(self.mean_stage))

    @classmethod
    def adopt(cls, ref_boundary, domain):
        """take over a reference File_boundary / Field_boundary object (attach.py): its interpolation
        function with the frames it precomputed, its point numbering, its mean stage"""
        fb = getattr(ref_boundary, "file_boundary", ref_boundary)
        self = cls.__new__(cls)
        self.domain = domain
        self.verbose = getattr(fb, "verbose", False)
        self.boundary_indices = dict(fb.boundary_indices)
        self.midpoint_coordinates = np.asarray(fb.midpoint_coordinates)
        self.F = fb.F
        default = getattr(fb, "default_boundary", None)
        self.default_boundary = None if default is None else default
        self.default_boundary_invoked = False
        self.mean_stage = getattr(ref_boundary, "mean_stage", 0.0)
        if cls is Field_boundary:
            self.file_boundary = self
        return self

    def evaluate(self, vol_id=None, edge_id=None):
        q = self.F(self.domain.get_time(), point_id=self.boundary_indices[(vol_id, edge_id)])
        q[0] += 0.0 if self.mean_stage == 0.0 else self.mean_stage
        return q

    def oracle_spec(self):
        return ("file", self)


class Field_boundary(File_boundary):
    """shallow_water/boundaries.py:993-1090: a File_boundary whose stage is offset by mean_stage"""
    device_kind = _b.BC_TIME_SPACE_TABLE_MEAN_STAGE

    def __init__(self, filename, domain, mean_stage=0.0, time_thinning=1, time_limit=None, boundary_polygon=None,
                 default_boundary=None, use_cache=False, verbose=False):
        File_boundary.__init__(self, filename, domain, time_thinning=time_thinning, time_limit=time_limit,
                               boundary_polygon=boundary_polygon, default_boundary=default_boundary,
                               use_cache=use_cache, verbose=verbose)
        self.file_boundary = self
        self.mean_stage = mean_stage

    def __repr__(self):
        return "Field boundary"

    def evaluate(self, vol_id=None, edge_id=None):
        q = self.F(self.domain.get_time(), point_id=self.boundary_indices[(vol_id, edge_id)])
        q[0] += self.mean_stage
        return q


class Time_space_boundary(_Table_boundary):
    """generic_boundary_conditions.py:419-516: values from a Python function of (t, x, y), evaluated at the
    edge midpoints on the host for every RK substep (three frames on the device, rewritten every step)."""

    def __init__(self, domain=None, function=None, default_boundary=None, verbose=False):
        if function is None:
            raise Exception("You must specify a function to Time_space_boundary")
        try:
            q = function(0.0, 0.0, 0.0)
        except Exception as e:
            raise Exception("Function for time_space_boundary could not be executed:\n%s" % e)
        q = np.array(q, dtype=np.float64)
        assert len(q.shape) == 1, "ERROR: Time_space_boundary function must return a 1d list or array "
        assert len(q) == len(domain.conserved_quantities), \
            "Return value for function must be a list or an array of length %d" % len(domain.conserved_quantities)
        self.domain = domain
        self.function = function
        self.default_boundary = default_boundary
        self.default_boundary_invoked = False
        self.verbose = verbose

    def __repr__(self):
        return "Time space boundary"

    def evaluate_all(self, ids, t):
        d = self.domain
        mid = d.edge_midpoint_coordinates[3 * d.boundary_cells[ids] + d.boundary_edges[ids]]
        out = np.empty((len(ids), 3))
        for j, (x, y) in enumerate(mid):
            out[j] = np.asarray(self.function(t, x, y), dtype=np.float64)[:3]
        return out

    def frames_for(self, ids):
        f = self.evaluate_all(np.asarray(ids, dtype=np.int64), self.domain.get_time())
        return np.ascontiguousarray(np.stack([f, f, f], axis=0))

    def values_for_substep(self, dev, seg, substep, t, ids=None):
        # (one object may be bound to several tags: the segment's own edges come with the call)
        dev.set_boundary_table_frame(seg, substep, self.evaluate_all(np.asarray(ids, dtype=np.int64), t))
        return (0.0, float(substep), 0.0)

    def device_values(self, t):
        return (0.0, 0.0, 0.0)

    def oracle_spec(self):
        return ("time_space", self.function)
