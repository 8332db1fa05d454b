"""Flow and energy through a cross section at yield time (host side, numpy).

Mirrors `Domain.get_flow_through_cross_section` / `get_energy_through_cross_section`
(shallow_water/shallow_water_domain.py:1750-1810), the `Cross_section` class behind them
(shallow_water/forcing.py:680-815) and `Mesh.get_intersecting_segments`
(abstract_2d_finite_volumes/neighbour_mesh.py:1124-1167, 1252-1438).  The reference visits every triangle
of the mesh in a Python loop; here only the triangles whose bounding box meets the line's are visited, with
the reference's scalar predicates in the reference's order, so the segments (end points, order, normals,
lengths) come out equal.  Values at the segment midpoints are the quantity's (discontinuous) vertex values
of the triangle the segment lies in, weighted as `Interpolate.interpolate_block` weights them
(fit_interpolate/interpolate.py:288, utilities/quad_tree.c:26-94)."""
import math

import numpy as np

from .file_boundary import barycentric_weights

_RTOL = 1.0e-5
_ATOL = 1.0e-8


def _allclose0(a):
    # numpy.allclose(a, 0.0, rtol, atol): |a - 0| <= atol + rtol * |0|
    return abs(a) <= _ATOL


def point_on_line(x, y, x0, y0, x1, y1, rtol=_RTOL, atol=_ATOL):
    """geometry/polygon.c:29-95"""
    a0 = x - x0
    a1 = y - y0
    b0 = x1 - x0
    b1 = y1 - y0
    nominator = abs(a1 * b0 + (-a0) * b1)
    denominator = b0 * b0 + b1 * b1
    if denominator == 0.0:
        parallel = nominator <= atol
    else:
        parallel = nominator / denominator <= rtol
    if not parallel:
        return False
    len_a = math.sqrt(a0 * a0 + a1 * a1)
    len_b = math.sqrt(b0 * b0 + b1 * b1)
    return (a0 * b0 + a1 * b1) >= -1.0e-308 and -1.0e-15 <= len_b - len_a


# collinear lines: which ends lie on the other line -> the shared part (geometry/polygon.py:55-109);
# p0, p1 the first line, p2, p3 the second.  None = the reference's lines_error.
_COLLINEAR = {
    (False, False, False, False): (3, None),
    (False, False, True, True): (2, (2, 3)),
    (False, True, False, True): (2, (3, 1)),
    (False, True, True, False): (2, (2, 1)),
    (False, True, True, True): (2, (2, 3)),
    (True, False, False, True): (2, (0, 3)),
    (True, False, True, False): (2, (0, 2)),
    (True, False, True, True): (2, (2, 3)),
    (True, True, False, False): (2, (0, 1)),
    (True, True, False, True): (2, (0, 1)),
    (True, True, True, False): (2, (0, 1)),
    (True, True, True, True): (2, (0, 1)),
}


def intersection(line0, line1):
    """geometry/polygon.py:112-186.  status 1: the point (x, y); status 2: the shared part ((x, y), (x, y));
    0 none, 3 collinear apart, 4 parallel."""
    (x0, y0), (x1, y1) = line0
    (x2, y2), (x3, y3) = line1
    denom = (y3 - y2) * (x1 - x0) - (x3 - x2) * (y1 - y0)
    u0 = (x3 - x2) * (y0 - y2) - (y3 - y2) * (x0 - x2)
    u1 = (x2 - x0) * (y1 - y0) - (y2 - y0) * (x1 - x0)
    if _allclose0(denom):
        if _allclose0(u0) and _allclose0(u1):
            state = (point_on_line(x0, y0, x2, y2, x3, y3), point_on_line(x1, y1, x2, y2, x3, y3),
                     point_on_line(x2, y2, x0, y0, x1, y1), point_on_line(x3, y3, x0, y0, x1, y1))
            if state not in _COLLINEAR:
                raise RuntimeError("INTERNAL ERROR: p1=%s, p2=%s, p3=%s, p4=%s"
                                   % ([x0, y0], [x1, y1], [x2, y2], [x3, y3]))
            status, pick = _COLLINEAR[state]
            if pick is None:
                return status, None
            p = ((x0, y0), (x1, y1), (x2, y2), (x3, y3))
            return status, (p[pick[0]], p[pick[1]])
        return 4, None
    u0 = u0 / denom
    u1 = u1 / denom
    x = x0 + u0 * (x1 - x0)
    y = y0 + u0 * (y1 - y0)
    if 0.0 <= u0 <= 1.0 and 0.0 <= u1 <= 1.0:
        return 1, (x, y)
    return 0, None


def _strictly_inside_triangle(x, y, tri):
    """is_inside_polygon(point, triangle, closed=False): geometry/polygon.c:642-728 (its own rtol = atol = 0)"""
    xs = [p[0] for p in tri]
    ys = [p[1] for p in tri]
    if x > max(xs) or x < min(xs) or y > max(ys) or y < min(ys):
        return False
    inside = False
    for i in range(3):
        j = (i + 1) % 3
        px_i, py_i = tri[i]
        px_j, py_j = tri[j]
        if point_on_line(x, y, px_i, py_i, px_j, py_j, 0.0, 0.0):
            return False
        if (py_i < y and py_j >= y) or (py_j < y and py_i >= y):
            if px_i + (y - py_i) / (py_j - py_i) * (px_j - px_i) < x:
                inside = not inside
    return inside


class Triangle_intersection:
    """neighbour_mesh.py:1216-1248: segment ((x0, y0), (x1, y1)), its right-hand normal, length, triangle."""

    def __init__(self, segment=None, normal=None, length=None, triangle_id=None):
        self.segment = segment
        self.normal = normal
        self.length = length
        self.triangle_id = triangle_id

    def __repr__(self):
        return ("Triangle_intersection(segment=%s, normal=%s, length=%s, triangle_id=%s)"
                % (self.segment, self.normal, self.length, self.triangle_id))


def _candidates(V3, line):
    """ids (ascending) of the triangles whose bounding box, grown by a fraction of its size, meets the line's:
    a triangle outside can give neither a proper intersection (0 <= u <= 1 on both lines) nor a collinear
    overlap (relative tolerance 1e-5 of the edge)."""
    (xa, ya), (xb, yb) = line
    lo = V3.min(axis=1)
    hi = V3.max(axis=1)
    pad = 1.0e-3 * (hi - lo).max(axis=1) + 1.0e-7
    keep = ((hi[:, 0] + pad >= min(xa, xb)) & (lo[:, 0] - pad <= max(xa, xb))
            & (hi[:, 1] + pad >= min(ya, yb)) & (lo[:, 1] - pad <= max(ya, yb)))
    return np.flatnonzero(keep)


def _segments_of_line(V3, line):
    """neighbour_mesh.py:1252-1395 for one straight piece of the polyline"""
    line = ((float(line[0][0]), float(line[0][1])), (float(line[1][0]), float(line[1][1])))
    xi0, eta0 = line[0]
    found = {}
    for i in _candidates(V3, line):
        tri = [(float(V3[i, j, 0]), float(V3[i, j, 1])) for j in range(3)]
        hits = {}
        for j in range(3):
            status, value = intersection(line, (tri[j], tri[(j + 1) % 3]))
            if status == 1:
                hits[value] = i
            elif status == 2:
                hits[value[0]] = i
                hits[value[1]] = i
        if len(hits) == 1:
            if _strictly_inside_triangle(line[1][0], line[1][1], tri):
                hits[line[1]] = i
            elif _strictly_inside_triangle(line[0][0], line[0][1], tri):
                hits[line[0]] = i
            else:
                continue
        assert len(hits) in (0, 2), "There can be only two or no intersections"
        if len(hits) != 2:
            continue
        (x0, y0), (x1, y1) = list(hits.keys())
        d0 = math.sqrt((x0 - xi0) * (x0 - xi0) + (y0 - eta0) * (y0 - eta0))
        d1 = math.sqrt((x1 - xi0) * (x1 - xi0) + (y1 - eta0) * (y1 - eta0))
        if d1 < d0:
            x0, y0, x1, y1 = x1, y1, x0, y0
        vx, vy = x1 - x0, y1 - y0
        length = math.sqrt(vx * vx + vy * vy)
        normal = np.array([vy, -vx], dtype=np.float64) / length
        segment = ((x0, y0), (x1, y1))
        if segment not in found:
            found[segment] = Triangle_intersection(segment=segment, normal=normal, length=length,
                                                   triangle_id=int(i))
    return list(found.values())


def get_intersecting_segments(vertex_coordinates, polyline):
    """neighbour_mesh.py:1398-1438; `polyline` relative to the mesh origin"""
    assert len(polyline) >= 2, "Polyline must contain at least two points"
    V3 = np.asarray(vertex_coordinates, dtype=np.float64).reshape(-1, 3, 2)
    out = []
    for p0, p1 in zip(polyline[:-1], polyline[1:]):
        out += _segments_of_line(V3, (p0, p1))
    assert len(out) > 0, "No segments found"
    return out


def segment_midpoints(segments):
    """neighbour_mesh.py:1444-1462"""
    return [np.sum(np.array(s.segment, dtype=np.float64), axis=0) / 2 for s in segments]


class Cross_section:
    """shallow_water/forcing.py:680-815.  `polyline` in absolute coordinates."""

    def __init__(self, domain, polyline=None, verbose=False):
        self.domain = domain
        self.polyline = polyline
        self.verbose = verbose
        geo = getattr(domain.mesh, "geo_reference", None)
        origin = np.array([geo.get_xllcorner(), geo.get_yllcorner()]) if geo is not None else np.zeros(2)
        rel = np.asarray(polyline, dtype=np.float64).reshape(-1, 2) - origin
        V = np.asarray(domain.vertex_coordinates, dtype=np.float64)
        self.segments = get_intersecting_segments(V, [tuple(p) for p in rel])
        self.midpoints = np.array(segment_midpoints(self.segments))
        # weights of each midpoint in its segment's own triangle (interior point: the triangle the
        # reference's search finds as well)
        self.triangle_ids = np.array([s.triangle_id for s in self.segments], dtype=np.int64)
        self.weights = np.zeros((len(self.segments), 3))
        local = np.arange(3, dtype=np.int64).reshape(1, 3)
        for k, (tid, mid) in enumerate(zip(self.triangle_ids, self.midpoints)):
            found, w = barycentric_weights(V[3 * tid:3 * tid + 3], local, mid.reshape(1, 2))
            if found[0] < 0:        # a segment along an edge, rounding put the midpoint a hair outside
                w = _plain_weights(V[3 * tid:3 * tid + 3], mid).reshape(1, 3)
            self.weights[k] = w[0]

    def set_verbose(self, verbose=True):
        self.verbose = verbose

    def _values(self, name):
        v = self.domain.quantities[name].vertex_values[self.triangle_ids]
        out = np.zeros(len(v))
        for j in range(3):            # summed in vertex order from 0.0, as the sparse product does
            out = out + self.weights[:, j] * v[:, j]
        return out

    def get_flow_through_cross_section(self):
        uh = self._values("xmomentum")
        vh = self._values("ymomentum")
        total_flow = 0
        for i in range(len(uh)):
            normal = self.segments[i].normal
            normal_momentum = uh[i] * normal[0] + vh[i] * normal[1]
            total_flow += normal_momentum * self.segments[i].length
        return total_flow

    def get_energy_through_cross_section(self, kind="total"):
        g, epsilon, h0 = 9.8, 1.0e-12, 1.0e-6       # anuga/config.py: g, epsilon, velocity_protection
        w = self._values("stage")
        z = self._values("elevation")
        uh = self._values("xmomentum")
        vh = self._values("ymomentum")
        h = w - z
        total_line_length = 0.0
        for s in self.segments:
            total_line_length += s.length
        if kind not in ("specific", "total"):
            raise Exception('Energy kind must be either "specific" or "total". I got %s' % kind)
        average_energy = 0.0
        for i in range(len(w)):
            if h[i] > epsilon:
                u = uh[i] / (h[i] + h0 / h[i])
                v = vh[i] / (h[i] + h0 / h[i])
            else:
                u = v = 0.0
            kinetic_energy = 0.5 * (u * u + v * v) / g
            segment_energy = (h[i] if kind == "specific" else w[i]) + kinetic_energy
            average_energy += segment_energy * (self.segments[i].length / total_line_length)
        return average_energy


def _plain_weights(tri, p):
    (x1, y1), (x2, y2), (x3, y3) = tri
    det = (y2 - y3) * (x1 - x3) + (x3 - x2) * (y1 - y3)
    a = ((y2 - y3) * (p[0] - x3) + (x3 - x2) * (p[1] - y3)) / det
    b = ((y3 - y1) * (p[0] - x3) + (x1 - x3) * (p[1] - y3)) / det
    return np.array([a, b, 1.0 - a - b])
