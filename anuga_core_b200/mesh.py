"""Host-side triangular mesh: geometry, connectivity and boundary structure.

This is the setup-time producer of the static inputs of the DE hot path
(SURVEY.md section 8(a) row a14).  Everything here runs once on the host with
vectorised numpy; the arithmetic (operation order, divisions, square roots)
follows the reference so that the static arrays are bit-identical:

* ``rectangular_cross``        - anuga/abstract_2d_finite_volumes/mesh_factory.py:138-175,
                                 mesh_factory_ext.pyx:30-88
* ``Mesh`` geometry            - general_mesh.py:156-277 (areas :170, normals :202-232,
                                 edgelengths :234-236, centroids :238-239, radii :262-267),
                                 edge midpoints :599-625
* neighbour structure          - neighbour_mesh.py:234-294 (semantics of neighbour_table.cpp)
* surrogate neighbours         - neighbour_mesh.py:318-345
* boundary enumeration         - neighbour_mesh.py:441-502, neighbour_mesh_ext.pyx:9-31
"""
import os

import numpy as np

DEFAULT_BOUNDARY_TAG = "exterior"      # anuga/config.py default_boundary_tag


def rectangular_cross(m, n, len1=1.0, len2=1.0, origin=(0.0, 0.0)):
    """Rectangular grid of m x n cells, each split in four triangles.

    Returns (points, elements, boundary) with the reference's numbering:
    grid points first (index i*(n+1)+j), then one centre point per cell in
    (i, j) order; cell c=i*n+j owns triangles 4c+0..3 = left, bottom, right,
    top; boundary edges are edge 1 of the respective triangle.
    """
    m = int(m)
    n = int(n)
    len1 = float(len1)
    len2 = float(len2)
    delta1 = len1 / m
    delta2 = len2 / n
    ng = (m + 1) * (n + 1)
    points = np.empty((ng + m * n, 2), dtype=np.float64)

    ii = np.arange(m + 1, dtype=np.float64)
    jj = np.arange(n + 1, dtype=np.float64)
    gx = ii * delta1 + origin[0]
    gy = jj * delta2 + origin[1]
    points[:ng, 0] = np.repeat(gx, n + 1)
    points[:ng, 1] = np.tile(gy, m + 1)

    ci = np.repeat(np.arange(m, dtype=np.int64), n)
    cj = np.tile(np.arange(n, dtype=np.int64), m)
    v1 = ci * (n + 1) + cj + 1            # (i, j+1)
    v2 = ci * (n + 1) + cj                # (i, j)
    v3 = (ci + 1) * (n + 1) + cj + 1      # (i+1, j+1)
    v4 = (ci + 1) * (n + 1) + cj          # (i+1, j)
    v5 = ng + ci * n + cj
    px = points[:ng, 0]
    py = points[:ng, 1]
    points[ng:, 0] = (px[v1] + px[v2] + px[v3] + px[v4]) * 0.25
    points[ng:, 1] = (py[v1] + py[v2] + py[v3] + py[v4]) * 0.25

    elements = np.empty((4 * m * n, 3), dtype=np.int64)
    elements[0::4, 0] = v2
    elements[0::4, 1] = v5
    elements[0::4, 2] = v1
    elements[1::4, 0] = v4
    elements[1::4, 1] = v5
    elements[1::4, 2] = v2
    elements[2::4, 0] = v3
    elements[2::4, 1] = v5
    elements[2::4, 2] = v4
    elements[3::4, 0] = v1
    elements[3::4, 1] = v5
    elements[3::4, 2] = v3

    boundary = {}

    def cell(i, j):             # the reference's per-cell order: left, bottom, right, top
        c = 4 * (i * n + j)
        if i == 0:
            boundary[(c + 0, 1)] = "left"
        if j == 0:
            boundary[(c + 1, 1)] = "bottom"
        if i == m - 1:
            boundary[(c + 2, 1)] = "right"
        if j == n - 1:
            boundary[(c + 3, 1)] = "top"
    # the cells at which the reference's (i, j) scan first meets each tag come first, so that
    # get_boundary_tags() lists the tags in the reference's order (left, bottom, top, right) ...
    for i, j in ((0, 0), (0, n - 1), (m - 1, 0)):
        cell(i, j)
    # ... then all boundary edges (the enumeration itself is by sorted key, neighbour_mesh.py:441-502)
    for j in range(n):
        boundary[(4 * (0 * n + j) + 0, 1)] = "left"
        boundary[(4 * ((m - 1) * n + j) + 2, 1)] = "right"
    for i in range(m):
        boundary[(4 * (i * n + 0) + 1, 1)] = "bottom"
        boundary[(4 * (i * n + n - 1) + 3, 1)] = "top"
    return points, elements, boundary


def rectangular_cross_neighbours(m, n):
    """Closed form of build_neighbour_structure for rectangular_cross(m, n) (same arrays, -1 where
    there is no neighbour): inside a cell, edge 0 of triangle k meets edge 2 of triangle k-1 and
    edge 2 meets edge 0 of triangle k+1 (mod 4); edge 1 is the cell side and meets edge 1 of the
    facing triangle of the adjacent cell (left<->right, bottom<->top)."""
    m = int(m)
    n = int(n)
    N = 4 * m * n
    nb = np.empty((N, 3), dtype=np.int64)
    ne = np.empty((N, 3), dtype=np.int64)
    base = 4 * np.arange(m * n, dtype=np.int64)
    ci = np.repeat(np.arange(m, dtype=np.int64), n)
    cj = np.tile(np.arange(n, dtype=np.int64), m)
    outer = ((base - 4 * n + 2, ci > 0), (base - 4 + 3, cj > 0), (base + 4 * n + 0, ci < m - 1),
             (base + 4 + 1, cj < n - 1))
    for k in range(4):
        nb[k::4, 0] = base + (k + 3) % 4
        nb[k::4, 2] = base + (k + 1) % 4
        target, exists = outer[k]
        nb[k::4, 1] = np.where(exists, target, -1)
        ne[k::4, 1] = np.where(exists, 1, -1)
    ne[:, 0] = 2
    ne[:, 2] = 0
    nbnd = (nb[:, 1] < 0).astype(np.int64)
    return nb, ne, nbnd


def rectangular(m, n, len1=1.0, len2=1.0, origin=(0.0, 0.0)):
    """Rectangular grid, two triangles per cell (mesh_factory.py:64-135): cell (i, j) owns
    the lower triangle 2(i*n+j) = [i4, i3, i2] and the upper one [i1, i2, i3]."""
    m = int(m)
    n = int(n)
    delta1 = float(len1) / m
    delta2 = float(len2) / n
    Np = (m + 1) * (n + 1)
    points = np.zeros((Np, 2), dtype=np.float64)
    ii = np.arange(m + 1, dtype=np.float64)
    jj = np.arange(n + 1, dtype=np.float64)
    points[:, 0] = np.repeat(ii * delta1 + origin[0], n + 1)
    points[:, 1] = np.tile(jj * delta2 + origin[1], m + 1)
    ci = np.repeat(np.arange(m, dtype=np.int64), n)
    cj = np.tile(np.arange(n, dtype=np.int64), m)
    i1 = ci * (n + 1) + cj + 1
    i2 = ci * (n + 1) + cj
    i3 = (ci + 1) * (n + 1) + cj + 1
    i4 = (ci + 1) * (n + 1) + cj
    elements = np.zeros((2 * m * n, 3), dtype=np.int64)
    elements[0::2] = np.stack([i4, i3, i2], axis=1)
    elements[1::2] = np.stack([i1, i2, i3], axis=1)
    boundary = {}
    for i in range(m):
        for j in range(n):
            nt = 2 * (i * n + j)
            if i == m - 1:
                boundary[(nt, 2)] = "right"
            if j == 0:
                boundary[(nt, 1)] = "bottom"
            if i == 0:
                boundary[(nt + 1, 2)] = "left"
            if j == n - 1:
                boundary[(nt + 1, 1)] = "top"
    return points, elements, boundary


def build_neighbour_structure(triangles, number_of_nodes):
    """neighbours, neighbour_edges, number_of_boundaries from the triangle table.

    Edge e of a triangle is the edge opposite vertex e: edge 0 = (v1, v2),
    edge 1 = (v2, v0), edge 2 = (v0, v1).  The neighbour across a directed edge
    (a, b) is the triangle owning (b, a).  Missing neighbours stay -1.
    """
    tri = np.ascontiguousarray(triangles, dtype=np.int64)
    N = tri.shape[0]
    if N >= 200000 and not os.environ.get("SWK_NO_NATIVE_SETUP"):   # large meshes: libswk's native helper (same result, ~10x faster)
        from . import backend
        res = backend.build_neighbour_structure_native(tri, number_of_nodes)
        if res is not None:
            return res
    nn = np.int64(number_of_nodes)
    src = np.empty((N, 3), dtype=np.int64)
    dst = np.empty((N, 3), dtype=np.int64)
    src[:, 0] = tri[:, 1]; dst[:, 0] = tri[:, 2]
    src[:, 1] = tri[:, 2]; dst[:, 1] = tri[:, 0]
    src[:, 2] = tri[:, 0]; dst[:, 2] = tri[:, 1]
    key = (src * nn + dst).ravel()
    rev = (dst * nn + src).ravel()
    order = np.argsort(key, kind="stable")
    skey = key[order]
    if skey.size > 1 and np.any(skey[1:] == skey[:-1]):
        dup = int(order[1:][skey[1:] == skey[:-1]][0])
        raise Exception("Edge %d of triangle %d is duplicating an edge of another triangle"
                        % (dup % 3, dup // 3))
    pos = np.searchsorted(skey, rev)
    pos[pos >= skey.size] = skey.size - 1
    found = skey[pos] == rev
    owner = order[pos]
    neighbours = np.where(found, owner // 3, -1).reshape(N, 3)
    neighbour_edges = np.where(found, owner % 3, -1).reshape(N, 3)
    number_of_boundaries = 3 - found.reshape(N, 3).sum(axis=1)
    return (neighbours.astype(np.int64), neighbour_edges.astype(np.int64),
            number_of_boundaries.astype(np.int64))


class Mesh:
    """Static mesh data with the reference's attribute names and layouts."""

    def __init__(self, coordinates, triangles, boundary=None,
                 use_inscribed_circle=False, neighbour_structure=None):
        self.nodes = np.ascontiguousarray(coordinates, dtype=np.float64)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int64)
        if self.nodes.ndim != 2 or self.nodes.shape[1] != 2:
            raise ValueError("coordinates must be an (M, 2) array")
        if self.triangles.ndim != 2 or self.triangles.shape[1] != 3:
            raise ValueError("triangles must be an (N, 3) array")
        self.number_of_nodes = self.nodes.shape[0]
        N = self.number_of_triangles = self.triangles.shape[0]
        self.use_inscribed_circle = use_inscribed_circle

        native = None
        if N >= 200000 and not os.environ.get("SWK_NO_NATIVE_SETUP"):   # large meshes: one pass in libswk's host helper (same bits)
            from . import backend
            native = backend.mesh_geometry_native(self.nodes, self.triangles, use_inscribed_circle)
        if native is not None:
            (self.vertex_coordinates, self.areas, self.normals, self.edgelengths, self.centroid_coordinates,
             self.radii, self.edge_midpoint_coordinates, bad) = native
            if bad >= 0:
                raise AssertionError("Degenerate Triangle(s) " + str(np.where(self.areas <= 0.0)[0]))
        else:
            self._geometry_numpy(use_inscribed_circle)

        if neighbour_structure is None:
            neighbour_structure = build_neighbour_structure(self.triangles, self.number_of_nodes)
        # (a sub-mesh cut out of a larger mesh may pass the structure it already knows)
        self.neighbours, self.neighbour_edges, self.number_of_boundaries = neighbour_structure
        rng = np.arange(N, dtype=np.int64)[:, None]
        self.surrogate_neighbours = np.where(self.neighbours < 0, rng, self.neighbours)

        self.vertex_value_indices = None
        self._build_boundary(boundary)

    def _geometry_numpy(self, use_inscribed_circle):
        N = self.number_of_triangles
        i0 = self.triangles[:, 0]
        i1 = self.triangles[:, 1]
        i2 = self.triangles[:, 2]
        V = np.empty((3 * N, 2), dtype=np.float64)
        V[0::3] = self.nodes[i0]
        V[1::3] = self.nodes[i1]
        V[2::3] = self.nodes[i2]
        self.vertex_coordinates = V
        x0 = V[0::3, 0]; y0 = V[0::3, 1]
        x1 = V[1::3, 0]; y1 = V[1::3, 1]
        x2 = V[2::3, 0]; y2 = V[2::3, 1]

        self.areas = -((x1 * y0 - x0 * y1) + (x2 * y1 - x1 * y2) + (x0 * y2 - x2 * y0)) / 2.0
        if not np.all(self.areas > 0.0):
            bad = np.where(self.areas <= 0.0)[0]
            raise AssertionError("Degenerate Triangle(s) " + str(bad))

        self.normals = np.empty((N, 6), dtype=np.float64)
        self.edgelengths = np.empty((N, 3), dtype=np.float64)
        for e, (xa, ya, xb, yb) in enumerate(((x1, y1, x2, y2), (x2, y2, x0, y0), (x0, y0, x1, y1))):
            xn = xb - xa
            yn = yb - ya
            l = np.sqrt(xn * xn + yn * yn)
            xn = xn / l
            yn = yn / l
            self.normals[:, 2 * e] = yn
            self.normals[:, 2 * e + 1] = -xn
            self.edgelengths[:, e] = l

        self.centroid_coordinates = np.empty((N, 2), dtype=np.float64)
        self.centroid_coordinates[:, 0] = (x0 + x1 + x2) / 3
        self.centroid_coordinates[:, 1] = (y0 + y1 + y2) / 3
        cx = self.centroid_coordinates[:, 0]
        cy = self.centroid_coordinates[:, 1]

        if not use_inscribed_circle:
            xm0 = (x1 + x2) / 2; ym0 = (y1 + y2) / 2
            xm1 = (x2 + x0) / 2; ym1 = (y2 + y0) / 2
            xm2 = (x0 + x1) / 2; ym2 = (y0 + y1) / 2
            d0 = np.sqrt((cx - xm0) ** 2 + (cy - ym0) ** 2)
            d1 = np.sqrt((cx - xm1) ** 2 + (cy - ym1) ** 2)
            d2 = np.sqrt((cx - xm2) ** 2 + (cy - ym2) ** 2)
            self.radii = np.minimum(np.minimum(d0, d1), d2)
        else:
            a = np.sqrt((x0 - x1) ** 2 + (y0 - y1) ** 2)
            b = np.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2)
            c = np.sqrt((x2 - x0) ** 2 + (y2 - y0) ** 2)
            self.radii = 2.0 * self.areas / (a + b + c)

        E = np.empty((3 * N, 2), dtype=np.float64)
        E[0::3] = 0.5 * (V[1::3] + V[2::3])
        E[1::3] = 0.5 * (V[2::3] + V[0::3])
        E[2::3] = 0.5 * (V[0::3] + V[1::3])
        self.edge_midpoint_coordinates = E

    def __len__(self):
        return self.number_of_triangles

    # -- boundary ---------------------------------------------------------
    def _build_boundary(self, boundary):
        N = self.number_of_triangles
        boundary = dict(boundary) if boundary else {}
        for (vol_id, edge_id) in boundary:
            assert vol_id < N and edge_id < 3, "Segment (%d, %d) does not exist" % (vol_id, edge_id)
        vols, edges = np.nonzero(self.neighbours < 0)
        for v, e in zip(vols.tolist(), edges.tolist()):
            if (v, e) not in boundary:
                boundary[(v, e)] = DEFAULT_BOUNDARY_TAG
        self.boundary = boundary
        self.boundary_length = len(boundary)

        X = sorted(boundary.keys())
        M = len(X)
        self.boundary_cells = np.zeros((M,), dtype=np.int64)
        self.boundary_edges = np.zeros((M,), dtype=np.int64)
        self.boundary_enumeration = {}
        for j, (vid, e) in enumerate(X):
            self.neighbours[vid, e] = -(j + 1)
            self.boundary_enumeration[(vid, e)] = j
            self.boundary_cells[j] = vid
            self.boundary_edges[j] = e
        self.boundary_tags_by_index = [boundary[k] for k in X]
        self.tag_boundary_cells = {}
        for tag in self.get_boundary_tags():
            self.tag_boundary_cells[tag] = []
        for j, tag in enumerate(self.boundary_tags_by_index):
            self.tag_boundary_cells[tag].append(j)

    def build_inverted_triangle_structure(self):
        """number_of_triangles_per_node and vertex_value_indices: the vertex slots 3*triangle+vertex
        grouped by node, in the order the reference uses (general_mesh.py:740-850: an argsort of the
        flattened triangle table with numpy's default sort)."""
        if getattr(self, "vertex_value_indices", None) is None:
            flat = self.triangles.reshape(-1)
            count = np.bincount(flat, minlength=self.number_of_nodes).astype(np.int64)
            if np.any(count == 0):
                raise NotImplementedError("nodes that belong to no triangle are not supported")
            self.number_of_triangles_per_node = count
            self.vertex_value_indices = np.argsort(flat).astype(np.int64)
        return self.number_of_triangles_per_node, self.vertex_value_indices

    def get_boundary_tags(self):
        tags = {}
        for v in self.boundary.values():
            tags[v] = 1
        return list(tags.keys())

    # -- accessors with the reference's names ---------------------------
    def get_centroid_coordinates(self):
        return self.centroid_coordinates

    def get_vertex_coordinates(self):
        return self.vertex_coordinates

    def get_edge_midpoint_coordinates(self):
        return self.edge_midpoint_coordinates

    def get_areas(self):
        return self.areas


class Topology:
    """neighbours + boundary of a triangle table without the geometry (what the partition needs)"""

    def __init__(self, number_of_nodes, triangles, boundary, neighbour_structure=None):
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int64)
        self.number_of_triangles = len(self.triangles)
        if neighbour_structure is None:
            neighbour_structure = build_neighbour_structure(self.triangles, number_of_nodes)
        self.neighbours, self.neighbour_edges, self.number_of_boundaries = neighbour_structure
        b = dict(boundary) if boundary else {}
        vols, edges = np.nonzero(self.neighbours < 0)
        for v, e in zip(vols.tolist(), edges.tolist()):
            if (v, e) not in b:
                b[(v, e)] = DEFAULT_BOUNDARY_TAG
        self.boundary = b


def morton_order(centroid_coordinates, bits=21):
    """Locality ordering of triangles: permutation that sorts centroids along a
    Z-order (Morton) curve.  ``perm[new] = old``.  Used by the device backend so
    that neighbour gathers hit nearby sectors (north_star, subsystem 1)."""
    c = np.asarray(centroid_coordinates, dtype=np.float64)
    lo = c.min(axis=0)
    span = c.max(axis=0) - lo
    span[span == 0.0] = 1.0
    q = ((c - lo) / span * ((1 << bits) - 1)).astype(np.uint64)

    def spread(v):
        v = v & np.uint64(0x1FFFFF)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v

    # interleave x and y (use the 3-way spread with a zero lane: still monotone)
    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))
    return np.argsort(code, kind="stable").astype(np.int64)
