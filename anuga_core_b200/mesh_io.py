""".tsh mesh files: the ASCII triangulation format of the reference's mesh generator, as far as a Domain
needs it (anuga/load_mesh/loadASCII.py:143-250 _read_triangulation, abstract_2d_finite_volumes/
pmesh2domain.py:44-160, anuga/extras.py:49 create_domain_from_file).

Read: the triangulation vertices with their attribute columns (elevation ...), the triangles, and the
tagged boundary segments.  The mesh outline that follows in the file (what the generator was given) and
the .msh NetCDF variant are not needed to run a model and are not read.  Mesh set-up, outside the hot path.
"""
import numpy as np


def read_tsh(filename):
    """-> dict(vertices (P,2), vertex_attributes (P,A) | None, vertex_attribute_titles [A],
    triangles (N,3), triangle_tags [N], segments (S,2), segment_tags [S])"""
    with open(filename, "r") as fd:
        first = fd.readline().split()
        n_vert, n_att = (int(first[0]), int(first[1])) if first else (0, 0)
        vertices = np.empty((n_vert, 2), dtype=np.float64)
        attributes = np.empty((n_vert, n_att), dtype=np.float64)
        for i in range(n_vert):
            f = fd.readline().split()
            vertices[i] = (float(f[1]), float(f[2]))
            attributes[i] = [float(x) for x in f[3:3 + n_att]]
        fd.readline()                                   # "# attribute column titles"
        titles = [fd.readline().strip() for _ in range(n_att)]
        n_tri = int(fd.readline().split()[0])
        triangles = np.empty((n_tri, 3), dtype=np.int64)
        triangle_tags = []
        for i in range(n_tri):
            f = fd.readline().split()
            triangles[i] = (int(f[1]), int(f[2]), int(f[3]))
            triangle_tags.append(" ".join(f[7:]))       # after index, 3 vertices, 3 neighbours
        n_seg = int(fd.readline().split()[0])
        segments = np.empty((n_seg, 2), dtype=np.int64)
        segment_tags = []
        for i in range(n_seg):
            f = fd.readline().split()
            segments[i] = (int(f[1]), int(f[2]))
            segment_tags.append(" ".join(f[3:]))
    return dict(vertices=vertices, vertex_attributes=attributes if n_att else None,
                vertex_attribute_titles=titles, triangles=triangles, triangle_tags=triangle_tags,
                segments=segments, segment_tags=segment_tags)


def boundary_tags_from_segments(triangles, segments, segment_tags):
    """{(triangle, edge): tag} for the tagged segments that are triangle sides (pmesh2domain.py:125-160):
    side (v0,v1) is edge 2, (v1,v2) edge 0, (v2,v0) edge 1; untagged ("") segments are skipped"""
    tri = np.asarray(triangles, dtype=np.int64)
    sides = {}
    for e, (a, b) in ((2, (0, 1)), (0, (1, 2)), (1, (2, 0))):
        for k, (p, q) in enumerate(zip(tri[:, a].tolist(), tri[:, b].tolist())):
            sides[(p, q)] = (k, e)
    tags = {}
    for (v1, v2), tag in zip(np.asarray(segments).tolist(), segment_tags):
        if tag == "":
            continue
        for key in ((v1, v2), (v2, v1)):
            if key in sides:
                tags[sides[key]] = tag
    return tags


def create_domain_from_file(filename, DomainClass=None, **domain_kwargs):
    """anuga.create_domain_from_file: a Domain on the mesh of a .tsh file, its vertex attribute columns
    loaded as quantities (stage = elevation when the file has no stage column)"""
    from .domain import Domain
    if not str(filename).endswith(".tsh"):
        raise NotImplementedError("only .tsh mesh files are read (.msh needs netCDF)")
    m = read_tsh(filename)
    boundary = boundary_tags_from_segments(m["triangles"], m["segments"], m["segment_tags"])
    domain = (DomainClass or Domain)(m["vertices"], m["triangles"], boundary, **domain_kwargs)
    quantities = {}
    if m["vertex_attributes"] is not None:
        for title, column in zip(m["vertex_attribute_titles"], m["vertex_attributes"].T):
            quantities[title] = np.array(column)
    if "elevation" in quantities and "stage" not in quantities:
        quantities["stage"] = quantities["elevation"]
    for name, values in quantities.items():
        if name in domain.quantities:
            domain.set_quantity(name, values, location="vertices")
    return domain
