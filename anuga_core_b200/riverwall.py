"""Riverwall edge tables from breaklines: ``domain.riverwallData.create_riverwalls(...)``.

Host-side set-up for the weir branch of the flux kernels.  Produces what
anuga/structures/riverwall.py:151-457 leaves on the domain - ``edge_flux_type`` (1 on wall edges),
``edge_river_wall_counter`` (1-based running number in (triangle, edge) order), one crest elevation and
one hydraulic-table row index per wall edge, and the hydraulic property table with the column order the
flux code assumes (Qfactor, s1, s2, h1, h2; sw_domain_openmp.c:607-617) - with the reference's
arithmetic for the membership test and the crest interpolation, so that the tables are equal bit for
bit (tests/test_host_setup.py compares them live with the reference).
"""
import numpy as np

HYDRAULIC_VARIABLE_NAMES = ("Qfactor", "s1", "s2", "h1", "h2")     # order fixed by the flux code
MAX_FLOAT = 1.0e36                                                  # anuga/config.py: max_float


class RiverWall:
    def __init__(self, domain):
        self.domain = domain
        self._reset()

    def _reset(self):
        default_float = -9.0e+20
        self.riverwall_elevation = np.array([default_float])
        self.hydraulic_properties_rowIndex = np.array([-1_000_000_000], dtype=np.int64)
        self.names = []
        self.default_riverwallPar = {"Qfactor": 1.0, "s1": 0.9, "s2": 0.95, "h1": 1.0, "h2": 1.5}
        self.hydraulic_variable_names = HYDRAULIC_VARIABLE_NAMES
        self.ncol_hydraulic_properties = len(HYDRAULIC_VARIABLE_NAMES)
        self.hydraulic_properties = np.array([[default_float]])
        self.riverwall_edges = np.array([-1_000_000_000], dtype=np.int64)
        self.input_riverwall_geo = None
        self.input_riverwallPar = None

    # ------------------------------------------------------------------------------------
    def create_riverwalls(self, riverwalls, riverwallPar={}, default_riverwallPar={}, tol=1.0e-4,
                          verbose=True, output_dir=None):
        """riverwalls: {name: [[x, y, z], ...]} polylines that coincide with mesh edges;
        riverwallPar: {name: {hydraulic parameter: value}}; see structures/riverwall.py:151-215."""
        domain = self.domain
        if not domain.get_using_discontinuous_elevation():
            raise Exception("Riverwalls are currently only supported for discontinuous elevation flow algorithms")
        if len(self.names) > 0:
            self._reset()                       # existing data is replaced, never edited
        self.input_riverwall_geo = riverwalls
        self.input_riverwallPar = riverwallPar
        for key in default_riverwallPar:
            if key not in self.default_riverwallPar:
                raise Exception("Key %s in default_riverwallPar not recognized" % key)
        self.default_riverwallPar.update(default_riverwallPar)
        defaults = self.default_riverwallPar
        for name, par in riverwallPar.items():
            if name not in riverwalls:
                raise Exception("Key %s in riverwallPar has no corresponding key in riverwall" % name)
            for key in par:
                if key not in defaults:
                    raise Exception("Hydraulic parameter named %s not recognised in default_riverwallPar" % key)

        exy = domain.edge_coordinates
        geo = getattr(domain.mesh, "geo_reference", None)
        llx = geo.get_xllcorner() if geo is not None else 0.0
        lly = geo.get_yllcorner() if geo is not None else 0.0
        n_edges = exy.shape[0]
        flux_type = np.zeros(n_edges, dtype=np.int64)
        crest = exy[:, 0] * 0. - MAX_FLOAT
        row = np.full(n_edges, -1.0)
        names = list(riverwalls.keys())
        for i, name in enumerate(names):
            line = riverwalls[name]
            for start, end in zip(line[:-1], line[1:]):
                if len(start) != 3 or len(end) != 3:
                    raise Exception("Each riverwall coordinate must have at exactly 3 values [xyz]")
                seg_len = ((start[0] - end[0]) ** 2 + (start[1] - end[1]) ** 2) ** 0.5
                if seg_len < tol:
                    continue
                # unit vector along the segment; vector from its start to every edge midpoint
                se_0 = -(start[0] - end[0]) / seg_len
                se_1 = -(start[1] - end[1]) / seg_len
                pv_0 = exy[:, 0] - (start[0] - llx)
                pv_1 = exy[:, 1] - (start[1] - lly)
                pv_len = (pv_0 ** 2 + pv_1 ** 2) ** 0.5
                along = pv_0 * se_0 + pv_1 * se_1
                perp_sq = pv_len ** 2. - along ** 2.
                on = np.flatnonzero((perp_sq < tol ** 2) * (along > 0. - tol) * (along < seg_len + tol))
                if len(on) == 0:
                    continue
                flux_type[on] = 1
                w0 = along[on] / seg_len
                w0 = w0 * (w0 >= 0.0)
                w0 = w0 * (w0 <= 1.0) + 1.0 * (w0 > 1.0)
                crest[on] = start[2] * (1.0 - w0) + w0 * end[2]
                row[on] = i
        wall = np.flatnonzero(flux_type == 1)
        self.riverwall_elevation = crest[wall]
        self.hydraulic_properties_rowIndex = row[wall].astype(int)
        self.riverwall_edges = wall
        self.names = names

        table = np.zeros((len(names), len(HYDRAULIC_VARIABLE_NAMES))) * np.nan
        for i, name in enumerate(names):
            par = riverwallPar.get(name)
            for j, var in enumerate(HYDRAULIC_VARIABLE_NAMES):
                table[i, j] = par[var] if (par is not None and var in par) else defaults[var]
        for i, name in enumerate(names):
            if table[i, 1] >= table[i, 2]:
                raise Exception("s1 >= s2 on riverwall %s. This is not allowed" % name)
            if table[i, 1] < 0. or table[i, 2] < 0.:
                raise Exception("s1 and s2 must be positive, with s1<s2")
            if table[i, 3] >= table[i, 4]:
                raise Exception("h1 >= h2 on riverwall %s. This is not allowed" % name)
            if table[i, 3] < 0. or table[i, 4] < 0.:
                raise Exception("h1 and h2 must be positive, with h1<h2")
        self.hydraulic_properties = table

        # hand the tables to the domain (edge_river_wall_counter, number_of_riverwall_edges: :395-404)
        domain.set_riverwall_tables(flux_type, self.riverwall_elevation, self.hydraulic_properties_rowIndex, table)
        ok, report = self.check_riverwall_connectedness()
        if verbose:
            print(report)
        if not ok:
            raise Exception("Riverwall discontinuity -- possible round-off error in finding edges on wall -- "
                            "try increasing value of tol")

    # ------------------------------------------------------------------------------------
    def get_centroids_corresponding_to_edgeInds(self, riverwalledgeInds):
        return np.floor(np.asarray(riverwalledgeInds) / 3.).astype(int)

    def check_riverwall_connectedness(self):
        """Every wall should be one chain of mesh edges: with round-off an edge can be missed, which shows
        as more than two edge end points that occur only once along a wall (riverwall.py:603-690)."""
        d = self.domain
        if len(self.names) == 0:
            return True, "No riverwalls"
        V = d.vertex_coordinates                      # (3N, 2): vertices 3k, 3k+1, 3k+2 of triangle k
        ok, lines = True, []
        for i, name in enumerate(self.names):
            edges = self.riverwall_edges[self.hydraulic_properties_rowIndex == i]
            if len(edges) == 0:
                lines.append("Riverwall %s has no edges on this mesh" % name)
                continue
            k, e = edges // 3, edges % 3
            a = V[3 * k + (e + 1) % 3]
            b = V[3 * k + (e + 2) % 3]
            pts = np.round(np.concatenate([a, b]), 9)
            # each wall edge is listed from both of its triangles (interior) or once (mesh boundary)
            uniq, counts = np.unique(np.unique(np.concatenate([np.minimum(a, b), np.maximum(a, b)], axis=1).round(9),
                                               axis=0).reshape(-1, 2), axis=0, return_counts=True)
            loose_ends = int(np.sum(counts == 1))
            connected = loose_ends <= 2
            ok = ok and connected
            lines.append("Riverwall %s: %d edges, %d end points%s" % (name, len(edges), loose_ends,
                                                                   "" if connected else "  <-- DISCONNECTED"))
            del pts, uniq
        return ok, "\n".join(lines)
