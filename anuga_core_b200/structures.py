"""Inlet_operator: water added / removed over a small set of triangles each timestep.

Mirrors anuga/structures/inlet.py:11-237 (Inlet: the exchange region and its averages,
set_stages_evenly :192-227) and anuga/structures/inlet_operator.py:9-157 (Inlet_operator.__call__).
The hydraulics are scalar / O(inlet triangles) numpy on the HOST, on values gathered from the
device (swk_gather_centroids) and scattered back (swk_scatter_centroids); the big arrays stay
resident.  SURVEY.md section 8(f) row 1.

Structure_operator / Boyd_box_operator (structures/structure_operator.py:13-330,
structures/boyd_box_operator.py:8-460, structures/inlet_enquiry.py): a culvert moving water between
two such exchange regions, driven by the energy difference at two enquiry triangles.  Same split:
gather (inlet triangles + enquiry triangle), scalar hydraulics on the host, scatter.

Region resolution is mesh set-up geometry (anuga/geometry), outside the hot path: a Region here is
a list of triangle indices, a circle, a polygon (centroids inside, optionally plus every triangle
cut by the outline: the reference's expand_polygon) or a line (triangles it cuts).
"""
import math

import numpy as np

velocity_protection = 1.0e-6        # anuga/config.py:18
g = 9.8                             # anuga/config.py


class Region:
    """abstract_2d_finite_volumes/region.py:24-150, for indices / center+radius / polygon (centroids)"""

    def __init__(self, domain, indices=None, polygon=None, center=None, radius=None, line=None, poly=None,
                 expand_polygon=False, verbose=False):
        self.domain = domain
        # how the region was specified: a distributed domain re-evaluates it on the gathered patch
        self.spec = dict(indices=indices, polygon=polygon, center=center, radius=radius, line=line, poly=poly,
                         expand_polygon=expand_polygon)
        c = domain.centroid_coordinates
        if poly is not None:                      # region.py:118-131: 2 points = line, more = polygon
            assert indices is None and polygon is None and line is None and center is None
            poly = np.asarray(poly, dtype=np.float64)
            if len(poly) > 2:
                polygon = poly
            else:
                line = poly
        if indices is not None:
            self.indices = np.asarray(indices, dtype=np.int64)
            self.type = "user_defined"
        elif center is not None and radius is not None:
            d2 = (c[:, 0] - center[0]) ** 2 + (c[:, 1] - center[1]) ** 2
            self.indices = np.flatnonzero(d2 < radius ** 2).astype(np.int64)
            self.type = "circle"
        elif polygon is not None:
            polygon = np.asarray(polygon, dtype=np.float64)
            cand = _near(domain, polygon)
            ids = cand[_inside_polygon(c[cand], polygon)].astype(np.int64)
            if expand_polygon:                    # region.py:252-257
                n = len(polygon)
                for j in range(n):
                    ids = np.union1d(ids, triangles_cut_by_segment(domain, polygon[j], polygon[(j + 1) % n]))
            self.indices = ids.astype(np.int64)
            self.type = "polygon"
        elif line is not None:
            line = np.asarray(line, dtype=np.float64)
            self.indices = triangles_cut_by_segment(domain, line[0], line[1])
            self.type = "line"
        else:
            self.indices = None
            self.type = "all"
        if self.indices is None:
            self.full_indices = np.flatnonzero(domain.tri_full_flag == 1)
        else:
            self.full_indices = self.indices[domain.tri_full_flag[self.indices] == 1]

    def get_indices(self, full_only=True):
        return self.full_indices if full_only else self.indices


def _inside_polygon(points, poly):
    x, y = points[:, 0], points[:, 1]
    inside = np.zeros(len(points), dtype=bool)
    n = len(poly)
    j = n - 1
    for i in range(n):
        xi, yi = poly[i]
        xj, yj = poly[j]
        cross = ((yi > y) != (yj > y)) & (x < (xj - xi) * (y - yi) / (yj - yi + 1e-300) + xi)
        inside ^= cross
        j = i
    return inside


def _near(domain, pts):
    """ids of the triangles that can touch the bounding box of `pts`: centroid within the box grown
    by the longest edge of the mesh (keeps the exact predicates below off the other millions)"""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    c = domain.centroid_coordinates
    pad = float(np.max(domain.edgelengths))
    lo = pts.min(axis=0) - pad
    hi = pts.max(axis=0) + pad
    return np.flatnonzero((c[:, 0] >= lo[0]) & (c[:, 0] <= hi[0]) & (c[:, 1] >= lo[1]) & (c[:, 1] <= hi[1]))


def triangles_cut_by_segment(domain, p0, p1):
    """ids (ascending) of the triangles a segment touches: one of their sides meets the segment
    (closed ends), or the segment lies strictly inside.  Same predicate as the reference's
    geometry/polygon.c:446-523, evaluated for all candidate triangles at once."""
    p0 = np.asarray(p0, dtype=np.float64)
    p1 = np.asarray(p1, dtype=np.float64)
    cand = _near(domain, [p0, p1])
    V = np.asarray(domain.vertex_coordinates, dtype=np.float64).reshape(-1, 3, 2)[cand]
    u = p1 - p0
    perp_u = np.array([-u[1], u[0]])
    hit = np.zeros(len(V), dtype=bool)
    beyond = np.zeros(len(V), dtype=np.int64)      # sides crossed by the supporting line past p1
    before = np.zeros(len(V), dtype=np.int64)      # ... and ahead of p0
    for j in range(3):
        t0 = V[:, j]
        w = V[:, (j + 1) % 3] - t0
        perp_w = np.stack([-w[:, 1], w[:, 0]], axis=1)
        den = u[0] * perp_w[:, 0] + u[1] * perp_w[:, 1]
        ok = den != 0.0                           # parallel sides never count
        v = t0 - p0
        with np.errstate(divide="ignore", invalid="ignore"):
            a = (v[:, 0] * perp_w[:, 0] + v[:, 1] * perp_w[:, 1]) / den
            b = -(v[:, 0] * perp_u[0] + v[:, 1] * perp_u[1]) / (w[:, 0] * perp_u[0] + w[:, 1] * perp_u[1])
        on_side = ok & (b >= 0.0) & (b <= 1.0)
        hit |= on_side & (a >= 0.0) & (a <= 1.0)
        beyond += on_side & (a > 1.0)
        before += on_side & (a < 0.0)
    hit |= (beyond >= 1) & (before >= 1)
    return cand[hit].astype(np.int64)


def triangle_containing_point(domain, point):
    """lowest id of a triangle whose closed hull holds the point (neighbour_mesh.py:1057-1080)"""
    cand = _near(domain, [point])
    V = np.asarray(domain.vertex_coordinates, dtype=np.float64).reshape(-1, 3, 2)[cand]
    x, y = float(point[0]), float(point[1])
    inside = np.ones(len(V), dtype=bool)
    for j in range(3):
        a, b = V[:, j], V[:, (j + 1) % 3]
        cross = (b[:, 0] - a[:, 0]) * (y - a[:, 1]) - (b[:, 1] - a[:, 1]) * (x - a[:, 0])
        inside &= cross >= -1.0e-12
    ids = cand[inside]
    if len(ids) == 0:
        raise Exception("Point %s not found within a triangle" % str(point))
    return int(ids[0])


# ----------------------------------------------------------------------------------------
# structures created directly on a distributed domain (the reference's parallel scripts create them after
# distribute(): parallel/parallel_operator_factory.py, parallel_inlet.py, parallel_structure_operator.py)
# ----------------------------------------------------------------------------------------
def is_distributed(domain):
    return getattr(domain, "numproc", 1) > 1 and getattr(domain, "_comm", None) is not None \
        and getattr(domain, "tri_l2s", None) is not None and not isinstance(domain, PatchDomain)


class _PatchQuantity:
    def __init__(self, values):
        self.centroid_values = values
        self.host_dirty = False


class PatchDomain:
    """The triangles of a distributed domain that lie near a structure, gathered on EVERY rank in the order
    of their ids in the undistributed numbering.  It stands in for the sequential domain while an inlet or a
    culvert resolves its geometry (regions, enquiry triangles, areas, the initial state that primes the
    smoothing memory), so that every rank ends up with the same object the sequential constructor would have
    made; `seq_ids` then takes patch indices back to undistributed ids."""

    conserved_quantities = ["stage", "xmomentum", "ymomentum"]

    def __getattr__(self, name):
        # scalars of the model (timestep, g, yieldstep, ...) are the distributed domain's
        if name.startswith("__") or "_sub" not in self.__dict__:
            raise AttributeError(name)
        return getattr(self.__dict__["_sub"], name)

    def __init__(self, sub, lo, hi):
        self._sub = sub
        comm = sub._comm
        nf = sub.number_of_full_triangles
        # reach of the exact predicates' candidate filter (_near): the longest edge of the WHOLE mesh
        self.pad = comm.allreduce_max(float(np.max(sub.edgelengths)))
        lo = np.asarray(lo, dtype=np.float64) - 2.0 * self.pad
        hi = np.asarray(hi, dtype=np.float64) + 2.0 * self.pad
        c = sub.centroid_coordinates[:nf]
        mine = np.flatnonzero((c[:, 0] >= lo[0]) & (c[:, 0] <= hi[0]) & (c[:, 1] >= lo[1]) & (c[:, 1] <= hi[1]))
        q = sub.quantities
        rows = np.concatenate([
            np.asarray(sub.tri_l2s[mine], dtype=np.int64).view(np.float64)[:, None],   # ids travel as bit patterns
            np.asarray(sub.vertex_coordinates, dtype=np.float64).reshape(-1, 6)[mine],
            c[mine], sub.areas[mine][:, None], sub.edgelengths[mine],
            np.stack([q[n].centroid_values[mine] for n in ("stage", "xmomentum", "ymomentum", "elevation")], axis=1),
        ], axis=1)
        counts = np.zeros(comm.size)
        counts[comm.rank] = len(mine)
        counts = comm.merge_disjoint(counts).astype(np.int64)
        start = int(np.sum(counts[:comm.rank]))
        allrows = np.zeros((int(np.sum(counts)), rows.shape[1]))
        allrows[start:start + len(mine)] = rows
        allrows = comm.merge_disjoint(allrows)                 # bit-exact: every row has one owner
        ids = np.ascontiguousarray(allrows[:, 0]).view(np.int64)
        order = np.argsort(ids, kind="stable")
        allrows = allrows[order]
        self.seq_ids = ids[order]
        P = len(self.seq_ids)
        self.number_of_triangles = P
        self.vertex_coordinates = np.ascontiguousarray(allrows[:, 1:7]).reshape(3 * P, 2)
        self.centroid_coordinates = np.ascontiguousarray(allrows[:, 7:9])
        self.areas = np.ascontiguousarray(allrows[:, 9])
        self.edgelengths = np.ascontiguousarray(allrows[:, 10:13])
        if P:
            self.edgelengths[0, 0] = max(self.edgelengths[0, 0], self.pad)     # _near reads the maximum
        self.tri_full_flag = np.ones(P, dtype=np.int64)
        self.quantities = {n: _PatchQuantity(np.ascontiguousarray(allrows[:, 13 + j]))
                           for j, n in enumerate(("stage", "xmomentum", "ymomentum", "elevation"))}
        self.operators = []

    def set_fractional_step_operator(self, op):
        self.operators.append(op)

    def get_centroid_coordinates(self, absolute=False):
        return self.centroid_coordinates


def _bbox(*point_sets):
    pts = np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 2) for p in point_sets if p is not None])
    return pts.min(axis=0), pts.max(axis=0)


def _relabel_inlet(inlet, patch):
    inlet.triangle_indices = patch.seq_ids[inlet.triangle_indices]
    inlet.region = None
    if hasattr(inlet, "enquiry_index"):
        inlet.enquiry_index = int(patch.seq_ids[inlet.enquiry_index])
        inlet._extra_ids = np.array([inlet.enquiry_index], dtype=np.int64)
    return inlet


def fetch_inlets(inlets):
    """Load the rows of several inlets that are read at the same point of the time loop (the two ends of a
    culvert): every rank gathers the rows it owns and ONE exact merge makes all of them visible everywhere."""
    parts = [i._gather_local() for i in inlets]
    if all(rows is None for rows, _ in parts):
        for inlet, (_, got) in zip(inlets, parts):
            inlet._take(got)
        return
    sizes = [i._n_rows() + i._n_extra for i in inlets]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    allgot = np.zeros((int(offs[-1]), 4), dtype=np.float64)
    for (rows, got), o in zip(parts, offs[:-1]):
        if len(rows):
            allgot[o + rows] = got
    comm = getattr(inlets[0].domain, "_comm", None)
    if comm is None:
        raise RuntimeError("a distributed inlet needs domain.attach_communicator(comm) before evolve")
    allgot = comm.merge_disjoint(allgot)
    for inlet, o, n in zip(inlets, offs[:-1], sizes):
        inlet._take(allgot[o:o + n].copy())


def _owned(sub, global_ids):
    """rows of `global_ids` (ids in the undistributed numbering) that are FULL triangles of
    sub-domain `sub`, and their local ids"""
    nf = sub.number_of_full_triangles
    l2s = np.asarray(sub.tri_l2s[:nf], dtype=np.int64)
    order = np.argsort(l2s, kind="stable")
    sorted_ids = l2s[order]
    gid = np.asarray(global_ids, dtype=np.int64).reshape(-1)
    pos = np.clip(np.searchsorted(sorted_ids, gid), 0, max(nf - 1, 0))
    hit = sorted_ids[pos] == gid if nf else np.zeros(len(gid), dtype=bool)
    return np.flatnonzero(hit).astype(np.int64), order[pos[hit]].astype(np.int64)


class Inlet:
    """The exchange region: views of the inlet triangles' centroid values, held on the host for the
    duration of one operator call (structures/inlet.py:11-237)."""

    def __init__(self, domain, region, verbose=False):
        self.domain = domain
        if isinstance(region, Region):
            self.region = region
        else:
            arr = np.asarray(region)
            if arr.ndim == 2 and arr.shape[1] == 2:      # a line or polygon, as the reference takes it (inlet.py:26-29)
                self.region = Region(domain, poly=arr.astype(np.float64), expand_polygon=True)
            else:
                self.region = Region(domain, indices=arr)
        self.triangle_indices = np.asarray(self.region.indices, dtype=np.int64)
        if len(self.triangle_indices) == 0:
            raise Exception("No triangles have been identified in region ")
        self.areas = domain.areas[self.triangle_indices]
        self.area = float(np.sum(self.areas))
        assert self.area > 0.0
        self.values = None          # (n,4) stage, xmom, ymom, elevation; loaded by fetch()
        self.local_rows = None      # distributed: rows of `values` whose triangle is full on this rank
        self._extra_ids = np.zeros(0, dtype=np.int64)      # further triangles fetched with the inlet
        self._extra_rows = np.zeros(0, dtype=np.int64)

    # -- multi-GPU ---------------------------------------------------------------------
    def localise(self, sub):
        """The same inlet seen from sub-domain `sub` of a distributed run (parallel/parallel_inlet.py
        restated): `values` / `areas` keep the layout of the undistributed inlet (rows in ascending
        global id), every rank fills the rows of the triangles it owns, one exact merge makes all
        rows visible everywhere and every rank evaluates the same scalar hydraulics; only owned rows
        are written back.  Ghost copies follow at the update_ghosts that ends the step."""
        import copy
        new = copy.copy(self)
        new.domain = sub
        new.local_rows, new.triangle_indices = _owned(sub, self.triangle_indices)
        new.values = None
        return new

    def _n_rows(self):
        return len(self.areas)

    # -- device exchange ---------------------------------------------------------------
    def fetch(self):
        fetch_inlets([self])

    def _gather_local(self):
        """(rows of this rank in the inlet's full layout, their values) - or (None, all rows) undistributed"""
        dev = self.domain._dev
        ids = np.concatenate([self.triangle_indices, self._extra_ids]).astype(np.int64)
        if self.local_rows is None:
            return None, dev.gather_centroids(ids)
        rows = np.concatenate([self.local_rows, self._extra_rows]).astype(np.int64)
        return rows, (dev.gather_centroids(ids) if len(ids) else np.zeros((0, 4)))

    def _take(self, got):
        self.values = got

    def commit(self):
        if self.local_rows is None:
            self.domain._dev.scatter_centroids(self.triangle_indices, self.values[:, :3])
        elif len(self.triangle_indices):
            self.domain._dev.scatter_centroids(self.triangle_indices, self.values[self.local_rows, :3])
    _n_extra = 0

    # -- the reference's accessors ----------------------------------------------------
    def get_area(self):
        return self.area

    def get_areas(self):
        return self.areas

    def get_stages(self):
        return self.values[:, 0]

    def get_elevations(self):
        return self.values[:, 3]

    def get_xmoms(self):
        return self.values[:, 1]

    def get_ymoms(self):
        return self.values[:, 2]

    def get_depths(self):
        return self.get_stages() - self.get_elevations()

    def get_total_water_volume(self):
        return np.sum(self.get_depths() * self.get_areas())

    def get_average_depth(self):
        return self.get_total_water_volume() / self.area

    def get_average_stage(self):
        return np.sum(self.get_stages() * self.get_areas()) / self.area

    def get_average_xmom(self):
        return np.sum(self.get_xmoms() * self.get_areas()) / self.area

    def get_average_ymom(self):
        return np.sum(self.get_ymoms() * self.get_areas()) / self.area

    def get_average_elevation(self):
        return np.sum(self.get_elevations() * self.get_areas()) / self.area

    def get_velocities(self):
        depths = self.get_depths()
        u = self.get_xmoms() * depths / (depths * depths + velocity_protection)
        v = self.get_ymoms() * depths / (depths * depths + velocity_protection)
        return u, v

    def get_xvelocities(self):
        return self.get_velocities()[0]

    def get_yvelocities(self):
        return self.get_velocities()[1]

    def get_average_speed(self):
        u, v = self.get_velocities()
        average_u = np.sum(u * self.get_areas()) / self.area
        average_v = np.sum(v * self.get_areas()) / self.area
        return math.sqrt(average_u ** 2 + average_v ** 2)

    def get_average_velocity_head(self):
        return 0.5 * self.get_average_speed() ** 2 / g

    def get_average_total_energy(self):
        return self.get_average_velocity_head() + self.get_average_stage()

    def get_average_specific_energy(self):
        return self.get_average_velocity_head() + self.get_average_depth()

    def set_depths_evenly(self, volume):
        """the volume spread as one extra depth over the region (inlet.py:230-236)"""
        self.set_depths(self.get_average_depth() + volume / self.get_area())

    def set_depths(self, depth):
        self.values[:, 0] = self.get_elevations() + depth

    def set_stages(self, stage):
        self.values[:, 0] = stage

    def set_xmoms(self, xmom):
        self.values[:, 1] = xmom

    def set_ymoms(self, ymom):
        self.values[:, 2] = ymom

    def set_stages_evenly(self, volume):
        """level surface over the lowest cells that `volume` can fill (inlet.py:192-227)"""
        assert volume >= 0.0
        areas = self.get_areas()
        stages = self.get_stages().copy()
        stages_order = stages.argsort()
        summed_areas = np.cumsum(areas[stages_order])
        summed_volume = np.zeros_like(areas)
        summed_volume[1:] = np.cumsum(summed_areas[:-1] * np.diff(stages[stages_order]))
        index = np.nonzero(summed_volume <= volume)[0][-1]
        depth = (volume - summed_volume[index]) / summed_areas[index]
        stages[stages_order[0:index + 1]] = stages[stages_order[index]] + depth
        self.set_stages(stages)


class Inlet_operator:
    """anuga.Inlet_operator(domain, region, Q=..., velocity=None, zero_velocity=False)"""
    time_dependent = True            # evaluated on the host every step
    host_side = True

    def __init__(self, domain, region, Q=0.0, velocity=None, zero_velocity=False, default=0.0,
                 description=None, label=None, logging=False, verbose=False):
        if is_distributed(domain):
            # resolve the region on the gathered patch (identical on every rank), then keep the owned rows
            sub = domain
            spec = region.spec if isinstance(region, Region) else dict(poly=np.asarray(region, dtype=np.float64),
                                                                      expand_polygon=True)
            if spec.get("indices") is not None:
                raise NotImplementedError("a Region given by local indices cannot be resolved across ranks: "
                                          "give a line, polygon or circle")
            geom = [spec.get(k) for k in ("polygon", "line", "poly")]
            if spec.get("center") is not None:
                ctr, r = np.asarray(spec["center"], dtype=np.float64), float(spec["radius"])
                geom.append(np.array([ctr - r, ctr + r]))
            patch = PatchDomain(sub, *_bbox(*geom))
            inlet = _relabel_inlet(Inlet(patch, Region(patch, **spec), verbose=verbose), patch)
            self.domain = sub
            self.inlet = inlet.localise(sub)
        else:
            self.domain = domain
            self.inlet = Inlet(domain, region, verbose=verbose)
        self.Q = Q
        if velocity is not None:
            assert len(velocity) == 2
        self.velocity = velocity
        self.zero_velocity = zero_velocity
        self.default = default
        self.applied_Q = 0.0
        self.total_applied_volume = 0.0
        self.total_requested_volume = 0.0
        self.total_applied_volume = 0.0
        self.description = " " if description is None else description
        self.label = "inlet" if label is None else label
        self.verbose = verbose
        domain.set_fractional_step_operator(self)

    def update_Q(self, t):
        if callable(self.Q):
            try:
                return float(self.Q(t))
            except Exception:
                return float(self.default)
        return float(self.Q)

    def set_Q(self, Q):
        self.Q = Q

    def get_Q(self):
        return self.Q

    def get_applied_Q(self):
        return self.applied_Q

    def get_total_applied_volume(self):
        return self.total_applied_volume

    def get_inlet(self):
        return self.inlet

    def set_label(self, label):
        self.label = label

    # -- reporting (inlet_operator.py:186-233) ----------------------------------------------
    def statistics(self):
        inlet = self.inlet
        message = "=====================================\n"
        message += "Inlet Operator: %s\n" % self.label
        message += "=====================================\n"
        message += "Description\n"
        message += "%s" % self.description
        message += "\n"
        message += "-------------------------------------\n"
        message += "Inlet\n"
        message += "-------------------------------------\n"
        message += "inlet triangle indices and centres\n"
        message += "%s" % inlet.triangle_indices
        message += "\n"
        message += "%s" % self.domain.get_centroid_coordinates()[inlet.triangle_indices]
        message += "\n"
        message += "region\n"
        message += "%s" % inlet
        message += "\n"
        message += "=====================================\n"
        return message

    def print_statistics(self):
        print(self.statistics())

    def timestepping_statistics(self):
        message = "---------------------------\n"
        message += "Inlet report for %s:\n" % self.label
        message += "--------------------------\n"
        message += "Q [m^3/s]: %.2f\n" % self.applied_Q
        message += "Total volume [m^3]: %.2f\n" % self.total_applied_volume
        return message

    def print_timestepping_statistics(self):
        print(self.timestepping_statistics())

    print_timestepping_statisitics = print_timestepping_statistics       # (the reference's spelling)

    def __call__(self):
        """inlet_operator.py:78-157; returns the volume added to fractional_step_volume_integral"""
        domain, inlet = self.domain, self.inlet
        timestep = domain.get_timestep()
        t = domain.get_time()
        inlet.fetch()
        current_volume = inlet.get_total_water_volume()
        total_area = inlet.get_area()
        assert current_volume >= 0.0
        Q1 = self.update_Q(t)
        Q2 = self.update_Q(t + timestep)
        Q = 0.5 * (Q1 + Q2)
        volume = Q * timestep
        self.applied_Q = Q
        u, v = inlet.get_velocities()
        added = 0.0
        if volume >= 0.0:
            inlet.set_stages_evenly(volume)
            added = volume
            self.total_requested_volume += volume
            self._set_momenta(u, v)
        elif current_volume + volume >= 0.0:
            depth = (current_volume + volume) / total_area
            inlet.set_depths(depth)
            self.total_requested_volume += volume
            added = volume
            self._set_momenta(u, v)
        else:
            inlet.set_depths(0.0)
            self.total_requested_volume += volume
            volume = -current_volume
            self.applied_Q = -current_volume / timestep
            added = -current_volume
            inlet.set_xmoms(0.0)                     # inlet_operator.py:159-160
            inlet.set_ymoms(0.0)
        self.total_applied_volume += volume          # :166
        inlet.commit()
        ref_op = getattr(self, "_mirror", None)      # attach.py: statistics the reference's reporting reads
        if ref_op is not None:
            ref_op.applied_Q = self.applied_Q
            ref_op.total_requested_volume = self.total_requested_volume
            ref_op.total_applied_volume = self.total_applied_volume
        return added

    def _set_momenta(self, u, v):
        inlet = self.inlet
        depths = inlet.get_depths()
        if self.velocity is not None:
            inlet.set_xmoms(depths * self.velocity[0])
            inlet.set_ymoms(depths * self.velocity[1])
        else:
            inlet.set_xmoms(depths * u)
            inlet.set_ymoms(depths * v)
        if self.zero_velocity:
            inlet.set_xmoms(0.0)
            inlet.set_ymoms(0.0)

    def localise(self, sub):
        """this operator on sub-domain `sub` of a distributed run (called by parallel.distribute)"""
        import copy
        new = copy.copy(self)
        new.domain = sub
        new.inlet = self.inlet.localise(sub)
        sub.set_fractional_step_operator(new)
        return new

    def oracle_spec(self):
        return ("inlet", dict(indices=self.inlet.triangle_indices.copy(), Q=self.Q, velocity=self.velocity,
                              zero_velocity=self.zero_velocity, default=self.default))


# ======================================================================================
# culverts: Inlet_enquiry, Structure_operator, Boyd_box_operator
# ======================================================================================
class Inlet_enquiry(Inlet):
    """An exchange region plus the triangle whose state drives the structure
    (structures/inlet_enquiry.py:10-161)."""

    def __init__(self, domain, region, enquiry_pt, invert_elevation=None, outward_culvert_vector=None,
                 verbose=False):
        if not isinstance(region, Region):
            region = Region(domain, poly=region, expand_polygon=True)      # inlet.py:26-29
        Inlet.__init__(self, domain, region, verbose)
        self.enquiry_pt = enquiry_pt
        self.invert_elevation = invert_elevation
        self.outward_culvert_vector = outward_culvert_vector
        self.enquiry_index = triangle_containing_point(domain, enquiry_pt)
        self.enquiry = None         # (4,) stage, xmom, ymom, elevation of the enquiry triangle
        self._extra_ids = np.array([self.enquiry_index], dtype=np.int64)
        self._extra_rows = np.array([len(self.triangle_indices)], dtype=np.int64)
    _n_extra = 1

    @classmethod
    def from_indices(cls, domain, triangle_indices, enquiry_index, invert_elevation=None,
                     outward_culvert_vector=None):
        """an inlet whose geometry was resolved elsewhere (e.g. by the reference package)"""
        self = cls.__new__(cls)
        Inlet.__init__(self, domain, Region(domain, indices=np.asarray(triangle_indices, dtype=np.int64)))
        self.enquiry_pt = None
        self.invert_elevation = invert_elevation
        self.outward_culvert_vector = None if outward_culvert_vector is None else np.array(outward_culvert_vector)
        self.enquiry_index = int(enquiry_index)
        self.enquiry = None
        self._extra_ids = np.array([self.enquiry_index], dtype=np.int64)
        self._extra_rows = np.array([len(self.triangle_indices)], dtype=np.int64)
        return self

    def localise(self, sub):
        new = Inlet.localise(self, sub)
        hit, local = _owned(sub, [self.enquiry_index])
        new._extra_ids = local
        new._extra_rows = np.array([self._n_rows()], dtype=np.int64)[:len(local)]
        return new

    def _take(self, got):
        self.values = got[:-1]
        self.enquiry = got[-1]

    def get_enquiry_stage(self):
        return self.enquiry[0]

    def get_enquiry_xmom(self):
        return self.enquiry[1]

    def get_enquiry_ymom(self):
        return self.enquiry[2]

    def get_enquiry_elevation(self):
        return self.enquiry[3]

    def get_enquiry_invert_elevation(self):
        return self.get_enquiry_elevation() if self.invert_elevation is None else self.invert_elevation

    def get_enquiry_depth(self):
        return max(self.get_enquiry_stage() - self.get_enquiry_invert_elevation(), 0.0)

    def get_enquiry_water_depth(self):
        return self.get_enquiry_stage() - self.get_enquiry_elevation()

    def get_enquiry_velocity(self):
        depth = self.get_enquiry_water_depth()
        u = depth * self.get_enquiry_xmom() / (depth ** 2 + velocity_protection)
        v = depth * self.get_enquiry_ymom() / (depth ** 2 + velocity_protection)
        return u, v

    def get_enquiry_xvelocity(self):
        return self.get_enquiry_velocity()[0]

    def get_enquiry_yvelocity(self):
        return self.get_enquiry_velocity()[1]

    def get_enquiry_position(self):
        return self.enquiry_pt

    def get_enquiry_speed(self):
        u, v = self.get_enquiry_velocity()
        return math.sqrt(u ** 2 + v ** 2)

    def get_enquiry_velocity_head(self):
        if getattr(self.domain, "use_new_velocity_head", False):
            u, v = self.get_enquiry_velocity()
            n1, n2 = self.outward_culvert_vector
            normal_speed = min(u * n1 + v * n2, 0.0)     # only flow INTO the culvert counts
            return 0.5 * normal_speed ** 2 / g
        return 0.5 * self.get_enquiry_speed() ** 2 / g

    def get_enquiry_total_energy(self):
        return self.get_enquiry_velocity_head() + self.get_enquiry_stage()

    def get_enquiry_specific_energy(self):
        return self.get_enquiry_velocity_head() + self.get_enquiry_depth()


def _unit(v, what):
    length = math.sqrt(float(np.sum(v ** 2)))
    assert length > 0.0, "The length of %s is less than 0" % what
    return v / length, length


class Structure_operator:
    """Base of the culvert family: geometry of the two exchange regions and the semi-implicit
    transfer of `Q` from the inflow to the outflow region (structure_operator.py:13-330).
    Subclasses provide discharge_routine() -> (Q, barrel_speed, outlet_depth) and set
    self.inflow / self.outflow."""
    time_dependent = True
    host_side = True
    counter = 0

    def __init__(self, domain, end_points=None, exchange_lines=None, enquiry_points=None,
                 invert_elevations=None, width=None, height=None, diameter=None, z1=None, z2=None,
                 blockage=None, barrels=None, apron=None, manning=None, enquiry_gap=None,
                 use_momentum_jet=False, zero_outflow_momentum=True, use_old_momentum_method=True,
                 always_use_Q_wetdry_adjustment=True, force_constant_inlet_elevations=False,
                 description=None, label=None, structure_type=None, logging=None, verbose=None):
        sub = None
        if is_distributed(domain):
            # build on the gathered patch around the culvert (identical on every rank), localise below
            sub = domain
            reach = 2.0 * sum(abs(float(v)) for v in (width, height, diameter, apron, enquiry_gap) if v is not None)
            lo, hi = _bbox(end_points, exchange_lines, enquiry_points)
            domain = PatchDomain(sub, lo - reach, hi + reach)
        self._sub = sub
        self.domain = domain
        as_array = lambda a: None if a is None else np.array(a, dtype=np.float64)
        self.end_points = as_array(end_points)
        self.exchange_lines = as_array(exchange_lines)
        self.enquiry_points = as_array(enquiry_points)
        self.invert_elevations = as_array(invert_elevations)
        assert self.end_points is None or self.exchange_lines is None
        self.force_constant_inlet_elevations = force_constant_inlet_elevations
        if height is None:
            height = width
        if width is None:
            width = diameter
        if apron is None:
            apron = width
        assert width is not None
        self.width, self.height, self.diameter = width, height, diameter
        self.z1, self.z2, self.blockage, self.barrels = z1, z2, blockage, barrels
        self.apron, self.manning, self.enquiry_gap = apron, manning, enquiry_gap
        if use_momentum_jet and zero_outflow_momentum:
            raise Exception("Can't have use_momentum_jet and zero_outflow_momentum both True")
        self.use_momentum_jet = use_momentum_jet
        self.zero_outflow_momentum = zero_outflow_momentum
        self.use_old_momentum_method = use_old_momentum_method
        self.always_use_Q_wetdry_adjustment = always_use_Q_wetdry_adjustment
        self.description = " " if description is None else description
        self.label = ("structure" if label is None else label) + "_%g" % Structure_operator.counter
        self.structure_type = "generic structure" if structure_type is None else structure_type
        self.verbose = verbose
        Structure_operator.counter += 1

        self.accumulated_flow = 0.0
        self.discharge = 0.0
        self.discharge_abs_timemean = 0.0
        self.velocity = 0.0
        self.outlet_depth = 0.0
        self.delta_total_energy = 0.0
        self.driving_energy = 0.0

        if self.exchange_lines is not None:
            self._skew_geometry()
        elif self.end_points is not None:
            self._straight_geometry()
        else:
            raise Exception("Define either exchange_lines or end_points")

        self.inlets = []
        for k, outward in ((0, self.outward_vector_0), (1, self.outward_vector_1)):
            line = self.exchange_lines[k]
            if self.apron is None:
                poly = line
            else:                      # the exchange region: the line swept `apron` away from the culvert
                offset = -self.apron * outward
                poly = np.array([line[0], line[1], line[1] + offset, line[0] + offset])
            invert = None if self.invert_elevations is None else self.invert_elevations[k]
            self.inlets.append(Inlet_enquiry(domain, poly, self.enquiry_points[k], invert_elevation=invert,
                                             outward_culvert_vector=outward, verbose=verbose))
            if force_constant_inlet_elevations:
                # one bed level over the exchange region: its area-weighted mean (structure_operator.py:170-173,
                # inlet.py:83-86, 176-178); done on the host arrays, before the state is uploaded
                inlet = self.inlets[-1]
                z = domain.quantities["elevation"]
                ids = inlet.triangle_indices
                level = np.sum(z.centroid_values[ids] * inlet.areas) / inlet.area
                z.centroid_values[ids] = level
                z.host_dirty = True
                if sub is not None:
                    # distributed: the same level on this rank's copies (full and ghost) of those triangles
                    here = np.flatnonzero(np.isin(np.asarray(sub.tri_l2s, dtype=np.int64), domain.seq_ids[ids]))
                    zs = sub.quantities["elevation"]
                    zs.centroid_values[here] = level
                    zs.host_dirty = True
        self.inflow, self.outflow = self.inlets
        if sub is None:
            domain.set_fractional_step_operator(self)
        # (a distributed structure registers itself in _localise_from_patch, once its subclass has primed its
        # smoothing memory on the patch's copy of the initial state)

    def _localise_from_patch(self):
        """second half of the distributed construction: patch indices -> undistributed ids, keep the owned rows"""
        sub, patch = self._sub, self.domain
        if sub is None:
            return
        for inlet in self.inlets:
            _relabel_inlet(inlet, patch)
        self.domain = sub
        self.inlets = [i.localise(sub) for i in self.inlets]
        self.inflow, self.outflow = self.inlets
        self._sub = None
        sub.set_fractional_step_operator(self)

    # -- geometry (structure_operator.py:396-480) -----------------------------------------
    def _straight_geometry(self):
        self.culvert_vector, self.culvert_length = _unit(self.end_points[1] - self.end_points[0], "culvert")
        self.outward_vector_0 = self.culvert_vector
        self.outward_vector_1 = -self.culvert_vector
        normal = np.array([-self.culvert_vector[1], self.culvert_vector[0]])
        half = 0.5 * self.width * normal
        self.exchange_lines = [np.array([self.end_points[i] + half, self.end_points[i] - half]) for i in (0, 1)]
        if self.enquiry_points is None:
            gap = (self.apron + self.enquiry_gap) * self.culvert_vector
            self.enquiry_points = [self.end_points[i] + (2 * i - 1) * gap for i in (0, 1)]

    def _skew_geometry(self):
        lines = self.exchange_lines
        centre0 = 0.5 * (lines[0][0] + lines[0][1])
        centre1 = 0.5 * (lines[1][0] + lines[1][1])
        n0, n1 = len(lines[0]), len(lines[1])
        assert n0 == n1, "There should be the same number of points in both exchange_lines"
        if n0 == 2:
            vec = centre1 - centre0 if self.end_points is None else self.end_points[1] - self.end_points[0]
            out0, out1 = vec, -vec
        elif n0 == 4:
            out0 = lines[0][3] - lines[0][2]
            out1 = lines[1][3] - lines[1][2]
            vec = centre1 - centre0
        else:
            raise Exception("n_exchange_0 != 2 or 4")
        self.culvert_vector, self.culvert_length = _unit(np.array(vec, dtype=np.float64), "culvert")
        self.outward_vector_0, _ = _unit(np.array(out0, dtype=np.float64), "outlet_vector_0")
        self.outward_vector_1, _ = _unit(np.array(out1, dtype=np.float64), "outlet_vector_1")
        if self.enquiry_points is None:
            raise Exception("enquiry_points are required with exchange_lines")

    # -- accessors ---------------------------------------------------------------------
    def get_culvert_length(self):
        return self.culvert_length

    def get_culvert_width(self):
        return self.width

    def get_culvert_height(self):
        return self.height

    def get_culvert_blockage(self):
        return self.blockage

    def get_culvert_barrels(self):
        return self.barrels

    def get_inlets(self):
        return self.inlets

    def get_culvert_diameter(self):
        return self.diameter

    def get_culvert_z1(self):
        return self.z1

    def get_culvert_z2(self):
        return self.z2

    def get_culvert_apron(self):
        return self.apron

    def get_culvert_slope(self):
        a, b = self.inlets
        return (b.get_enquiry_invert_elevation() - a.get_enquiry_invert_elevation()) / self.get_culvert_length()

    def get_master_proc(self):
        return 0

    def _both(self, what):
        return [getattr(i, "get_enquiry_" + what)() for i in self.inlets]

    # the state of the two enquiry triangles as the last call of the structure saw it; refresh() re-reads
    # it from the device (on a distributed domain: called by every rank)
    def get_enquiry_stages(self):
        return self._both("stage")

    def get_enquiry_depths(self):
        return self._both("depth")

    def get_enquiry_positions(self):
        return self._both("position")

    def get_enquiry_xmoms(self):
        return self._both("xmom")

    def get_enquiry_ymoms(self):
        return self._both("ymom")

    def get_enquiry_elevations(self):
        return self._both("elevation")

    def get_enquiry_water_depths(self):
        return self._both("water_depth")

    def get_enquiry_invert_elevations(self):
        return self._both("invert_elevation")

    def get_enquiry_velocitys(self):
        return self._both("velocity")

    def get_enquiry_xvelocitys(self):
        return self._both("xvelocity")

    def get_enquiry_yvelocitys(self):
        return self._both("yvelocity")

    def get_enquiry_speeds(self):
        return self._both("speed")

    def get_enquiry_velocity_heads(self):
        return self._both("velocity_head")

    def get_enquiry_total_energys(self):
        return self._both("total_energy")

    def get_enquiry_specific_energys(self):
        return self._both("specific_energy")

    def refresh(self):
        """re-read both exchange regions and enquiry triangles (device -> host), e.g. before reporting at a yield"""
        if getattr(self.domain, "_dev", None) is not None:
            self._fetch()

    # -- reporting (structure_operator.py:498-640) ------------------------------------------
    def statistics(self):
        message = "=====================================\n"
        message += "Structure Operator: %s\n" % self.label
        message += "=====================================\n"
        message += "Structure Type: %s\n" % self.structure_type
        message += "Description\n"
        message += "%s" % self.description
        if self.structure_type == "boyd_pipe":
            message += "Culvert Diameter: %s\n" % self.diameter
        elif self.structure_type == "boyd_box":
            message += "Culvert  Height: %s\n" % self.height
            message += "Culvert    Width: %s\n" % self.width
        else:
            message += "Culvert Height: %s\n" % self.height
            message += "Culvert  Width: %s\n" % self.width
            message += "Batter Slope 1: %s\n" % self.z1
            message += "Batter Slope 2: %s\n" % self.z2
        message += "Culvert Blockage: %s\n" % self.blockage
        message += "No.  of  barrels: %s\n" % self.barrels
        message += "\n"
        for i, inlet in enumerate(self.inlets):
            message += "-------------------------------------\n"
            message += "Inlet %i\n" % i
            message += "-------------------------------------\n"
            message += "inlet triangle indices and centres and elevations\n"
            message += "%s" % inlet.triangle_indices
            message += "\n"
            message += "%s" % self.domain.get_centroid_coordinates()[inlet.triangle_indices]
            message += "\n"
            elev = self.domain.quantities["elevation"].centroid_values[inlet.triangle_indices]
            message += "%s" % elev
            message += "\n"
            if len(elev) and not np.allclose(elev.max() - elev.min(), 0.0):
                message += "Warning: non-constant inlet elevation can cause well-balancing problems"
            message += "region\n"
            message += "%s" % inlet.region
            message += "\n"
        message += "=====================================\n"
        return message

    def print_statistics(self):
        print(self.statistics())

    def print_timestepping_statistics(self):
        self.refresh()
        message = "---------------------------\n"
        message += "Structure report for %s:\n" % self.label
        message += "--------------------------\n"
        message += "Type: %s\n" % self.structure_type
        for k in (0, 1):
            inlet = self.inlets[k]
            for what, unit in (("depth", "m"), ("speed", "m/s"), ("stage", "m"), ("elevation", "m")):
                message += "inlets[%d]_enquiry_%s [%s]:  %.2f\n" % (k, what, unit, getattr(inlet, "get_enquiry_" + what)())
            for what, unit in (("depth", "m"), ("speed", "m/s"), ("stage", "m"), ("elevation", "m")):
                message += "inlets[%d]_average_%s [%s]:  %.2f\n" % (k, what, unit, getattr(inlet, "get_average_" + what)())
            if k == 0:
                message += "\n"
        message += "Discharge [m^3/s]: %.2f\n" % self.discharge
        message += "Discharge_function_value [m^3/s]: %.2f\n" % self.discharge_abs_timemean
        message += "Velocity  [m/s]: %.2f\n" % self.velocity
        message += "Outlet Depth  [m]: %.2f\n" % self.outlet_depth
        message += "Accumulated Flow [m^3]: %.2f\n" % self.accumulated_flow
        message += "Inlet Driving Energy %.2f\n" % self.driving_energy
        message += "Delta Total Energy %.2f\n" % self.delta_total_energy
        message += "Control at this instant: %s\n" % getattr(self, "case", "N/A")
        print(message)

    def timestepping_statistics(self):
        """one csv row: time, discharge, its time mean since the last call, velocity, accumulated flow,
        driving energy, head difference; the time mean starts again"""
        message = "%.5f, " % self.domain.get_time()
        message += "%.5f, " % self.discharge
        message += "%.5f, " % self.discharge_abs_timemean
        message += "%.5f, " % self.velocity
        message += "%.5f, " % self.accumulated_flow
        message += "%.5f, " % self.driving_energy
        message += "%.5f" % self.delta_total_energy
        self.discharge_abs_timemean = 0.0
        return message

    def set_label(self, label):
        self.label = label

    # used while evolving to close or resize a culvert (structure_operator.py:366-388): the rating reads culvert_*
    def set_culvert_height(self, height):
        self.culvert_height = height

    def set_culvert_width(self, width):
        self.culvert_width = width

    def set_culvert_z1(self, z1):
        self.culvert_z1 = z1

    def set_culvert_z2(self, z2):
        self.culvert_z2 = z2

    def set_culvert_blockage(self, blockage):
        self.culvert_blockage = blockage

    def set_culvert_barrels(self, barrels):
        self.culvert_barrels = barrels

    def discharge_routine(self):
        raise NotImplementedError

    def _fetch(self):
        fetch_inlets(self.inlets)          # both ends in one gather + one cross-rank merge

    def localise(self, sub):
        """this structure on sub-domain `sub` of a distributed run (parallel_structure_operator.py
        restated: see Inlet.localise); the smoothing memory primed on the whole domain is kept"""
        import copy
        new = copy.copy(self)
        new.domain = sub
        new.inlets = [i.localise(sub) for i in self.inlets]
        new.inflow, new.outflow = new.inlets
        sub.set_fractional_step_operator(new)
        return new

    # -- the transfer (structure_operator.py:215-372) --------------------------------------
    def __call__(self):
        timestep = self.domain.get_timestep()
        self._fetch()
        Q, barrel_speed, outlet_depth = self.discharge_routine()
        inflow, outflow = self.inflow, self.outflow
        area_in, area_out = inflow.get_area(), outflow.get_area()

        depth0 = inflow.get_average_depth()
        xmom0 = inflow.get_average_xmom()
        ymom0 = inflow.get_average_ymom()

        dt_Q_on_d = timestep * Q / depth0 if depth0 > 0.0 else 0.0
        rescale = self.always_use_Q_wetdry_adjustment or (depth0 * area_in <= Q * timestep)
        factor = 1.0 / (1.0 + dt_Q_on_d / area_in)
        if rescale:           # Q scaled by new/old depth: the inflow region can never be over-drawn
            depth1 = depth0 * factor
            timestep_star = timestep * depth1 / depth0 if depth0 > 0.0 else 0.0
        else:
            depth1 = depth0 - timestep * Q / area_in
            timestep_star = timestep

        if self.use_old_momentum_method:
            mom_factor = factor
        elif depth0 > 0.0:
            if rescale:
                mom_factor = 1.0 / (1.0 + dt_Q_on_d * depth1 / (depth0 * area_in))
            else:
                mom_factor = 1.0 / (1.0 + timestep * Q / (depth0 * area_in))
        else:
            mom_factor = 0.0
        xmom1 = xmom0 * mom_factor
        ymom1 = ymom0 * mom_factor

        inflow.set_depths(depth1)
        inflow.set_xmoms(xmom1)
        inflow.set_ymoms(ymom1)

        loss = (depth0 - depth1) * area_in
        xmom_loss = (xmom0 - xmom1) * area_in
        ymom_loss = (ymom0 - ymom1) * area_in

        extra_depth = Q * timestep_star / area_out
        direction = -outflow.outward_culvert_vector
        gain = extra_depth * area_out
        assert np.allclose(gain - loss, 0.0)

        self.accumulated_flow += gain
        self.discharge = Q * timestep_star / timestep
        self.discharge_abs_timemean += gain / self.domain.yieldstep
        self.velocity = barrel_speed
        self.outlet_depth = outlet_depth

        out_depth = outflow.get_average_depth() + extra_depth
        outflow.set_depths(out_depth)
        if self.use_momentum_jet:
            out_xmom = barrel_speed * out_depth * direction[0]
            out_ymom = barrel_speed * out_depth * direction[1]
        elif self.zero_outflow_momentum:
            out_xmom = out_ymom = 0.0
        else:
            out_xmom = outflow.get_average_xmom() + xmom_loss / area_out
            out_ymom = outflow.get_average_ymom() + ymom_loss / area_out
        outflow.set_xmoms(out_xmom)
        outflow.set_ymoms(out_ymom)

        inflow.commit()
        outflow.commit()
        return 0.0                    # water is moved, not created: nothing for the volume integral


def boyd_box_function(width, depth, blockage, barrels, flow_width, length, driving_energy,
                      delta_total_energy, outlet_enquiry_depth, sum_loss, manning):
    """Boyd's box-culvert rating: inlet control (weir / orifice), then outlet control when the
    tailwater matters.  Returns Q, barrel velocity, outlet depth, flow area, case
    (boyd_box_operator.py:265-400; constants are Boyd's)."""
    open_fraction = 1 - blockage
    if blockage >= 1.0:
        return 0.0, 0.0, 0.0, 0.00001, "100 blocked culvert"
    span = open_fraction * width * barrels                  # total clear width of all barrels
    Q_weir = 0.544 * g ** 0.5 * open_fraction * width * barrels * driving_energy ** 1.50
    Q_orifice = 0.702 * g ** 0.5 * open_fraction * width * barrels * depth ** 0.89 * driving_energy ** 0.61
    if Q_weir < Q_orifice:
        Q = Q_weir
    else:
        Q = Q_orifice

    def section(Q):
        """flow depth (critical, capped by the barrel), wetted area and perimeter"""
        d = (Q ** 2 / g / span ** 2) ** 0.333333
        if d > depth:
            return depth, span * depth, 2 * (span + depth), True
        return d, span * d, span + 2 * d, False

    outlet_culvert_depth, flow_area, perimeter, full = section(Q)
    case = "Inlet CTRL Outlet unsubmerged PIPE PART FULL" if full else \
        "INLET CTRL Culvert is open channel flow we will for now assume critical depth"

    if delta_total_energy < driving_energy:                 # outlet control can govern
        if outlet_enquiry_depth > depth:
            outlet_culvert_depth = depth
            flow_area = span * depth
            perimeter = 2.0 * (span + depth)
            case = "Outlet submerged"
        else:
            outlet_culvert_depth, flow_area, perimeter, full = section(Q)
            if full:
                perimeter = 2.0 * (span + depth)
            else:
                perimeter = span + 2.0 * outlet_culvert_depth
            case = "Outlet is Flowing Full" if full else "Outlet is open channel flow"
        hyd_rad = flow_area / perimeter
        culvert_velocity = math.sqrt(delta_total_energy /
                                     ((sum_loss / 2 / g) + (manning ** 2 * length) / hyd_rad ** 1.33333))
        Q = min(Q, flow_area * culvert_velocity)

    barrel_velocity = Q / (flow_area + velocity_protection / flow_area)
    return Q, barrel_velocity, outlet_culvert_depth, flow_area, case


def total_energy(smooth_delta_total_energy, delta_total_energy, timestep, smoothing_timescale,
                 forward_Euler_smooth=True):
    """exponential smoothing of the head difference (boyd_box_operator.py:402-418)"""
    if forward_Euler_smooth:
        ts = timestep / max(timestep, smoothing_timescale, 1.0e-06) if timestep > 0.0 else 1.0
        smoothed = smooth_delta_total_energy + ts * (delta_total_energy - smooth_delta_total_energy)
    else:
        ts = timestep / max(smoothing_timescale, 1.0e-06)
        smoothed = (smooth_delta_total_energy + ts * delta_total_energy) / (1.0 + ts)
    return smoothed, ts


def smooth_discharge(smooth_delta_total_energy, smooth_Q, Q, flow_area, ts, forward_Euler_smooth=True):
    """signed, smoothed discharge; a sign flip shuts the culvert for the step
    (boyd_box_operator.py:420-441)"""
    sign = np.sign(smooth_delta_total_energy)
    if forward_Euler_smooth:
        smooth_Q = smooth_Q + ts * (Q * sign - smooth_Q)
    else:
        smooth_Q = (smooth_Q + ts * (Q * sign)) / (1.0 + ts)
    Q = 0.0 if np.sign(smooth_Q) != sign else min(abs(smooth_Q), Q)
    barrel_velocity = 0.0 if flow_area == 0 else Q / flow_area
    return smooth_Q, Q, barrel_velocity


class _Boyd_operator(Structure_operator):
    """What Boyd_box_operator and Boyd_pipe_operator share (boyd_box_operator.py:96-262,
    boyd_pipe_operator.py:78-196): losses, the smoothed head difference that decides the flow
    direction, the rating of the barrel (`_rating`, per subclass) and the smoothed discharge."""

    def _init_boyd(self, losses, use_momentum_jet, use_velocity_head, smoothing_timescale):
        if isinstance(losses, dict):
            self.sum_loss = sum(losses.values())
        elif isinstance(losses, list):
            self.sum_loss = sum(losses)
        else:
            self.sum_loss = losses
        self.use_momentum_jet = use_momentum_jet
        self.zero_outflow_momentum = not use_momentum_jet
        self.use_velocity_head = use_velocity_head
        self.culvert_length = self.get_culvert_length()
        self.culvert_width = self.get_culvert_width()
        self.culvert_height = self.get_culvert_height()
        self.culvert_diameter = self.diameter
        self.culvert_blockage = self.get_culvert_blockage()
        self.culvert_barrels = self.get_culvert_barrels()
        self.max_velocity = 10.0
        self.case = "N/A"
        # one evaluation on the initial state primes the smoothing memory (:118-126)
        self.smoothing_timescale = 0.0
        self.smooth_delta_total_energy = 0.0
        self.smooth_Q = 0.0
        self._fetch_initial()
        Qvd = self.discharge_routine()
        self.smooth_delta_total_energy = 1.0 * self.delta_total_energy
        self.smooth_Q = Qvd[0]
        self.smoothing_timescale = smoothing_timescale
        self._localise_from_patch()          # distributed construction: now take the owned rows

    def _fetch_initial(self):
        """the constructor runs before the device handle exists: read the host arrays"""
        q = self.domain.quantities
        for inlet in self.inlets:
            ids = np.append(inlet.triangle_indices, inlet.enquiry_index).astype(np.int64)
            both = np.stack([q["stage"].centroid_values[ids], q["xmomentum"].centroid_values[ids],
                             q["ymomentum"].centroid_values[ids], q["elevation"].centroid_values[ids]], axis=1)
            inlet.values, inlet.enquiry = both[:-1], both[-1]

    def _blocked(self):
        raise NotImplementedError

    def _rating(self):
        """-> Q, barrel velocity, outlet depth, flow area, case for the current driving_energy,
        delta_total_energy and outflow enquiry depth"""
        raise NotImplementedError

    def discharge_routine(self):
        if self._blocked():
            self.case = "Culvert blocked"
            self.inflow, self.outflow = self.inlets
            return 0.0, 0.0, 0.0
        a, b = self.inlets
        if self.use_velocity_head:
            self.delta_total_energy = a.get_enquiry_total_energy() - b.get_enquiry_total_energy()
        else:
            self.delta_total_energy = a.get_enquiry_stage() - b.get_enquiry_stage()
        self.smooth_delta_total_energy, ts = total_energy(self.smooth_delta_total_energy, self.delta_total_energy,
                                                          self.domain.timestep, self.smoothing_timescale, True)
        if self.smooth_delta_total_energy >= 0.0:
            self.inflow, self.outflow = a, b
            self.delta_total_energy = self.smooth_delta_total_energy
        else:
            self.inflow, self.outflow = b, a
            self.delta_total_energy = -self.smooth_delta_total_energy

        flow_area = None
        if self.inflow.get_enquiry_depth() > 0.01:
            assert self.inflow.get_enquiry_specific_energy() >= 0.0, "Specific energy at inlet is negative"
            if self.use_velocity_head:
                self.driving_energy = self.inflow.get_enquiry_specific_energy()
            else:
                self.driving_energy = self.inflow.get_enquiry_depth()
            Q, barrel_velocity, outlet_culvert_depth, flow_area, case = self._rating()
            self.smooth_Q, Q, barrel_velocity = smooth_discharge(self.smooth_delta_total_energy, self.smooth_Q,
                                                                 Q, flow_area, ts, True)
        else:
            Q = barrel_velocity = outlet_culvert_depth = 0.0
            case = "Inlet dry"
        self.case = case
        if barrel_velocity > self.max_velocity:
            barrel_velocity = self.max_velocity
            Q = flow_area * barrel_velocity
        return Q, barrel_velocity, outlet_culvert_depth

    _MIRRORED = ("accumulated_flow", "discharge", "discharge_abs_timemean", "velocity", "outlet_depth",
                 "delta_total_energy", "driving_energy", "case", "smooth_delta_total_energy", "smooth_Q")

    @classmethod
    def adopt(cls, domain, ref_op):
        """This structure for a Boyd operator object of the reference package (attach.py): its resolved
        geometry, parameters and smoothing memory are taken over, and the statistics the reference's
        reporting reads are written back to it after every call."""
        self = cls.__new__(cls)
        self.domain = domain
        for k in ("width", "height", "diameter", "z1", "z2", "blockage", "barrels", "apron", "manning", "enquiry_gap",
                  "use_momentum_jet", "zero_outflow_momentum", "use_old_momentum_method",
                  "always_use_Q_wetdry_adjustment", "use_velocity_head", "sum_loss", "culvert_length",
                  "max_velocity", "smoothing_timescale", "description", "label", "structure_type") + cls._MIRRORED:
            setattr(self, k, getattr(ref_op, k, None))
        self.culvert_width, self.culvert_height = self.width, self.height
        self.culvert_diameter = self.diameter
        self.culvert_z1, self.culvert_z2 = self.z1, self.z2
        self.culvert_blockage, self.culvert_barrels = self.blockage, self.barrels
        self.inlets = [Inlet_enquiry.from_indices(domain, i.triangle_indices, i.enquiry_index, i.invert_elevation,
                                                  i.outward_culvert_vector) for i in ref_op.inlets]
        self.inflow, self.outflow = self.inlets
        self._mirror = ref_op
        domain.set_fractional_step_operator(self)
        return self

    def __call__(self):
        out = Structure_operator.__call__(self)
        ref_op = getattr(self, "_mirror", None)
        if ref_op is not None:
            for k in self._MIRRORED:
                setattr(ref_op, k, getattr(self, k))
        return out

    _oracle_kind = None

    def oracle_spec(self):
        return (self._oracle_kind, dict(
            inlet_indices=[i.triangle_indices.copy() for i in self.inlets],
            enquiry_indices=[i.enquiry_index for i in self.inlets],
            invert_elevations=[i.invert_elevation for i in self.inlets],
            outward_vectors=[np.array(i.outward_culvert_vector) for i in self.inlets],
            width=self.culvert_width, height=self.culvert_height, diameter=self.culvert_diameter,
            blockage=self.culvert_blockage,
            barrels=self.culvert_barrels, length=self.culvert_length, sum_loss=self.sum_loss,
            manning=self.manning, use_velocity_head=self.use_velocity_head,
            use_momentum_jet=self.use_momentum_jet, zero_outflow_momentum=self.zero_outflow_momentum,
            use_old_momentum_method=self.use_old_momentum_method,
            always_use_Q_wetdry_adjustment=self.always_use_Q_wetdry_adjustment,
            smoothing_timescale=self.smoothing_timescale, max_velocity=self.max_velocity,
            smooth_delta_total_energy=self.smooth_delta_total_energy, smooth_Q=self.smooth_Q,
            use_new_velocity_head=getattr(self.domain, "use_new_velocity_head", False)))


class Boyd_box_operator(_Boyd_operator):
    """anuga.Boyd_box_operator(domain, losses, width, height=None, barrels=1.0, blockage=0.0, ...,
    end_points | exchange_lines, enquiry_points, invert_elevations, apron, manning, enquiry_gap,
    smoothing_timescale, use_momentum_jet, use_velocity_head)  (boyd_box_operator.py:8-262)"""
    _oracle_kind = "boyd_box"

    def __init__(self, domain, losses, width, height=None, barrels=1.0, blockage=0.0, z1=0.0, z2=0.0,
                 end_points=None, exchange_lines=None, enquiry_points=None, invert_elevations=None,
                 apron=0.1, manning=0.013, enquiry_gap=0.0, smoothing_timescale=0.0,
                 use_momentum_jet=True, use_velocity_head=True, description=None, label=None,
                 structure_type="boyd_box", logging=False, verbose=False):
        Structure_operator.__init__(self, domain, end_points=end_points, exchange_lines=exchange_lines,
                                    enquiry_points=enquiry_points, invert_elevations=invert_elevations,
                                    width=width, height=height, blockage=blockage, barrels=barrels,
                                    diameter=None, apron=apron, manning=manning, enquiry_gap=enquiry_gap,
                                    description=description, label=label, structure_type=structure_type,
                                    logging=logging, verbose=verbose)
        self._init_boyd(losses, use_momentum_jet, use_velocity_head, smoothing_timescale)

    def _blocked(self):
        return self.culvert_height <= 0.0

    def _rating(self):
        return boyd_box_function(
            width=self.culvert_width, depth=self.culvert_height, blockage=self.culvert_blockage,
            barrels=self.culvert_barrels, flow_width=self.culvert_width, length=self.culvert_length,
            driving_energy=self.driving_energy, delta_total_energy=self.delta_total_energy,
            outlet_enquiry_depth=self.outflow.get_enquiry_depth(), sum_loss=self.sum_loss,
            manning=self.manning)


def boyd_pipe_function(depth, diameter, blockage, barrels, length, driving_energy, delta_total_energy,
                       outlet_enquiry_depth, sum_loss, manning):
    """Boyd's circular-culvert rating (boyd_pipe_operator.py:199-372): inlet control (unsubmerged /
    submerged), section properties of a part-full circle, then the barrel's energy loss caps the
    discharge.  Returns Q, barrel velocity, outlet depth, flow area, case."""
    if blockage >= 1.0:
        return 0.0, 0.0, 0.0, 0.00001, "100 blocked culvert"
    if blockage > 0.9:
        bf = 3.333 - 3.333 * blockage
    else:
        bf = 1.0 - 0.4012316798 * blockage - 0.3768350138 * (blockage ** 2)
    D = bf * diameter                                       # clear diameter
    Q_unsub = barrels * (0.421 * g ** 0.5 * (D ** 0.87) * driving_energy ** 1.63)
    Q_sub = barrels * (0.530 * g ** 0.5 * (D ** 1.87) * driving_energy ** 0.63)
    Q = min(Q_unsub, Q_sub)

    def critical_depth(Q):
        d1 = D / 1.26 * (Q / g ** 0.5 * (D ** 2.5)) ** (1 / 3.75)
        d2 = D / 0.95 * (Q / g ** 0.5 * (D ** 2.5)) ** (1 / 1.95)
        return d2 if d1 / D > 0.85 else d1

    def full():
        return D, barrels * (D / 2) ** 2 * math.pi, barrels * bf * diameter * math.pi, barrels * bf * diameter

    def part_full(d):
        alpha = math.acos(1 - 2 * d / D) * 2
        return (d, barrels * D ** 2 / 8 * (alpha - math.sin(alpha)), barrels * (alpha * bf * diameter / 2.0),
                barrels * bf * diameter * math.sin(alpha / 2.0))

    d = critical_depth(Q)
    if d >= D:
        outlet_culvert_depth, flow_area, perimeter, flow_width = full()
        case = "Inlet CTRL Outlet submerged Circular PIPE FULL"
    else:
        outlet_culvert_depth, flow_area, perimeter, flow_width = part_full(d)
        case = "INLET CTRL Culvert is open channel flow we will for now assume critical depth"

    if delta_total_energy < driving_energy:                 # outlet control
        if outlet_enquiry_depth > D:
            outlet_culvert_depth, flow_area, perimeter, flow_width = full()
            case = "Outlet submerged"
        else:
            d = critical_depth(Q)
            if d > D:
                outlet_culvert_depth, flow_area, perimeter, flow_width = full()
                case = "Outlet unsubmerged PIPE FULL"
            else:
                outlet_culvert_depth, flow_area, perimeter, flow_width = part_full(d)
                perimeter = barrels * alpha_of(d, D) * bf * diameter / 2.0
                case = "Outlet is open channel flow we will for now assume critical depth"
    hyd_rad = flow_area / perimeter
    culvert_velocity = math.sqrt(delta_total_energy /
                                 ((sum_loss / 2 / g) + (manning ** 2 * length) / hyd_rad ** 1.33333))
    Q = min(Q, flow_area * culvert_velocity)
    barrel_velocity = Q / (flow_area + velocity_protection / flow_area)
    return Q, barrel_velocity, outlet_culvert_depth, flow_area, case


def alpha_of(d, D):
    """angle subtended by the free surface of depth d in a circle of diameter D"""
    return math.acos(1 - 2 * d / D) * 2


class Boyd_pipe_operator(_Boyd_operator):
    """anuga.Boyd_pipe_operator(domain, losses, diameter=None, barrels=1.0, blockage=0.0, ...)
    (boyd_pipe_operator.py:7-196)"""
    _oracle_kind = "boyd_pipe"

    def __init__(self, domain, losses, diameter=None, barrels=1.0, blockage=0.0, z1=0.0, z2=0.0,
                 end_points=None, exchange_lines=None, enquiry_points=None, invert_elevations=None,
                 apron=0.1, manning=0.013, enquiry_gap=0.2, smoothing_timescale=0.0,
                 use_momentum_jet=True, use_velocity_head=True, description=None, label=None,
                 structure_type="boyd_pipe", logging=False, verbose=False):
        Structure_operator.__init__(self, domain, end_points=end_points, exchange_lines=exchange_lines,
                                    enquiry_points=enquiry_points, invert_elevations=invert_elevations,
                                    width=None, height=None, diameter=diameter, blockage=blockage,
                                    barrels=barrels, apron=apron, manning=manning, enquiry_gap=enquiry_gap,
                                    description=description, label=label, structure_type=structure_type,
                                    logging=logging, verbose=verbose)
        self._init_boyd(losses, use_momentum_jet, use_velocity_head, smoothing_timescale)

    def _blocked(self):
        return self.culvert_diameter <= 0.0

    def _rating(self):
        return boyd_pipe_function(
            depth=self.inflow.get_enquiry_depth(), diameter=self.culvert_diameter,
            blockage=self.culvert_blockage, barrels=self.culvert_barrels, length=self.culvert_length,
            driving_energy=self.driving_energy, delta_total_energy=self.delta_total_energy,
            outlet_enquiry_depth=self.outflow.get_enquiry_depth(), sum_loss=self.sum_loss,
            manning=self.manning)


def weir_orifice_trapezoid_function(width, depth, blockage, barrels, z1, z2, flow_width, length, driving_energy,
                                    delta_total_energy, outlet_enquiry_depth, sum_loss, manning):
    """Rating of a trapezoidal opening (bottom width `width`, side slopes z1, z2): weir flow when the inlet
    is unsubmerged, orifice flow when submerged, critical depth of the trapezoid by Newton iteration, then
    the barrel's energy loss when the tailwater matters.  Returns Q, barrel velocity, outlet depth, flow
    area, case (weir_orifice_trapezoid_operator.py:279-485)."""
    bf = 1 - blockage
    if blockage >= 1.0:
        return 0.0, 0.0, 0.0, 0.00001, "100% blocked culvert"
    Q_weir = 1.7 * bf * barrels * ((2 * width + depth * (z1 + z2)) / 2) * driving_energy ** 1.50
    Q_orifice = 0.8 * bf * barrels * g ** 0.5 * (0.5 * depth * (2 * width + depth * (z1 + z2))) * driving_energy ** 0.5
    Q = Q_weir if Q_weir < Q_orifice else Q_orifice

    def critical_depth(Q):
        """Ac^1.5 / Tc^0.5 = Q / sqrt(9.81), Newton from a thin film"""
        dcrit, dyc = 0.00001, 0.001
        while abs(dyc) > 0.00001:
            Tc = bf * barrels * width + (z1 + z2) * dcrit
            Ac = 0.5 * dcrit * (bf * barrels * width + Tc)
            fc = Ac ** 1.5 * Tc ** -0.5 - Q / (9.81 ** 0.5)
            ffc = Ac ** 1.5 * -0.5 * Tc ** -1.5 * (z1 + z2) + Tc ** -0.5 * 1.5 * Ac ** 0.5 * Tc
            dyc = -fc / ffc
            dcrit = dcrit + dyc
        return dcrit

    def area(d):
        return bf * barrels * width * d + 0.5 * (z1 + z2) * d ** 2

    def sides(d):
        return (d ** 2 + (z1 * d) ** 2) ** 0.5, (d ** 2 + (z2 * d) ** 2) ** 0.5

    # inlet control: depth in the barrel = critical depth, capped by the opening
    outlet_culvert_depth = critical_depth(Q)
    d = depth if outlet_culvert_depth > depth else outlet_culvert_depth
    if outlet_culvert_depth > depth:
        outlet_culvert_depth = depth
        case = "Inlet CTRL Outlet unsubmerged PIPE PART FULL"
    else:
        case = "INLET CTRL Culvert is open channel flow we will for now assume critical depth"
    s1, s2 = sides(d)
    flow_area = area(d)
    perimeter = 2.0 * bf * barrels * width + (z1 + z2) * d + s1 + s2
    hyd_rad = flow_area / perimeter
    culvert_velocity = math.sqrt(delta_total_energy /
                                 ((sum_loss / 2 / g) + (manning ** 2 * length) / hyd_rad ** 1.33333))
    Q_outlet_tailwater = flow_area * culvert_velocity

    if delta_total_energy < driving_energy:                 # outlet control
        if outlet_enquiry_depth > depth:
            outlet_culvert_depth = d = depth
            case = "Outlet submerged"
        else:
            Q = min(Q, Q_outlet_tailwater)
            outlet_culvert_depth = critical_depth(Q)
            if outlet_culvert_depth > depth:
                outlet_culvert_depth = d = depth
                case = "Outlet is Flowing Full"
            else:
                d = outlet_culvert_depth
                case = "Outlet is open channel flow"
        s1, s2 = sides(d)
        flow_area = area(d)
        perimeter = bf * barrels * width + s1 + s2
        hyd_rad = flow_area / perimeter
        culvert_velocity = math.sqrt(delta_total_energy /
                                     ((sum_loss / 2 / g) + (manning ** 2 * length) / hyd_rad ** 1.33333))
        Q = min(Q, flow_area * culvert_velocity)
    barrel_velocity = Q / (flow_area + velocity_protection / flow_area)
    return Q, barrel_velocity, outlet_culvert_depth, flow_area, case


class Weir_orifice_trapezoid_operator(_Boyd_operator):
    """anuga.Weir_orifice_trapezoid_operator(domain, losses, width, height=None, barrels=1.0, blockage=0.0,
    z1=0.0, z2=0.0, ...)  (weir_orifice_trapezoid_operator.py:11-278)"""
    _oracle_kind = "weir_orifice_trapezoid"

    def __init__(self, domain, losses, width, height=None, barrels=1.0, blockage=0.0, z1=0.0, z2=0.0,
                 end_points=None, exchange_lines=None, enquiry_points=None, invert_elevations=None,
                 apron=0.1, manning=0.013, enquiry_gap=0.0, smoothing_timescale=0.0,
                 use_momentum_jet=True, use_velocity_head=True, description=None, label=None,
                 structure_type="weir_orifice_trapezoid", logging=False, verbose=False):
        Structure_operator.__init__(self, domain, end_points=end_points, exchange_lines=exchange_lines,
                                    enquiry_points=enquiry_points, invert_elevations=invert_elevations,
                                    width=width, height=height, blockage=blockage, barrels=barrels,
                                    diameter=None, z1=z1, z2=z2, apron=apron, manning=manning,
                                    enquiry_gap=enquiry_gap, description=description, label=label,
                                    structure_type=structure_type, logging=logging, verbose=verbose)
        self.culvert_z1, self.culvert_z2 = self.z1, self.z2
        self._init_boyd(losses, use_momentum_jet, use_velocity_head, smoothing_timescale)

    def _blocked(self):
        return self.culvert_height <= 0.0

    def _rating(self):
        return weir_orifice_trapezoid_function(
            width=self.culvert_width, depth=self.culvert_height, blockage=self.culvert_blockage,
            barrels=self.culvert_barrels, z1=self.culvert_z1, z2=self.culvert_z2, flow_width=self.culvert_width,
            length=self.culvert_length, driving_energy=self.driving_energy,
            delta_total_energy=self.delta_total_energy, outlet_enquiry_depth=self.outflow.get_enquiry_depth(),
            sum_loss=self.sum_loss, manning=self.manning)

    def oracle_spec(self):
        kind, spec = _Boyd_operator.oracle_spec(self)
        spec.update(z1=self.culvert_z1, z2=self.culvert_z2)
        return kind, spec


class Internal_boundary_operator(Structure_operator):
    """anuga.Internal_boundary_operator(domain, internal_boundary_function, width, height, end_points |
    exchange_lines, enquiry_points, invert_elevation, apron, enquiry_gap, use_velocity_head,
    zero_outflow_momentum, force_constant_inlet_elevations, smoothing_timescale,
    compute_discharge_implicitly)  (structures/internal_boundary_operator.py:15-308).

    Q = internal_boundary_function(hw, tw), hw / tw the stage (or total energy) at enquiry points 0 / 1,
    positive from 0 to 1; the transfer is Structure_operator's, without momentum jet.  The implicit form
    solves for the change of both levels over the timestep (scipy.optimize.root, 'lm'), as the reference."""

    def __init__(self, domain, internal_boundary_function, width=1.0, height=1.0, end_points=None,
                 exchange_lines=None, enquiry_points=None, invert_elevation=None, apron=0.0, enquiry_gap=0.0,
                 use_velocity_head=False, zero_outflow_momentum=False, force_constant_inlet_elevations=True,
                 smoothing_timescale=0.0, compute_discharge_implicitly=True, description=None, label=None,
                 structure_type="internal_boundary", logging=False, verbose=True):
        Structure_operator.__init__(self, domain, end_points=end_points, exchange_lines=exchange_lines,
                                    enquiry_points=enquiry_points,
                                    invert_elevations=[invert_elevation, invert_elevation],
                                    width=width, height=height, diameter=None, apron=apron, manning=None,
                                    enquiry_gap=enquiry_gap, use_momentum_jet=False,
                                    zero_outflow_momentum=zero_outflow_momentum, use_old_momentum_method=False,
                                    always_use_Q_wetdry_adjustment=False,
                                    force_constant_inlet_elevations=force_constant_inlet_elevations,
                                    description=description, label=label, structure_type=structure_type,
                                    logging=logging, verbose=verbose)
        self.internal_boundary_function = internal_boundary_function
        self.use_velocity_head = use_velocity_head
        self.compute_discharge_implicitly = compute_discharge_implicitly
        self.max_velocity = 99999999999.0
        self.case = "N/A"
        # one evaluation on the initial state primes the smoothed discharge (:107-116)
        self.smoothing_timescale = 0.0
        self.smooth_Q = 0.0
        self.smooth_delta_total_energy = 0.0
        _Boyd_operator._fetch_initial(self)
        Qvd = self.discharge_routine()
        self.smooth_Q = Qvd[0]
        self.smoothing_timescale = smoothing_timescale
        self._localise_from_patch()

    def _levels(self):
        a, b = self.inlets
        if self.use_velocity_head:
            self.inlet0_energy, self.inlet1_energy = a.get_enquiry_total_energy(), b.get_enquiry_total_energy()
        else:
            self.inlet0_energy, self.inlet1_energy = a.get_enquiry_stage(), b.get_enquiry_stage()
        self.driving_energy = max(self.inlet0_energy, self.inlet1_energy)
        self.delta_total_energy = self.inlet0_energy - self.inlet1_energy

    def discharge_routine(self):
        if self.compute_discharge_implicitly:
            return self.discharge_routine_implicit()
        return self.discharge_routine_explicit()

    def discharge_routine_explicit(self):
        """:136-218"""
        if self.height <= 0.0:
            self.case = "Structure is blocked"
            self.inflow, self.outflow = self.inlets
            return 0.0, 0.0, 0.0
        self._levels()
        dt = self.domain.timestep
        ts = dt / max(dt, self.smoothing_timescale, 1.0e-30) if dt > 0.0 else 1.0
        self.smooth_delta_total_energy += ts * (self.delta_total_energy - self.smooth_delta_total_energy)
        if np.sign(self.smooth_delta_total_energy) != np.sign(self.delta_total_energy):
            self.smooth_delta_total_energy = 0.0
        if self.inlet0_energy >= self.inlet1_energy:
            hw = 1.0 * self.inlet0_energy
            tw = hw - self.smooth_delta_total_energy
        else:
            tw = 1.0 * self.inlet1_energy
            hw = tw + self.smooth_delta_total_energy
        Q = self.internal_boundary_function(hw, tw)
        self.smooth_Q = self.smooth_Q + ts * (Q - self.smooth_Q)
        if self.smooth_Q >= 0.0:
            self.inflow, self.outflow = self.inlets
        else:
            self.outflow, self.inflow = self.inlets
        if np.sign(self.smooth_Q) != np.sign(Q):
            Q = 0.0
        else:
            Q = min(abs(self.smooth_Q), abs(Q))
        return Q, np.nan, np.nan

    def discharge_routine_implicit(self):
        """:221-308: Q(H0 + dH, T0 + dT) with (dH, dT) the level changes that discharge causes over dt"""
        import scipy.optimize as sco
        self._levels()
        f = self.internal_boundary_function
        Q0 = f(self.inlet0_energy, self.inlet1_energy)
        dt = self.domain.get_timestep()
        if dt > 0.0:
            E0, E1 = self.inlet0_energy, self.inlet1_energy
            areas = np.array([self.inlets[0].get_area(), self.inlets[1].get_area()])
            theta = 1.0
            sign = np.array([-1.0, 1.0])

            def residual(sol):
                discharge = (1.0 - theta) * Q0 + theta * f(E0 + sol[0], E1 + sol[1])
                return sol * areas - discharge * dt * sign
            sol = sco.root(residual, np.array([0.0, 0.0]), method="lm").x
            Q = (1.0 - theta) * Q0 + theta * f(E0 + sol[0], E1 + sol[1])
            ts = dt / max(dt, self.smoothing_timescale, 1.0e-30)
        else:
            Q = Q0
            ts = 1.0
        self.smooth_Q = self.smooth_Q + ts * (Q - self.smooth_Q)
        if Q >= 0.0:
            self.inflow, self.outflow = self.inlets
        else:
            self.outflow, self.inflow = self.inlets
        if np.sign(self.smooth_Q) != np.sign(Q):
            Q = 0.0
            self.smooth_Q = 0.0
        else:
            Q = min(abs(self.smooth_Q), abs(Q))
        return Q, np.nan, np.nan

    def oracle_spec(self):
        raise NotImplementedError("Internal_boundary_operator carries a user function: no CPU-oracle scenario")


class pumping_station_function:
    """Rate of a pumping station as an internal_boundary_function: pumps ramp up to `pump_capacity` while the
    headwater is above `hw_to_start_pumping` and down to zero below `hw_to_stop_pumping`
    (structures/internal_boundary_functions.py:392-477)."""

    def __init__(self, domain, pump_capacity, hw_to_start_pumping, hw_to_stop_pumping, initial_pump_rate=0.0,
                 pump_rate_of_increase=1.0e+100, pump_rate_of_decrease=1.0e+100, verbose=True):
        self.pump_capacity = pump_capacity
        self.hw_to_start_pumping = hw_to_start_pumping
        self.hw_to_stop_pumping = hw_to_stop_pumping
        self.pump_rate_of_increase = pump_rate_of_increase
        self.pump_rate_of_decrease = pump_rate_of_decrease
        self.domain = domain
        self.last_time_called = domain.get_time()
        self.time = domain.get_time()
        if hw_to_start_pumping < hw_to_stop_pumping:
            raise Exception("hw_to_start_pumping should be >= hw_to_stop_pumping")
        if initial_pump_rate > pump_capacity:
            raise Exception("Initial pump rate is > pump capacity")
        if pump_rate_of_increase < 0.0 or pump_rate_of_decrease < 0.0:
            raise Exception("Pump rates of increase / decrease MUST be non-negative")
        if pump_capacity < 0.0 or initial_pump_rate < 0.0:
            raise Exception("Pump rates cannot be negative")
        self.pump_rate = initial_pump_rate

    def __call__(self, hw_in, tw_in):
        self.time = self.domain.get_time()
        if self.time > self.last_time_called:
            dt = self.time - self.last_time_called
            self.last_time_called = self.time
        else:
            dt = 0.0
            if self.time != self.last_time_called:
                raise Exception("Impossible timestepping: time %s before the last call at %s"
                                % (self.time, self.last_time_called))
        if hw_in < self.hw_to_stop_pumping:
            self.pump_rate = max(0.0, self.pump_rate - dt * self.pump_rate_of_decrease)
        elif hw_in > self.hw_to_start_pumping:
            self.pump_rate = min(self.pump_capacity, self.pump_rate + dt * self.pump_rate_of_increase)
        return self.pump_rate
