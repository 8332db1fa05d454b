"""Inlet_operator: water added / removed over a small set of triangles each timestep.

Mirrors anuga/structures/inlet.py:11-237 (Inlet: the exchange region and its averages,
set_stages_evenly :192-227) and anuga/structures/inlet_operator.py:9-157 (Inlet_operator.__call__).
The hydraulics are scalar / O(inlet triangles) numpy on the HOST, on values gathered from the
device (swk_gather_centroids) and scattered back (swk_scatter_centroids); the big arrays stay
resident.  SURVEY.md section 8(f) row 1.

Region resolution from polygons / lines is mesh set-up geometry (anuga/geometry), outside the hot
path: a Region here is a list of triangle indices, a circle or a polygon tested on centroids.
"""
import numpy as np

velocity_protection = 1.0e-6        # anuga/config.py:18


class Region:
    """abstract_2d_finite_volumes/region.py:24-150, for indices / center+radius / polygon (centroids)"""

    def __init__(self, domain, indices=None, polygon=None, center=None, radius=None, **unsupported):
        for k, v in unsupported.items():
            if v not in (None, False):
                raise NotImplementedError("Region(%s=...) is set-up geometry outside the hot path" % k)
        self.domain = domain
        c = domain.centroid_coordinates
        if indices is not None:
            self.indices = np.asarray(indices, dtype=np.int64)
            self.type = "user_defined"
        elif center is not None and radius is not None:
            d2 = (c[:, 0] - center[0]) ** 2 + (c[:, 1] - center[1]) ** 2
            self.indices = np.flatnonzero(d2 < radius ** 2).astype(np.int64)
            self.type = "circle"
        elif polygon is not None:
            self.indices = np.flatnonzero(_inside_polygon(c, np.asarray(polygon, dtype=np.float64))).astype(np.int64)
            self.type = "polygon"
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


class Inlet:
    """The exchange region: views of the inlet triangles' centroid values, held on the host for the
    duration of one operator call (structures/inlet.py:11-237)."""

    def __init__(self, domain, region, verbose=False):
        self.domain = domain
        self.region = region if isinstance(region, Region) else Region(domain, indices=region)
        self.triangle_indices = np.asarray(self.region.indices, dtype=np.int64)
        if len(self.triangle_indices) == 0:
            raise Exception("No triangles have been identified in region ")
        self.areas = domain.areas[self.triangle_indices]
        self.area = float(np.sum(self.areas))
        assert self.area > 0.0
        self.values = None          # (n,4) stage, xmom, ymom, elevation; loaded by fetch()

    # -- device exchange ---------------------------------------------------------------
    def fetch(self):
        self.values = self.domain._dev.gather_centroids(self.triangle_indices)

    def commit(self):
        self.domain._dev.scatter_centroids(self.triangle_indices, self.values[:, :3])

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

    def get_velocities(self):
        depths = self.get_depths()
        u = self.get_xmoms() * depths / (depths * depths + velocity_protection)
        v = self.get_ymoms() * depths / (depths * depths + velocity_protection)
        return u, v

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
        return self.applied_Q

    def get_inlet(self):
        return self.inlet

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
        inlet.commit()
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

    def oracle_spec(self):
        return ("inlet", dict(indices=self.inlet.triangle_indices.copy(), Q=self.Q, velocity=self.velocity,
                              zero_velocity=self.zero_velocity, default=self.default))
