"""Fractional-step operators that run on the device once per timestep.

Rate_operator mirrors anuga/operators/rate_operators.py:24-269 for the rate
types the hot-path configs use: a scalar, a function of time f(t) (evaluated on
the host each step), or a per-centroid array.  Spatial functions f(x, y[, t]),
Quantities and xarray rates are evaluated once on the host into a centroid
array when time-independent and are otherwise out of scope.
"""
import numpy as np


class Rate_operator:
    def __init__(self, domain, rate=0.0, factor=1.0, region=None, indices=None, polygon=None,
                 center=None, radius=None, default_rate=0.0, description=None, label=None,
                 logging=False, verbose=False, monitor=False):
        if region is not None or polygon is not None or center is not None or radius is not None:
            # base_operator.py:36-50: the region's triangles (centroid inside the polygon / circle)
            from .structures import Region
            if region is None:
                region = Region(domain, indices=indices, polygon=polygon, center=center, radius=radius)
            indices = region.indices
        self.domain = domain
        self.factor = factor
        self.indices = None if indices is None else np.asarray(indices, dtype=np.int64)
        self.rate_callable = None
        self.rate_array = None
        self.rate = 0.0
        self.set_rate(rate)
        # rate_operators.py:483-509: used when a rate time series is asked for a time outside its records
        assert default_rate is None or isinstance(default_rate, (int, float)) or callable(default_rate), \
            "Default_rate must be either None a scalar, or a function of time.\nI got %s." % str(default_rate)
        self.default_rate = default_rate
        self.default_rate_invoked = False
        self.op_id = None
        self.description = description
        self.label = ("rate_operator" if label is None else label) + "_%g" % len(domain.fractional_step_operators)
        self.verbose = verbose
        domain.set_fractional_step_operator(self)

    @property
    def rate_type(self):          # the reference's attribute (rate_operators.py:95-140)
        if self.rate_callable is not None:
            return "t"
        return "centroid_array" if self.rate_array is not None else "scalar"

    rate_spatial = False
    rate_xyt = None

    @property
    def host_side(self):
        return self.rate_xyt is not None

    @property
    def time_dependent(self):
        return self.rate_callable is not None or callable(self.factor) or self.rate_xyt is not None

    def __call__(self):
        """host-side application (spatial-temporal rates only): rate_operators.py:149-269 on gathered rows;
        returns the volume added to full cells"""
        d = self.domain
        t, dt = d.get_time(), d.get_timestep()
        factor = self.current_factor(t)
        ids = np.arange(d.number_of_triangles, dtype=np.int64) if self.indices is None else self.indices
        if len(ids) == 0:
            return 0.0
        rows = d._dev.gather_centroids(ids)              # stage, xmom, ymom, elevation
        c = d.centroid_coordinates
        rate = np.asarray(self.rate_xyt(c[ids, 0], c[ids, 1], t), dtype=np.float64) * np.ones(len(ids))
        if np.all(rate >= 0.0):
            local_rates = factor * dt * rate
            rows[:, 0] = rows[:, 0] + local_rates
        else:
            heights = rows[:, 0] - rows[:, 3]
            local_rates = np.maximum(factor * dt * rate, -heights)
            f = np.where(local_rates < 0.0, (local_rates + heights) / (heights + 1.0e-10), 1.0)
            rows[:, 0] = rows[:, 0] + local_rates
            rows[:, 1] = rows[:, 1] * f
            rows[:, 2] = rows[:, 2] * f
        d._dev.scatter_centroids(ids, rows[:, :3])
        full = d.tri_full_flag[ids] == 1
        return float(np.sum((local_rates * d.areas[ids])[full]))

    def set_rate(self, rate):
        self.rate_callable = None
        self.rate_array = None
        self.rate_xyt = None
        self.rate_input = rate
        if callable(rate):
            import inspect
            nargs = len(inspect.signature(rate).parameters)
            if nargs == 1:
                self.rate_callable = rate
            elif nargs == 2:
                C = self.domain.centroid_coordinates
                self.rate_array = np.asarray(rate(C[:, 0], C[:, 1]), dtype=np.float64) * np.ones(len(C))
            else:
                # rate(x, y, t): evaluated on the host every step over the operator's cells, which are
                # gathered from / scattered to the device (rate_operators.py:165-172, 276-300)
                self.rate_xyt = rate
        elif isinstance(rate, (list, tuple, np.ndarray)):
            self.rate_array = np.asarray(rate, dtype=np.float64)
            assert self.rate_array.shape == (self.domain.number_of_triangles,)
        elif hasattr(rate, "centroid_values"):
            self.rate_array = np.array(rate.centroid_values, dtype=np.float64)
        else:
            self.rate = float(rate)
        dev = getattr(self.domain, "_dev", None)
        if dev is not None and getattr(self, "op_id", None) is not None:
            if self.rate_array is not None:
                raise NotImplementedError("changing an array rate after the device upload")
            dev.set_rate(self.op_id, self.current_rate(self.domain.get_time()), self.current_factor(self.domain.get_time()))

    def set_factor(self, factor):
        """The reference reads self.factor at every call (rate_operators.py:160), so a new factor takes
        effect at the next step: also for an operator that already lives on the device."""
        self.factor = factor
        dev = getattr(self.domain, "_dev", None)
        if dev is not None and getattr(self, "op_id", None) is not None:
            t = self.domain.get_time()
            dev.set_rate(self.op_id, self.current_rate(t), self.current_factor(t))

    def current_rate(self, t):
        if self.rate_callable is None:
            return self.rate
        try:                     # evaluate_temporal_function (utilities/function_utils.py:85-119)
            return float(self.rate_callable(t))
        except BaseException as e:
            if type(e).__name__ not in ("Modeltime_too_late", "Modeltime_too_early") or self.default_rate is None:
                raise
            if not self.default_rate_invoked:
                import warnings
                self.default_rate_invoked = True
                warnings.warn("Using default_rate outside the time interval of the rate function")
            dr = self.default_rate
            return float(dr(t)) if callable(dr) else float(dr)

    def current_factor(self, t):
        return float(self.factor(t)) if callable(self.factor) else float(self.factor)

    def get_factor(self, t=None):
        return self.current_factor(self.domain.get_time() if t is None else t)

    def get_time(self):
        return self.domain.get_time()

    def get_timestep(self):
        return self.domain.get_timestep()

    def set_default_rate(self, default_rate):
        """rate_operators.py:483-509"""
        assert default_rate is None or isinstance(default_rate, (int, float)) or callable(default_rate), \
            "Default_rate must be either None a scalar, or a function of time.\nI got %s." % str(default_rate)
        self.default_rate = default_rate
        self.default_rate_invoked = False

    def _rates_now(self, t=None):
        """the rate [m/s] over the operator's cells at time t (before the factor)"""
        d = self.domain
        t = d.get_time() if t is None else t
        ids = np.arange(d.number_of_triangles, dtype=np.int64) if self.indices is None else self.indices
        if self.rate_xyt is not None:
            c = d.centroid_coordinates
            return ids, np.asarray(self.rate_xyt(c[ids, 0], c[ids, 1], t), dtype=np.float64) * np.ones(len(ids))
        if self.rate_array is not None:
            return ids, self.rate_array[ids]
        return ids, np.full(len(ids), self.current_rate(t))

    def get_Q(self, full_only=True):
        """current overall discharge [m^3/s] = sum(rate * area) * factor (rate_operators.py:444-481)"""
        d = self.domain
        ids, rate = self._rates_now()
        keep = d.tri_full_flag[ids] == 1 if full_only else np.ones(len(ids), dtype=bool)
        return np.sum(d.areas[ids][keep] * rate[keep]) * self.get_factor()

    def statistics(self):
        return "You need to implement operator statistics for your operator"

    def timestepping_statistics(self):
        """label: min and max rate over the cells and the volume of the last timestep (rate_operators.py:598-601)"""
        from .compat import indent
        ids, rate = self._rates_now()
        full = self.domain.tri_full_flag[ids] == 1
        per_second = (rate * self.get_factor())[full]
        lo, hi = (float(np.min(per_second)), float(np.max(per_second))) if len(per_second) else (0.0, 0.0)
        influx = float(np.sum(per_second * self.domain.areas[ids][full]) * self.domain.get_timestep())
        return indent + self.label + ": Min rate = %g m/s, Max rate = %g m/s, Total Q = %g m^3" % (lo, hi, influx)

    def print_statistics(self):
        print(self.statistics())

    def print_timestepping_statistics(self):
        print(self.timestepping_statistics())

    def set_label(self, label=None):
        self.label = label

    def oracle_spec(self):
        if self.rate_xyt is not None:
            assert not callable(self.factor)
            return ("rate", dict(rate_xyt=self.rate_xyt, rate=None, factor=float(self.factor), indices=self.indices))
        rate = self.rate_callable if self.rate_callable is not None else \
            (self.rate_array if self.rate_array is not None else self.rate)
        assert not callable(self.factor)
        return ("rate", dict(rate=rate, factor=float(self.factor), indices=self.indices))


# ----------------------------------------------------------------------------------------
# Set_quantity / Set_stage: assign values over a region (anuga/operators/set_quantity.py:24-140,
# set_stage.py:24-120).  Called by hand they edit the host arrays (uploaded before the next step);
# the *_operator forms run once per timestep on the region's cells gathered from the device.
# ----------------------------------------------------------------------------------------
def _function_type(value):
    if not callable(value):
        return "scalar"
    import inspect
    n = len(inspect.signature(value).parameters)
    return {1: "t", 2: "x,y", 3: "x,y,t"}[n]


class Set_quantity:
    def __init__(self, domain, quantity, value=None, region=None, indices=None, polygon=None, center=None,
                 radius=None, line=None, verbose=False, test_elevation=True, test_stage=True):
        from .structures import Region
        self.domain = domain
        self.region = region if isinstance(region, Region) else \
            Region(domain, indices=indices, polygon=polygon, center=center, radius=radius, line=line)
        self.indices = self.region.indices
        self.quantity = quantity
        assert quantity in domain.quantities, "quantity not found in domain"
        if test_elevation:
            assert quantity != "elevation", "Use Set_elevation to maintain mass continuity"
        if test_stage:
            assert quantity != "stage", "Use Set_stage to maintain non-negative water depth"
        self.set_value(value)
        self.coord_c = domain.centroid_coordinates

    def set_value(self, value=None):
        self.value = value
        self.value_type = _function_type(value)

    def get_value(self, x=None, y=None, t=None):
        if t is None:
            t = self.domain.get_time()
        if self.value_type == "t":
            return self.value(t)
        if self.value_type == "x,y":
            return self.value(x, y)
        if self.value_type == "x,y,t":
            return self.value(x, y, t)
        return float(self.value)

    def _ids(self):
        return slice(None) if self.indices is None else np.asarray(self.indices, dtype=np.int64)

    def _new_values(self, ids):
        return self.get_value(x=self.coord_c[ids, 0], y=self.coord_c[ids, 1])

    def __call__(self):
        if self.indices is not None and len(self.indices) == 0:
            return
        d = self.domain
        d.sync_to_host()
        ids = self._ids()
        q = d.quantities[self.quantity]
        q.centroid_values[ids] = self._new_values(ids)
        q.host_dirty = True


class Set_stage(Set_quantity):
    """stage over a region, never below the bed (set_stage.py:24-120)"""

    def __init__(self, domain, stage=None, indices=None, polygon=None, center=None, radius=None, line=None,
                 verbose=False):
        Set_quantity.__init__(self, domain, "stage", value=stage, indices=indices, polygon=polygon,
                              center=center, radius=radius, line=line, verbose=verbose, test_stage=False)

    def _new_values(self, ids):
        value = Set_quantity._new_values(self, ids)
        return np.maximum(self.domain.quantities["elevation"].centroid_values[ids], value)


class Set_quantity_operator(Set_quantity):
    """Set_quantity applied every timestep (set_quantity_operator.py:12-60): a host-side
    fractional-step operator on the region's cells.  Conserved quantities only."""
    time_dependent = True
    host_side = True
    _COLUMN = {"stage": 0, "xmomentum": 1, "ymomentum": 2}

    def __init__(self, domain, quantity, value=None, region=None, indices=None, polygon=None, center=None,
                 radius=None, line=None, description=None, label=None, logging=False, verbose=False,
                 test_stage=True, test_elevation=True):
        Set_quantity.__init__(self, domain, quantity, value, region=region, indices=indices, polygon=polygon,
                              center=center, radius=radius, line=line, test_stage=test_stage,
                              test_elevation=test_elevation)
        if quantity not in self._COLUMN:
            raise NotImplementedError("Set_quantity_operator on %s: only the conserved quantities live on the device"
                                      % quantity)
        domain.set_fractional_step_operator(self)

    def __call__(self):
        if self.indices is not None and len(self.indices) == 0:
            return 0.0
        d = self.domain
        ids = np.arange(d.number_of_triangles, dtype=np.int64) if self.indices is None \
            else np.asarray(self.indices, dtype=np.int64)
        rows = d._dev.gather_centroids(ids)              # stage, xmom, ymom, elevation
        rows[:, self._COLUMN[self.quantity]] = self.get_value(x=self.coord_c[ids, 0], y=self.coord_c[ids, 1])
        d._dev.scatter_centroids(ids, rows[:, :3])
        return 0.0

    def oracle_spec(self):
        return ("set_quantity", dict(indices=None if self.indices is None else np.asarray(self.indices).copy(),
                                     quantity=self.quantity, value=self.value, value_type=self.value_type))


class Set_stage_operator(Set_quantity_operator):
    """anuga.Set_stage_operator (set_stage_operator.py:20-50): the stage assigned as given"""

    def __init__(self, domain, stage=None, region=None, indices=None, polygon=None, center=None, radius=None,
                 line=None, description=None, label=None, logging=False, verbose=False):
        Set_quantity_operator.__init__(self, domain, "stage", value=stage, region=region, indices=indices,
                                       polygon=polygon, center=center, radius=radius, line=line,
                                       test_stage=False)
    get_stage = Set_quantity.get_value
    set_stage = Set_quantity.set_value


class Set_elevation(Set_quantity):
    """bed elevation over a region with the water depth kept (set_elevation.py:14-150, the
    discontinuous-elevation branch: elevation and stage move together at the centroids)"""

    def __init__(self, domain, elevation=None, region=None, indices=None, polygon=None, center=None,
                 radius=None, line=None, verbose=False):
        Set_quantity.__init__(self, domain, "elevation", value=elevation, region=region, indices=indices,
                              polygon=polygon, center=center, radius=radius, line=line, verbose=verbose,
                              test_elevation=False)

    def __call__(self):
        if self.value is None or (self.indices is not None and len(self.indices) == 0):
            return
        d = self.domain
        d.sync_to_host()
        ids = self._ids()
        w, z = d.quantities["stage"], d.quantities["elevation"]
        height = w.centroid_values[ids] - z.centroid_values[ids]
        z.centroid_values[ids] = self.get_value(x=self.coord_c[ids, 0], y=self.coord_c[ids, 1])
        w.centroid_values[ids] = z.centroid_values[ids] + height
        w.host_dirty = z.host_dirty = True


class Set_elevation_operator(Set_elevation):
    """Set_elevation applied every timestep (set_elevation_operator.py): erosion, breaches, slides"""
    time_dependent = True
    host_side = True

    def __init__(self, domain, elevation=None, region=None, indices=None, polygon=None, center=None,
                 radius=None, line=None, description=None, label=None, logging=False, verbose=False):
        Set_elevation.__init__(self, domain, elevation, region, indices, polygon, center, radius, line, verbose)
        domain.set_fractional_step_operator(self)

    def __call__(self):
        if self.value is None or (self.indices is not None and len(self.indices) == 0):
            return 0.0
        d = self.domain
        ids = np.arange(d.number_of_triangles, dtype=np.int64) if self.indices is None \
            else np.asarray(self.indices, dtype=np.int64)
        rows = d._dev.gather_centroids(ids)              # stage, xmom, ymom, elevation
        height = rows[:, 0] - rows[:, 3]
        bed = np.asarray(self.get_value(x=self.coord_c[ids, 0], y=self.coord_c[ids, 1]), dtype=np.float64) \
            * np.ones(len(ids))
        rows[:, 0] = bed + height
        d._dev.scatter_bed(ids, bed)
        d._dev.scatter_centroids(ids, rows[:, :3])
        d.quantities["elevation"].centroid_values[ids] = bed        # host copy of the (otherwise static) bed
        return 0.0

    def oracle_spec(self):
        return ("set_elevation", dict(indices=None if self.indices is None else np.asarray(self.indices).copy(),
                                      value=self.value, value_type=self.value_type))
