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
        self.op_id = None
        domain.set_fractional_step_operator(self)

    @property
    def rate_type(self):          # the reference's attribute (rate_operators.py:95-140)
        if self.rate_callable is not None:
            return "t"
        return "centroid_array" if self.rate_array is not None else "scalar"

    rate_spatial = False

    @property
    def time_dependent(self):
        return self.rate_callable is not None or callable(self.factor)

    def set_rate(self, rate):
        self.rate_callable = None
        self.rate_array = None
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
                raise NotImplementedError("rate(x, y, t) needs a host evaluation over all centroids every "
                                          "step; outside the hot-path scope")
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
        self.factor = factor

    def current_rate(self, t):
        return float(self.rate_callable(t)) if self.rate_callable is not None else self.rate

    def current_factor(self, t):
        return float(self.factor(t)) if callable(self.factor) else float(self.factor)

    def oracle_spec(self):
        rate = self.rate_callable if self.rate_callable is not None else \
            (self.rate_array if self.rate_array is not None else self.rate)
        assert not callable(self.factor)
        return ("rate", dict(rate=rate, factor=float(self.factor), indices=self.indices))
