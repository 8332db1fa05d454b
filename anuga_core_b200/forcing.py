"""Explicit forcing terms (anuga/shallow_water/forcing.py).

``domain.forcing_terms`` is the reference's list of callables ``f(domain)`` that
``compute_forcing_terms`` runs after every flux evaluation (generic_domain.py:2417-2427).  Here the
list holds descriptors: Manning friction is built into the update kernels (the placeholder below keeps
the list's shape: the reference's list starts with ``manning_friction_implicit``), and a ``Wind_stress``
is evaluated on the host into two per-triangle arrays that the update kernels add to the explicit
updates (swk_set_explicit_forcing); General_forcing, Rainfall and Inflow add their rate to the stage update the same way.
"""
import math

import numpy as np

# anuga/config.py:50-52
eta_w = 3.0e-3   # wind stress coefficient
rho_a = 1.2e-3   # atmospheric density
rho_w = 1023     # fluid density


def manning_friction_implicit(domain):
    """placeholder with the reference's name: friction runs inside the device update kernels"""


def _check_forcefield(f):
    """forcing.py:23-76: a callable f(t, x, y) returning one value per point, or a scalar"""
    if callable(f):
        x = np.ones(3, dtype=np.float64)
        y = np.ones(3, dtype=np.float64)
        try:
            q = f(1.0, x=x, y=y)
        except Exception as e:
            raise Exception("Function %s could not be executed:\n%s" % (f, e))
        try:
            q = np.array(q, dtype=np.float64)
        except Exception:
            raise Exception("Return value from vector function %s could not be converted into a numeric array "
                            "of floats.\nSpecified function should return either list or array." % f)
        assert len(q) == 3, "%s must return vector of length 3" % f
        return f
    try:
        return float(f)
    except Exception:
        raise Exception("Force field %s must be a scalar value coercible to float." % str(f))


class Wind_stress:
    """Wind stress on the water momentum from wind speed s [m/s] and direction phi [degrees]
    (forcing.py:80-186).  Wind_stress(s, phi), Wind_stress(s=..., phi=...) with scalars or functions
    f(t, x, y), or Wind_stress(F) with one function returning (s, phi)."""

    def __init__(self, *args, **kwargs):
        self.use_coordinates = True
        if len(args) == 2:
            s, phi = args
        elif len(args) == 1:
            vector_function = args[0]
            if len(kwargs) == 1:
                self.use_coordinates = kwargs["use_coordinates"]
            if self.use_coordinates:
                s = lambda t, x, y: vector_function(t, x=x, y=y)[0]
                phi = lambda t, x, y: vector_function(t, x=x, y=y)[1]
            else:
                s = lambda t, i: vector_function(t, point_id=i)[0]
                phi = lambda t, i: vector_function(t, point_id=i)[1]
        else:
            if len(kwargs) == 2:
                s, phi = kwargs["s"], kwargs["phi"]
            else:
                raise Exception("Assumes two keyword arguments: s=..., phi=....")
        if self.use_coordinates:
            self.speed = _check_forcefield(s)
            self.phi = _check_forcefield(phi)
        else:
            self.speed, self.phi = s, phi
        self.const = eta_w * rho_a / rho_w

    @property
    def time_dependent(self):
        return callable(self.speed) or callable(self.phi)

    def _field(self, f, t, xc):
        N = len(xc)
        if callable(f):
            if self.use_coordinates:
                return np.asarray(f(t, xc[:, 0], xc[:, 1]), dtype=np.float64) * np.ones(N)
            return np.array([f(t, i) for i in range(N)], dtype=np.float64)
        return f * np.ones(N, dtype=np.float64)

    def explicit_forcing(self, domain, t, F):
        """add this term to the explicit-update additions F = {quantity: (N,) array}"""
        fx, fy = self.momentum_forcing(domain, t)
        F["xmomentum"] += fx
        F["ymomentum"] += fy

    def momentum_forcing(self, domain, t):
        """(S*u, S*v) per triangle with the arithmetic of assign_windfield_values (forcing.py:189-215):
        Python's math functions on every distinct (s, phi) pair."""
        xc = domain.get_centroid_coordinates()
        s_vec = self._field(self.speed, t, xc)
        phi_vec = self._field(self.phi, t, xc)
        pairs, inverse = np.unique(np.stack([s_vec, phi_vec], axis=1), axis=0, return_inverse=True)
        fx = np.empty(len(pairs))
        fy = np.empty(len(pairs))
        for j, (s, phi) in enumerate(pairs):
            s, phi = float(s), float(phi)
            phi = phi * math.pi / 180.0
            u = s * math.cos(phi)
            v = s * math.sin(phi)
            S = self.const * math.sqrt(u ** 2 + v ** 2)
            fx[j] = S * u
            fy[j] = S * v
        inverse = np.asarray(inverse).reshape(-1)
        return fx[inverse], fy[inverse]

    def __call__(self, domain):
        """host form (the reference's call): add to the explicit updates of the host arrays"""
        fx, fy = self.momentum_forcing(domain, domain.get_time())
        domain.quantities["xmomentum"].explicit_update[:] += fx
        domain.quantities["ymomentum"].explicit_update[:] += fy

    def oracle_spec(self):
        return ("wind", self)


class General_forcing:
    """General explicit forcing term: a rate [quantity/s] added to the explicit update of one conserved
    quantity over a circle, a polygon or the whole domain (forcing.py:215-495).  rate: a number or a function
    of time; default_rate takes over when the rate function runs out of data (Modeltime_too_late)."""

    def __init__(self, domain, quantity_name, rate=0.0, center=None, radius=None, polygon=None,
                 default_rate=None, verbose=False):
        if center is None:
            assert radius is None, "I got radius but no center."
        if radius is None:
            assert center is None, "I got center but no radius."
        if quantity_name not in domain.conserved_quantities:
            raise Exception("%s is not a conserved quantity" % quantity_name)
        self.domain = domain
        self.quantity_name = quantity_name
        self.rate = rate
        self.center = None if center is None else np.asarray(center, dtype=np.float64)
        self.radius = radius
        self.polygon = polygon
        self.verbose = verbose
        self.value = 0.0
        points = domain.get_centroid_coordinates(absolute=True)
        self.exchange_indices = None
        if self.center is not None and self.radius is not None:
            assert len(self.center) == 2
            assert polygon is None, "Polygon cannot be specified when center and radius are"
            c = self.center
            inside = ((points[:, 0] - c[0]) ** 2 + (points[:, 1] - c[1]) ** 2) < self.radius ** 2
            self.exchange_indices = np.flatnonzero(inside)
        if self.polygon is not None:
            from .structures import Region
            self.exchange_indices = np.asarray(Region(domain, polygon=self.polygon).indices, dtype=np.int64)
        if self.exchange_indices is None:
            self.exchange_area = None          # (polygon_area of the mesh boundary in the reference; only Inflow divides by it)
        else:
            if len(self.exchange_indices) == 0:
                raise Exception("No triangles have been identified in specified region: center=%s, radius=%s"
                                % (self.center, self.radius))
            area = 0.0
            for i in self.exchange_indices:        # the reference's running sum, in index order (:373-375)
                area += domain.areas[i]
            self.exchange_area = area
            assert self.exchange_area > 0.0
        assert default_rate is None or isinstance(default_rate, (int, float)) or callable(default_rate), \
            "Keyword argument default_rate must be either None or a function of time.\nI got %s." % str(default_rate)
        if default_rate is not None and not callable(default_rate):
            tmp = default_rate
            default_rate = lambda t: tmp
        self.default_rate = default_rate
        self.default_rate_invoked = False

    @property
    def time_dependent(self):
        return callable(self.rate)

    def update_rate(self, t):
        return self.rate(t) if callable(self.rate) else self.rate

    def current_rate(self, t):
        try:
            rate = self.update_rate(t)
        except BaseException as e:
            if type(e).__name__ != "Modeltime_too_late":
                raise
            if self.default_rate is None:
                raise type(e)("%s: ANUGA is trying to run longer than specified data.\nYou can specify keyword "
                              "argument default_rate in the forcing function to tell it what to do in the absence "
                              "of time data." % str(e))
            rate = self.default_rate(t)
            if not self.default_rate_invoked:
                import warnings
                warnings.warn("%s\nInstead I will use the default rate: %s\nNote: Further warnings will be supressed"
                              % (str(e), str(self.default_rate)))
                self.default_rate_invoked = True
        if rate is None:
            raise Exception("Attribute rate must be specified in General_forcing or its descendants before "
                            "attempting to call it")
        return rate

    def explicit_forcing(self, domain, t, F):
        rate = self.current_rate(t)
        if self.exchange_indices is None:
            F[self.quantity_name][:] += rate
        else:
            F[self.quantity_name][self.exchange_indices] += rate

    def __call__(self, domain):
        """host form (the reference's call): add to the explicit update of the host array"""
        upd = domain.quantities[self.quantity_name].explicit_update
        rate = self.current_rate(domain.get_time())
        if self.exchange_indices is None:
            upd[:] += rate
        else:
            upd[self.exchange_indices] += rate

    def get_quantity_values(self, quantity_name=None):
        q = self.domain.quantities[quantity_name or self.quantity_name]
        v = q.centroid_values
        return v.copy() if self.exchange_indices is None else v[self.exchange_indices]

    def parallel_safe(self):
        return True

    def oracle_spec(self):
        return ("general_forcing", self)


class Rainfall(General_forcing):
    """rain [mm/s] over a region or the whole domain, added to the stage update (forcing.py:497-575)"""

    def __init__(self, domain, rate=0.0, center=None, radius=None, polygon=None, default_rate=None, verbose=False):
        if callable(rate):
            rain = lambda t: rate(t) / 1000.0
        else:
            rain = rate / 1000.0
        if default_rate is not None:
            if callable(default_rate):
                default_rain = lambda t: default_rate(t) / 1000.0
            else:
                default_rain = default_rate / 1000.0
        else:
            default_rain = None
        General_forcing.__init__(self, domain, "stage", rate=rain, center=center, radius=radius, polygon=polygon,
                                 default_rate=default_rain, verbose=verbose)


class Inflow(General_forcing):
    """flow [m^3/s] into (or out of) a region: divided by the region's area and added to the stage update
    (forcing.py:578-640)"""

    def __init__(self, domain, rate=0.0, center=None, radius=None, polygon=None, default_rate=None, verbose=False):
        General_forcing.__init__(self, domain, "stage", rate=rate, center=center, radius=radius, polygon=polygon,
                                 default_rate=default_rate, verbose=verbose)
        if self.exchange_area is None:
            raise NotImplementedError("Inflow over the whole domain (the reference divides by the area of the "
                                      "mesh boundary polygon): give center/radius or a polygon")

    def update_rate(self, t):
        if callable(self.rate):
            return self.rate(t) / self.exchange_area
        return self.rate / self.exchange_area
