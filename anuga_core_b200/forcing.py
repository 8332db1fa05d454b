"""Explicit forcing terms (anuga/shallow_water/forcing.py).

``domain.forcing_terms`` is the reference's list of callables ``f(domain)`` that
``compute_forcing_terms`` runs after every flux evaluation (generic_domain.py:2417-2427).  Here the
list holds descriptors: Manning friction is built into the update kernels (the placeholder below keeps
the list's shape: the reference's list starts with ``manning_friction_implicit``), and a ``Wind_stress``
is evaluated on the host into two per-triangle arrays that the update kernels add to the explicit
momentum updates (swk_set_momentum_forcing).
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
