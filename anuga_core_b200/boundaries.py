"""Boundary condition objects with the reference's constructor signatures.

Each object is a *descriptor*: it names one of the device boundary kinds of
include/swk.h and, for time-dependent conditions, provides the host-evaluated
scalars that are uploaded before every flux evaluation.  The arithmetic runs in
k_boundary_values (csrc/swk_kernels.cuh) and follows evaluate_segment of

  Reflective_boundary                      anuga/shallow_water/boundaries.py:235-288
  Transmissive_momentum_set_stage_boundary                         :344-372
  Transmissive_n_momentum_zero_t_momentum_set_stage_boundary       :477-517
  Transmissive_stage_zero_momentum_boundary                        :543-551
  Time_stage_zero_momentum_boundary                                :616-635
  Flather_external_stage_zero_velocity_boundary                    :1096-1266
  Characteristic_stage_boundary                                    :639-843
  Dirichlet_discharge_boundary                                     :845-890
  Transmissive_boundary    anuga/abstract_2d_finite_volumes/generic_boundary_conditions.py:173-193
  Dirichlet_boundary                                               :221-264
  Time_boundary                                                    :370-411
"""
import numpy as np

from . import backend as _b


class Modeltime_too_late(BaseException):
    """a time series (file function) was asked for a time after its last record
    (anuga/fit_interpolate/interpolate.py:55)"""


class Modeltime_too_early(BaseException):
    pass


def _is_named(e, name):
    # the reference's own exception classes arrive here when a reference function is attached
    return type(e).__name__ == name


def evaluate_with_default(owner, function, t, default):
    """function(t); when it runs out of data (Modeltime_too_late) the default value takes over, as in
    Boundary.get_boundary_values (generic_boundary_conditions.py:98-135).  `default` may be a value
    or a function of t; the first hand-over is flagged on `owner`."""
    try:
        return function(t)
    except BaseException as e:
        if not _is_named(e, "Modeltime_too_late") or default is None:
            raise
        if not getattr(owner, "default_boundary_invoked", False):
            owner.default_boundary_invoked = True
            if getattr(owner, "verbose", False):
                print("%s\nInstead I will use the default boundary value: %s\n"
                      "Note: Further warnings will be suppressed" % (e, default))
        import copy
        return default(t) if callable(default) else copy.deepcopy(default)


class Boundary:
    device_kind = _b.BC_NONE
    time_dependent = False

    def device_values(self, t):
        return (0.0, 0.0, 0.0)

    def values_for_substep(self, dev, seg, substep, t, ids=None):
        """the segment's three value-table entries for RK substep `substep` at time t; `ids` are the boundary
        indices of the segment (boundaries that keep per-substep data on the device, e.g. Time_space_boundary,
        refresh it here)"""
        return self.device_values(t)

    def oracle_spec(self):
        """Specification tuple understood by the CPU oracle (tests only)."""
        raise NotImplementedError


class Reflective_boundary(Boundary):
    device_kind = _b.BC_REFLECTIVE

    def __init__(self, domain=None):
        if domain is None:
            raise Exception("Domain must be specified for reflective boundary")
        self.domain = domain

    def __repr__(self):
        return "Reflective_boundary"

    def oracle_spec(self):
        return ("reflective",)


class Dirichlet_boundary(Boundary):
    device_kind = _b.BC_DIRICHLET

    def __init__(self, dirichlet_values=None):
        if dirichlet_values is None:
            raise Exception("Must specify one value for each quantity")
        self.dirichlet_values = np.array(dirichlet_values, dtype=np.float64)
        if self.dirichlet_values.size < 3:
            raise Exception("Dirichlet boundary needs stage, xmomentum, ymomentum")

    def __repr__(self):
        return "Dirichlet boundary (%s)" % self.dirichlet_values

    def device_values(self, t):
        return tuple(float(v) for v in self.dirichlet_values[:3])

    def oracle_spec(self):
        return ("dirichlet", [float(v) for v in self.dirichlet_values[:3]])


class Transmissive_boundary(Boundary):
    device_kind = _b.BC_TRANSMISSIVE

    def __init__(self, domain=None):
        if domain is None:
            raise Exception("Domain must be specified for transmissive boundary")
        self.domain = domain

    def __repr__(self):
        return "Transmissive_boundary(%s)" % self.domain

    def oracle_spec(self):
        return ("transmissive",)


class Time_boundary(Boundary):
    """Dirichlet values given as a function of time, evaluated on the host before
    each flux evaluation."""
    device_kind = _b.BC_DIRICHLET
    time_dependent = True

    def __init__(self, domain=None, function=None, default_boundary=None, verbose=False):
        if domain is None:
            raise Exception("You must specify a domain to Time_boundary")
        if function is None:
            raise Exception("You must specify a function to Time_boundary")
        q = np.asarray(function(0.0), dtype=np.float64)
        if q.size < 3:
            raise Exception("Time_boundary function must return (stage, xmomentum, ymomentum)")
        self.domain = domain
        self.function = function
        self.default_boundary = default_boundary
        self.default_boundary_invoked = False
        self.verbose = verbose

    def __repr__(self):
        return "Time boundary"

    def device_values(self, t):
        q = np.asarray(evaluate_with_default(self, self.function, t, self.default_boundary), dtype=np.float64)
        return (float(q[0]), float(q[1]), float(q[2]))

    def oracle_spec(self):
        return ("time", self.function)


class _Set_stage(Boundary):
    time_dependent = True
    _oracle_kind = None

    def __init__(self, domain=None, function=None, default_boundary=0.0):
        if domain is None:
            raise Exception("Domain must be specified for this type boundary")
        if function is None:
            raise Exception("Function must be specified for this type boundary")
        if isinstance(function, (int, float)):
            tmp = function
            function = lambda t: tmp
        self.domain = domain
        self.function = function
        self.default_boundary = default_boundary
        self.default_boundary_invoked = False

    def device_values(self, t):
        value = evaluate_with_default(self, self.function, t, self.default_boundary)
        try:
            x = float(value)
        except Exception:
            x = float(value[0])
        return (x, 0.0, 0.0)

    def oracle_spec(self):
        return (self._oracle_kind, self.function)


class Transmissive_n_momentum_zero_t_momentum_set_stage_boundary(_Set_stage):
    device_kind = _b.BC_TRANSMISSIVE_N_ZERO_T_SET_STAGE
    _oracle_kind = "transmissive_n_zero_t_set_stage"


class Transmissive_momentum_set_stage_boundary(_Set_stage):
    device_kind = _b.BC_TRANSMISSIVE_MOMENTUM_SET_STAGE
    _oracle_kind = "transmissive_momentum_set_stage"


class Time_stage_zero_momentum_boundary(_Set_stage):
    device_kind = _b.BC_DIRICHLET
    _oracle_kind = "time_stage_zero_momentum"


class Flather_external_stage_zero_velocity_boundary(_Set_stage):
    """weakly reflecting open boundary: external stage f(t), zero external velocity, combined with the
    interior state through characteristic variables (vectorised form, boundaries.py:1207-1266)"""
    device_kind = _b.BC_FLATHER_EXTERNAL_STAGE_ZERO_VELOCITY
    _oracle_kind = "flather_external_stage_zero_velocity"

    def __init__(self, domain=None, function=None):
        _Set_stage.__init__(self, domain, function)


class Characteristic_stage_boundary(_Set_stage):
    """stage from a function of time, momentum from the Riemann invariants of a simple incoming wave
    (boundaries.py:639-843, the vectorised evaluate_segment)"""
    device_kind = _b.BC_CHARACTERISTIC_STAGE
    _oracle_kind = "characteristic_stage"

    def __init__(self, domain=None, function=None, default_stage=0.0):
        _Set_stage.__init__(self, domain, function, default_boundary=None)
        self.default_stage = default_stage

    def __repr__(self):
        return "Characteristic_stage_boundary (%s) (%s) " % (self.domain, self.default_stage)


class Dirichlet_discharge_boundary(Boundary):
    """sets the stage (stage0) and the momentum wh0 along the inward normal (boundaries.py:845-890)"""
    device_kind = _b.BC_DIRICHLET_DISCHARGE

    def __init__(self, domain=None, stage0=None, wh0=None):
        if domain is None:
            raise Exception("Domain must be specified for this type of boundary")
        if stage0 is None:
            raise Exception("Stage must be specified for this type of boundary")
        self.domain = domain
        self.stage0 = stage0
        self.wh0 = 0.0 if wh0 is None else wh0

    def __repr__(self):
        return "Dirichlet_Discharge_boundary(%s)" % self.domain

    def device_values(self, t):
        return (float(self.stage0), float(self.wh0), 0.0)

    def oracle_spec(self):
        return ("dirichlet_discharge", float(self.stage0), float(self.wh0))


class Transmissive_stage_zero_momentum_boundary(Boundary):
    device_kind = _b.BC_TRANSMISSIVE_STAGE_ZERO_MOMENTUM

    def __init__(self, domain=None):
        if domain is None:
            raise Exception("Domain must be specified for Transmissive_stage_zero_momentum boundary")
        self.domain = domain

    def oracle_spec(self):
        return ("transmissive_stage_zero_momentum",)
