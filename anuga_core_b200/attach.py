"""Drop-in for a *reference* ``anuga.shallow_water.Domain``: the new device multiprocessor_mode.

    import anuga, anuga_core_b200
    domain = anuga.rectangular_cross_domain(...)          # the reference's own Domain
    ... set_flow_algorithm / set_quantity / set_boundary as usual ...
    anuga_core_b200.set_multiprocessor_mode_b200(domain)  # instead of domain.set_multiprocessor_mode(4)
    for t in domain.evolve(yieldstep=..., finaltime=...):  # the reference's evolve wrapper (SWW, checkpoints)
        ...

What it does:
  * builds a device handle from the reference domain's OWN mesh arrays (no recomputation) and keeps the
    reference's numpy arrays as the user-visible truth: ``domain.quantities[...].centroid_values`` are
    updated IN PLACE at every yield, so aliases held by user code / operators stay valid;
  * reads the scalars (betas, CFL, minimum_allowed_height, timestepping method, ...) at every evolve()
    start, not at mode selection (the reference's mode 4 snapshots them once and goes stale,
    sw_domain_cuda.py:36-171);
  * replaces ``domain._evolve_base`` (generic_domain.py:1715) by the device time loop - the
    reference's ``Domain.evolve`` wrapper (shallow_water_domain.py:2300-2407) keeps calling it, so
    ``store_timestep`` and checkpoint pickling keep working on host arrays;
  * maps the reference's boundary objects and Rate_operators to device kinds by class name; anything
    it cannot map raises (no silent CPU fallback).
"""
import numpy as np

from . import boundaries as _bnd
from . import file_boundary as _fb
from . import structures as _st
from .domain import Domain, MODE_B200
from .operators import Rate_operator

_BOUNDARY_BY_NAME = {
    "Reflective_boundary": lambda B, d: _bnd.Reflective_boundary(d),
    "Dirichlet_boundary": lambda B, d: _bnd.Dirichlet_boundary(B.dirichlet_values),
    "Transmissive_boundary": lambda B, d: _bnd.Transmissive_boundary(d),
    "Time_boundary": lambda B, d: _bnd.Time_boundary(d, B.function),
    "Transmissive_n_momentum_zero_t_momentum_set_stage_boundary":
        lambda B, d: _bnd.Transmissive_n_momentum_zero_t_momentum_set_stage_boundary(d, B.function),
    "Transmissive_momentum_set_stage_boundary":
        lambda B, d: _bnd.Transmissive_momentum_set_stage_boundary(d, B.function),
    "Transmissive_stage_zero_momentum_boundary":
        lambda B, d: _bnd.Transmissive_stage_zero_momentum_boundary(d),
    "Time_stage_zero_momentum_boundary": lambda B, d: _bnd.Time_stage_zero_momentum_boundary(d, B.f),
    "Flather_external_stage_zero_velocity_boundary":
        lambda B, d: _bnd.Flather_external_stage_zero_velocity_boundary(d, B.function),
    "File_boundary": lambda B, d: _fb.File_boundary.adopt(B, d),
    "Field_boundary": lambda B, d: _fb.Field_boundary.adopt(B, d),
    "Time_space_boundary": lambda B, d: _fb.Time_space_boundary(d, B.function, B.default_boundary),
    "Dirichlet_discharge_boundary": lambda B, d: _bnd.Dirichlet_discharge_boundary(d, B.stage0, B.wh0),
    "Characteristic_stage_boundary":
        lambda B, d: _bnd.Characteristic_stage_boundary(d, B.function, B.default_stage),
}

_SCALARS = ("epsilon", "H0", "g", "minimum_allowed_height", "maximum_allowed_speed", "evolve_max_timestep",
            "evolve_min_timestep", "max_smallsteps", "CFL", "beta_w", "beta_w_dry", "beta_uh", "beta_uh_dry",
            "beta_vh", "beta_vh_dry", "low_froude", "extrapolate_velocity_second_order", "use_sloped_mannings",
            "default_order", "ghost_layer_width", "centroid_transmissive_bc", "fixed_flux_timestep",
            "starttime")

_QUANTITIES = ("stage", "xmomentum", "ymomentum", "elevation", "friction", "height", "xvelocity", "yvelocity")


class B200_interface:
    """Counterpart of the reference's GPU_interface (shallow_water/sw_domain_cuda.py) for mode 5."""

    def __init__(self, ref_domain, device=0, reorder=True):
        self.ref = ref_domain
        if not ref_domain.get_using_discontinuous_elevation():
            raise Exception("the B200 mode implements the DE flow algorithms only (DE0, DE1, DE2, DE0_7, DE1_7)")
        d = self.dev_domain = Domain(mesh=ref_domain.mesh, device=device, reorder=reorder,
                                     full_send_dict=getattr(ref_domain, "full_send_dict", None),
                                     ghost_recv_dict=getattr(ref_domain, "ghost_recv_dict", None),
                                     processor=getattr(ref_domain, "processor", 0),
                                     numproc=getattr(ref_domain, "numproc", 1),
                                     ghost_layer_width=getattr(ref_domain, "ghost_layer_width", 2))
        d.tri_full_flag = np.ascontiguousarray(ref_domain.tri_full_flag, dtype=np.int64)
        # share the reference's arrays: same buffers, updated in place
        for name in _QUANTITIES:
            if name in ref_domain.quantities:
                rq, q = ref_domain.quantities[name], d.quantities[name]
                q.centroid_values = rq.centroid_values
                q.boundary_values = rq.boundary_values
                q._arrays["vertex_values"] = rq.vertex_values
                q._arrays["edge_values"] = rq.edge_values
                q._arrays["explicit_update"] = rq.explicit_update
                q._arrays["semi_implicit_update"] = rq.semi_implicit_update
        for k in ("edge_flux_type", "edge_river_wall_counter", "riverwall_elevation", "riverwall_rowIndex",
                  "riverwall_hydraulic_properties", "number_of_riverwall_edges",
                  "ncol_riverwall_hydraulic_properties"):
            if hasattr(ref_domain, k) and getattr(ref_domain, "number_of_riverwall_edges", 0) > 0:
                setattr(d, k, getattr(ref_domain, k))
        self.refresh()

    # -- state that may change between evolve() calls -----------------------------------
    def refresh(self):
        ref, d = self.ref, self.dev_domain
        for k in _SCALARS:
            if hasattr(ref, k):
                setattr(d, k, getattr(ref, k))
        d.timestepping_method = ref.get_timestepping_method()
        d.flow_algorithm = ref.get_flow_algorithm()
        d._params_dirty = True
        bmap = {}
        for tag, B in (ref.boundary_map or {}).items():
            if B is None:
                bmap[tag] = None
                continue
            name = type(B).__name__
            if name not in _BOUNDARY_BY_NAME:
                raise NotImplementedError("boundary %s has no device kind (SURVEY.md 8(f) row 4)" % name)
            bmap[tag] = _BOUNDARY_BY_NAME[name](B, d)
        if bmap:
            d.boundary_map = None
            d.set_boundary(bmap)
        # the operator list is rebuilt from the reference's at every evolve(): forget the device copies of
        # the previous call first, or every rain / rate operator would be applied once more per call
        d.fractional_step_operators = []
        if d._dev is not None:
            d._dev.clear_rate_operators()
        d._operators_dirty = True
        for op in getattr(ref, "fractional_step_operators", []):
            name = type(op).__name__
            if name == "boundary_flux_integral_operator":
                continue                      # built into the device step
            if name == "Inlet_operator":          # host-side hydraulics on gathered cells (structures.py)
                new = _st.Inlet_operator(d, _st.Region(d, indices=op.inlet.triangle_indices), Q=op.Q,
                                         velocity=op.velocity, zero_velocity=op.zero_velocity,
                                         default=getattr(op, "default", 0.0))
                new._mirror = op
                continue
            if name in ("Boyd_box_operator", "Boyd_pipe_operator", "Weir_orifice_trapezoid_operator"):
                getattr(_st, name).adopt(d, op)
                continue
            if name != "Rate_operator":
                raise NotImplementedError("operator %s has no device implementation (SURVEY.md 8(f))" % name)
            if getattr(op, "rate_type", None) not in ("scalar", "t", "centroid_array", "quantity") \
                    or getattr(op, "rate_spatial", False):
                raise NotImplementedError("Rate_operator with a spatial function rate")
            rate = getattr(op, "rate_input", op.rate)
            if op.rate_type == "quantity":
                rate = rate.centroid_values
            Rate_operator(d, rate=rate, factor=op.factor, indices=op.indices)
        for q in d.quantities.values():
            q.host_dirty = True
        d.relative_time = ref.relative_time

    # -- the reference's GPU_interface method names (per-call use by the dispatch sites) -----
    def allocate_gpu_arrays(self):
        self.dev_domain._ensure_device()

    def compile_gpu_kernels(self):
        from . import backend
        backend.load_library()                # kernels are compiled ahead of time for sm_100a

    def compute_fluxes_ext_central_kernel(self, timestep=None):
        d = self.dev_domain
        d._push_quantities()
        return d.compute_fluxes(0)

    def extrapolate_second_order_edge_sw_kernel(self, domain=None):
        self.dev_domain.distribute_to_vertices_and_edges()
        self.dev_domain.sync_to_host()

    def protect_against_infinitesimal_and_negative_heights_kernal(self, domain=None):
        me = self.dev_domain.protect_against_infinitesimal_and_negative_heights()
        self.dev_domain.sync_to_host()
        return me

    def update_conserved_quantities_kernal(self, domain=None):
        d = self.dev_domain
        d.timestep = self.ref.timestep
        d.update_conserved_quantities()
        d.sync_to_host()
        return 0

    # friction.py:133-147 calls these in mode 4 (manning_friction_implicit_gpu): the semi-implicit updates
    # of the reference domain's own momentum quantities through the per-call C ABI entry points
    def _friction(self, sloped):
        from . import backend as B
        lib = B.load_library()
        ref = self.ref
        q = ref.quantities
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        w, uh, vh = f64(q["stage"].centroid_values), f64(q["xmomentum"].centroid_values), f64(q["ymomentum"].centroid_values)
        eta = f64(q["friction"].centroid_values)
        xu, yu = q["xmomentum"].semi_implicit_update, q["ymomentum"].semi_implicit_update
        assert xu.flags.c_contiguous and yu.flags.c_contiguous
        N = len(w)
        dev = self.dev_domain.device
        if sloped:
            x = f64(ref.get_vertex_coordinates())
            zv = f64(q["elevation"].vertex_values)
            B._check(lib.swk_call_manning_friction_sloped(dev, float(ref.g), float(ref.minimum_allowed_height), N,
                                                          B._pd(x), B._pd(w), B._pd(zv), B._pd(uh), B._pd(vh),
                                                          B._pd(eta), B._pd(xu), B._pd(yu)))
        else:
            z = f64(q["elevation"].centroid_values)
            B._check(lib.swk_call_manning_friction_flat(dev, float(ref.g), float(ref.minimum_allowed_height), N,
                                                        B._pd(w), B._pd(z), B._pd(uh), B._pd(vh), B._pd(eta),
                                                        B._pd(xu), B._pd(yu)))

    def compute_forcing_terms_manning_friction_flat(self):
        self._friction(False)

    def compute_forcing_terms_manning_friction_sloped(self):
        self._friction(True)

    # -- resident time loop behind the reference's evolve wrapper --------------------------------
    def evolve_base(self, yieldstep=None, finaltime=None, duration=None, skip_initial_step=False):
        ref, d = self.ref, self.dev_domain
        self.refresh()
        d.evolved_called = ref.evolved_called
        d.evolve_starttime = getattr(ref, "evolve_starttime", d.relative_time)
        ref.evolved_called = True
        for t in d.evolve(yieldstep=yieldstep, finaltime=finaltime, duration=duration,
                          skip_initial_step=skip_initial_step):
            for k in ("relative_time", "timestep", "flux_timestep", "number_of_steps",
                      "number_of_first_order_steps", "recorded_min_timestep", "recorded_max_timestep",
                      "yieldstep", "finaltime", "relative_finaltime", "relative_yieldtime"):
                if hasattr(d, k):
                    setattr(ref, k, getattr(d, k))
            ref.boundary_flux_integral_value = d.boundary_flux_integral
            yield ref.get_time()


def set_multiprocessor_mode_b200(ref_domain, device=0, reorder=True):
    """The new device multiprocessor_mode for a reference Domain.  Raises when no sm_100 device /
    library is usable (north_star: no CPU fallback)."""
    from . import backend
    backend.load_library()
    if backend.device_count() < 1:
        raise backend.SwkError(-1, "no sm_100 device: multiprocessor_mode %d cannot be selected" % MODE_B200)
    iface = B200_interface(ref_domain, device=device, reorder=reorder)
    ref_domain.multiprocessor_mode = MODE_B200
    ref_domain.gpu_interface = iface
    ref_domain._evolve_base = iface.evolve_base
    ref_domain.distribute_to_vertices_and_edges = iface.extrapolate_second_order_edge_sw_kernel
    ref_domain.protect_against_infinitesimal_and_negative_heights = \
        iface.protect_against_infinitesimal_and_negative_heights_kernal
    return iface
