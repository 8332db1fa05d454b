"""ctypes binding of libswk.so (include/swk.h).

This is the thin host-side layer the north_star asks for: Python host code
calling hand-written sm_100a CUDA through a C ABI, no PyTorch requirement and
NO CPU FALLBACK - importing works without a GPU (so that CPU-only tools can
introspect the ABI), but every compute entry point raises ``SwkError`` when the
library or an sm_100 device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SWK_LIB", os.path.join(_HERE, "libswk.so"))   # SWK_LIB: kernel-variant experiments

SWK_OK = 0
ERRORS = {
    -1: "SWK_ERR_CUDA", -2: "SWK_ERR_ARG", -3: "SWK_ERR_DENOMINATOR", -4: "SWK_ERR_SMALLSTEP",
    -5: "SWK_ERR_OVERSHOOT", -6: "SWK_ERR_UNSUPPORTED", -7: "SWK_ERR_NCCL",
}

# quantity ids (include/swk.h)
Q = dict(
    STAGE_C=0, XMOM_C=1, YMOM_C=2, ELEVATION_C=3, FRICTION_C=4, HEIGHT_C=5,
    STAGE_E=10, XMOM_E=11, YMOM_E=12, HEIGHT_E=13, ELEVATION_E=14,
    STAGE_V=20, XMOM_V=21, YMOM_V=22, HEIGHT_V=23, ELEVATION_V=24,
    STAGE_B=30, XMOM_B=31, YMOM_B=32,
    STAGE_EU=40, XMOM_EU=41, YMOM_EU=42, MAX_SPEED=50,
    STAGE_BACKUP=60, XMOM_BACKUP=61, YMOM_BACKUP=62,
)

BC_NONE, BC_REFLECTIVE, BC_DIRICHLET, BC_TRANSMISSIVE = 0, 1, 2, 3
BC_TRANSMISSIVE_N_ZERO_T_SET_STAGE, BC_TRANSMISSIVE_MOMENTUM_SET_STAGE = 4, 5
BC_TRANSMISSIVE_STAGE_ZERO_MOMENTUM = 6
BC_FLATHER_EXTERNAL_STAGE_ZERO_VELOCITY = 7
BC_CHARACTERISTIC_STAGE = 8
BC_TIME_SPACE_TABLE, BC_TIME_SPACE_TABLE_MEAN_STAGE = 9, 10
BC_DIRICHLET_DISCHARGE = 11

_I = C.c_int64
_D = C.c_double
_PI = C.POINTER(C.c_int64)
_PD = C.POINTER(C.c_double)


class SwkError(RuntimeError):
    def __init__(self, code, message):
        self.code = code
        RuntimeError.__init__(self, "%s: %s" % (ERRORS.get(code, code), message))


class SwkParams(C.Structure):
    _fields_ = [
        ("epsilon", _D), ("H0", _D), ("g", _D), ("minimum_allowed_height", _D),
        ("maximum_allowed_speed", _D), ("evolve_max_timestep", _D), ("evolve_min_timestep", _D),
        ("beta_w", _D), ("beta_w_dry", _D), ("beta_uh", _D), ("beta_uh_dry", _D),
        ("beta_vh", _D), ("beta_vh_dry", _D), ("CFL", _D), ("fixed_flux_timestep", _D),
        ("extrapolate_velocity_second_order", _I), ("low_froude", _I), ("timestepping_method", _I),
        ("use_sloped_mannings", _I), ("max_smallsteps", _I), ("default_order", _I),
        ("ghost_layer_width", _I), ("centroid_transmissive_bc", _I), ("track_max_speed", _I),
    ]


class SwkMesh(C.Structure):
    _fields_ = [
        ("number_of_elements", _I), ("boundary_length", _I),
        ("neighbours", _PI), ("neighbour_edges", _PI), ("surrogate_neighbours", _PI),
        ("number_of_boundaries", _PI), ("tri_full_flag", _PI),
        ("normals", _PD), ("edgelengths", _PD), ("radii", _PD), ("areas", _PD),
        ("centroid_coordinates", _PD), ("edge_coordinates", _PD), ("vertex_coordinates", _PD),
        ("boundary_cells", _PI), ("boundary_edges", _PI),
        ("number_of_riverwall_edges", _I), ("ncol_riverwall_hydraulic_properties", _I),
        ("edge_flux_type", _PI), ("edge_river_wall_counter", _PI),
        ("riverwall_elevation", _PD), ("riverwall_rowIndex", _PI),
        ("riverwall_hydraulic_properties", _PD),
        ("permutation", _PI),
    ]


class SwkEvolveResult(C.Structure):
    _fields_ = [
        ("time", _D), ("timestep", _D), ("flux_timestep", _D),
        ("recorded_min_timestep", _D), ("recorded_max_timestep", _D),
        ("boundary_flux_integral", _D), ("fractional_step_volume_integral", _D),
        ("mass_error", _D), ("boundary_flux_sum", _D * 3),
        ("number_of_steps", _I), ("number_of_first_order_steps", _I), ("total_steps", _I),
        ("negative_cells", _I), ("stop_reason", _I), ("kernel_launches", _I),
    ]


_HOST_VIEW_ARRAYS = [
    "stage_centroid_values", "xmom_centroid_values", "ymom_centroid_values",
    "bed_centroid_values", "height_centroid_values", "friction_centroid_values",
    "stage_edge_values", "xmom_edge_values", "ymom_edge_values", "bed_edge_values", "height_edge_values",
    "stage_vertex_values", "xmom_vertex_values", "ymom_vertex_values", "bed_vertex_values", "height_vertex_values",
    "stage_boundary_values", "xmom_boundary_values", "ymom_boundary_values",
    "stage_explicit_update", "xmom_explicit_update", "ymom_explicit_update",
    "stage_semi_implicit_update", "xmom_semi_implicit_update", "ymom_semi_implicit_update",
    "max_speed", "boundary_flux_sum",
]


class SwkHostView(C.Structure):
    _fields_ = [("mesh", SwkMesh), ("params", SwkParams)] + [(n, _PD) for n in _HOST_VIEW_ARRAYS]


# every symbol include/swk.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "swk_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "swk_last_error": (C.c_char_p, []),
    "swk_abi_version": (C.c_int, []),
    "swk_build_neighbour_structure": (C.c_int, [_I, _I, _PI, _PI, _PI, _PI]),
    "swk_mesh_geometry": (C.c_int, [_I, _I, _PD, _PI, _I, _PD, _PD, _PD, _PD, _PD, _PD, _PD, _PI]),
    "swk_create": (C.c_int, [C.POINTER(SwkMesh), C.POINTER(SwkParams), C.c_int, C.POINTER(_H)]),
    "swk_destroy": (C.c_int, [_H]),
    "swk_set_params": (C.c_int, [_H, C.POINTER(SwkParams)]),
    "swk_set_quantity": (C.c_int, [_H, C.c_int, _PD, _I]),
    "swk_get_quantity": (C.c_int, [_H, C.c_int, _PD, _I]),
    "swk_set_boundary_segment": (C.c_int, [_H, C.c_int, C.c_int, _PI, _I, _PD]),
    "swk_set_boundary_values": (C.c_int, [_H, C.c_int, _PD]),
    "swk_add_rate_operator": (C.c_int, [_H, _D, _D, _PD, _PI, _I, C.POINTER(C.c_int)]),
    "swk_set_rate": (C.c_int, [_H, C.c_int, _D, _D]),
    "swk_clear_rate_operators": (C.c_int, [_H]),
    "swk_register_cells": (C.c_int, [_H, _PI, _I, C.POINTER(C.c_int)]),
    "swk_gather_set": (C.c_int, [_H, C.c_int, _PD]),
    "swk_scatter_set": (C.c_int, [_H, C.c_int, _PD]),
    "swk_update_ghosts_async": (C.c_int, [_H]),
    "swk_set_explicit_forcing": (C.c_int, [_H, _PD, _PD, _PD, _I]),
    "swk_set_rate_dynamic": (C.c_int, [_H, C.c_int, C.c_int]),
    "swk_set_boundary_values_substep": (C.c_int, [_H, C.c_int, C.c_int, _PD]),
    "swk_set_boundary_table": (C.c_int, [_H, C.c_int, _I, _I, _PD]),
    "swk_set_boundary_table_frame": (C.c_int, [_H, C.c_int, _I, _PD]),
    "swk_step_begin": (C.c_int, [_H, _D, _D]),
    "swk_step_first": (C.c_int, [_H, C.POINTER(SwkEvolveResult)]),
    "swk_step_rest": (C.c_int, [_H]),
    "swk_step_end": (C.c_int, [_H, C.POINTER(SwkEvolveResult)]),
    "swk_gather_centroids": (C.c_int, [_H, _PI, _I, _PD]),
    "swk_scatter_centroids": (C.c_int, [_H, _PI, _I, _PD]),
    "swk_scatter_bed": (C.c_int, [_H, _PI, _I, _PD]),
    "swk_add_fractional_step_volume": (C.c_int, [_H, _D]),
    "swk_set_local_ghost_copy": (C.c_int, [_H, _PI, _PI, _I]),
    "swk_set_time": (C.c_int, [_H, _D]),
    "swk_protect": (C.c_int, [_H, _PD]),
    "swk_extrapolate_second_order_edge_sw": (C.c_int, [_H]),
    "swk_distribute_to_vertices_and_edges": (C.c_int, [_H]),
    "swk_update_boundary": (C.c_int, [_H]),
    "swk_compute_fluxes": (C.c_int, [_H, C.c_int, _PD]),
    "swk_update_conserved_quantities": (C.c_int, [_H, _D, _PI]),
    "swk_backup_conserved_quantities": (C.c_int, [_H]),
    "swk_saxpy_conserved_quantities": (C.c_int, [_H, _D, _D, _D]),
    "swk_update_ghosts": (C.c_int, [_H]),
    "swk_apply_fractional_steps": (C.c_int, [_H, _D]),
    "swk_get_statistics": (C.c_int, [_H, C.POINTER(SwkEvolveResult)]),
    "swk_evolve": (C.c_int, [_H, _D, _D, _I, C.POINTER(SwkEvolveResult)]),
    "swk_reset_yield_statistics": (C.c_int, [_H]),
    "swk_run_steps": (C.c_int, [_H, _I, C.c_int, C.POINTER(C.c_float)]),
    "swk_kernel_timing": (C.c_int, [_H, _PD, _PI]),
    "swk_stream": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "swk_synchronize": (C.c_int, [_H]),
    "swk_pin_host_buffer": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "swk_unpin_host_buffer": (C.c_int, [_H, C.c_void_p]),
    "swk_kernel_launch_count": (C.c_int, [_H, _PI]),
    "swk_bytes_per_triangle_step": (C.c_int, [_H, _PD, _PD]),
    "swk_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "swk_comm_init": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int]),
    "swk_set_halo": (C.c_int, [_H, C.c_int, C.POINTER(C.c_int), _PI, C.POINTER(_PI), _PI, C.POINTER(_PI)]),
    "swk_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "swk_comm_destroy": (C.c_int, [C.c_void_p]),
    "swk_comm_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]),
    "swk_comm_attach": (C.c_int, [_H, C.c_void_p]),
    "swk_call_open": (C.c_int, [C.POINTER(SwkHostView), C.c_int, C.POINTER(_H)]),
    "swk_call_close": (C.c_int, [_H]),
    "swk_call_compute_fluxes_ext_central": (C.c_int, [_H, C.POINTER(SwkHostView), _D, C.c_int, _PD]),
    "swk_call_extrapolate_second_order_edge_sw": (C.c_int, [_H, C.POINTER(SwkHostView)]),
    "swk_call_protect_new": (C.c_int, [_H, C.POINTER(SwkHostView), _PD]),
    "swk_call_fix_negative_cells": (C.c_int, [_H, C.POINTER(SwkHostView), _PI]),
    "swk_call_manning_friction_flat": (C.c_int, [C.c_int, _D, _D, _I] + [_PD] * 7),
    "swk_call_manning_friction_sloped": (C.c_int, [C.c_int, _D, _D, _I] + [_PD] * 8),
    "swk_call_update": (C.c_int, [C.c_int, _I, _D, _PD, _PD, _PD]),
    "swk_call_backup_centroid_values": (C.c_int, [C.c_int, _I, _PD, _PD]),
    "swk_call_saxpy_centroid_values": (C.c_int, [C.c_int, _I, _D, _D, _PD, _PD]),
}

_lib = None


def load_library():
    """Load libswk.so and bind every declared symbol.  Raises when the library
    has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SwkError(-1, "%s is missing: the CUDA extension was not built and there is no CPU "
                           "fallback (run anuga_core_b200.build.build())" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        f = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib


def _check(code):
    if code != SWK_OK:
        msg = load_library().swk_last_error()
        raise SwkError(code, msg.decode() if msg else "")


def device_count():
    n = C.c_int(0)
    code = load_library().swk_device_count(C.byref(n))
    if code != SWK_OK:
        return 0
    return n.value


def _pd(a):
    return a.ctypes.data_as(_PD) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(_PI) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def make_params(p):
    """p: mapping with the reference's attribute names."""
    sp = SwkParams()
    method = p.get("timestepping_method", "rk2")
    sp.timestepping_method = {"euler": 1, "rk2": 2, "rk3": 3, 1: 1, 2: 2, 3: 3}[method]
    fixed = p.get("fixed_flux_timestep")
    sp.fixed_flux_timestep = -1.0 if fixed is None else float(fixed)
    for k in ("epsilon", "H0", "g", "minimum_allowed_height", "maximum_allowed_speed",
              "evolve_max_timestep", "evolve_min_timestep", "beta_w", "beta_w_dry", "beta_uh",
              "beta_uh_dry", "beta_vh", "beta_vh_dry", "CFL"):
        setattr(sp, k, float(p[k]))
    sp.extrapolate_velocity_second_order = int(bool(p.get("extrapolate_velocity_second_order", True)))
    sp.low_froude = int(p.get("low_froude", 0))
    sp.use_sloped_mannings = int(bool(p.get("use_sloped_mannings", False)))
    sp.max_smallsteps = int(p.get("max_smallsteps", 50))
    sp.default_order = int(p.get("default_order", 2))
    sp.ghost_layer_width = int(p.get("ghost_layer_width", 2))
    sp.centroid_transmissive_bc = int(bool(p.get("centroid_transmissive_bc", False)))
    sp.track_max_speed = int(bool(p.get("track_max_speed", False)))
    return sp


class _MeshArrays:
    """Keeps the contiguous numpy arrays alive while a SwkMesh points at them."""

    def __init__(self, m, permutation=None):
        self.keep = {}
        sm = self.struct = SwkMesh()
        N = int(np.asarray(m["areas"]).shape[0])
        M = int(np.asarray(m["boundary_cells"]).shape[0])
        sm.number_of_elements = N
        sm.boundary_length = M
        for k in ("neighbours", "neighbour_edges", "surrogate_neighbours", "number_of_boundaries",
                  "tri_full_flag", "boundary_cells", "boundary_edges"):
            a = self.keep[k] = _i64(m[k])
            setattr(sm, k, _pi(a))
        for k in ("normals", "edgelengths", "radii", "areas", "centroid_coordinates", "edge_coordinates"):
            a = self.keep[k] = _f64(m[k])
            setattr(sm, k, _pd(a))
        if m.get("vertex_coordinates") is not None:
            a = self.keep["vertex_coordinates"] = _f64(m["vertex_coordinates"])
            sm.vertex_coordinates = _pd(a)
        if m.get("edge_flux_type") is not None and int(m.get("number_of_riverwall_edges", 0)) > 0:
            sm.number_of_riverwall_edges = int(m["number_of_riverwall_edges"])
            sm.ncol_riverwall_hydraulic_properties = int(m.get("ncol_riverwall_hydraulic_properties", 5))
            for k in ("edge_flux_type", "edge_river_wall_counter", "riverwall_rowIndex"):
                a = self.keep[k] = _i64(m[k])
                setattr(sm, k, _pi(a))
            for k in ("riverwall_elevation", "riverwall_hydraulic_properties"):
                a = self.keep[k] = _f64(m[k])
                setattr(sm, k, _pd(a))
        if permutation is not None:
            a = self.keep["permutation"] = _i64(permutation)
            sm.permutation = _pi(a)
        self.N, self.M = N, M


class DeviceDomain:
    """One swk_domain handle: arrays resident in HBM, time loop on the device."""

    def __init__(self, mesh_arrays, params, device=0, permutation=None):
        self.lib = load_library()
        self._mesh = _MeshArrays(mesh_arrays, permutation)
        self.N, self.M = self._mesh.N, self._mesh.M
        self.params = make_params(params)
        h = _H()
        _check(self.lib.swk_create(C.byref(self._mesh.struct), C.byref(self.params), int(device), C.byref(h)))
        self.h = h
        self.device = int(device)
        self._mesh = None      # host mesh arrays are only borrowed during swk_create
        self._segments = {}
        self._pinned = {}

    def close(self):
        if getattr(self, "h", None):
            self.unpin_all()
            self.lib.swk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters / quantities ------------------------------------------
    def set_params(self, params):
        self.params = make_params(params)
        _check(self.lib.swk_set_params(self.h, C.byref(self.params)))

    def set_quantity(self, name, array):
        a = _f64(array).reshape(-1)
        _check(self.lib.swk_set_quantity(self.h, Q[name], _pd(a), a.size))

    def get_quantity(self, name, out=None):
        q = Q[name]
        n = self.N
        shape = (self.N,)
        if 10 <= q < 30:
            n, shape = 3 * self.N, (self.N, 3)
        elif 30 <= q < 40:
            n, shape = self.M, (self.M,)
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        flat = out.reshape(-1)
        assert flat.size == n and flat.flags.c_contiguous and flat.dtype == np.float64
        _check(self.lib.swk_get_quantity(self.h, q, _pd(flat), n))
        return out

    # -- boundaries / operators / ghosts -------------------------------------
    def set_boundary_segment(self, segment, kind, ids, values=(0.0, 0.0, 0.0)):
        ids = _i64(ids)
        v = (C.c_double * 3)(*[float(x) for x in values])
        _check(self.lib.swk_set_boundary_segment(self.h, int(segment), int(kind), _pi(ids), ids.size, v))

    def set_boundary_values(self, segment, values):
        v = (C.c_double * 3)(*[float(x) for x in values])
        _check(self.lib.swk_set_boundary_values(self.h, int(segment), v))

    def add_rate_operator(self, rate=0.0, factor=1.0, rate_array=None, indices=None):
        op = C.c_int(-1)
        ra = _f64(rate_array) if rate_array is not None else None
        idx = _i64(indices) if indices is not None else None
        _check(self.lib.swk_add_rate_operator(self.h, float(rate), float(factor), _pd(ra), _pi(idx),
                                              0 if idx is None else idx.size, C.byref(op)))
        return op.value

    def set_explicit_forcing(self, fs, fx, fy):
        """per-triangle additions to the explicit updates of stage, xmomentum, ymomentum (None: switch off)"""
        if fs is None and fx is None and fy is None:
            _check(self.lib.swk_set_explicit_forcing(self.h, None, None, None, 0))
            return
        arrs = [np.ascontiguousarray(a if a is not None else np.zeros(self.N), dtype=np.float64) for a in (fs, fx, fy)]
        _check(self.lib.swk_set_explicit_forcing(self.h, _pd(arrs[0]), _pd(arrs[1]), _pd(arrs[2]), arrs[0].size))

    def clear_rate_operators(self):
        _check(self.lib.swk_clear_rate_operators(self.h))

    def set_rate(self, op_id, rate, factor=1.0):
        _check(self.lib.swk_set_rate(self.h, int(op_id), float(rate), float(factor)))

    _SET_LIMIT = 1 << 16      # id lists up to this length are registered once and reused (inlets, enquiry cells)

    def _use_set(self, ids):
        # (a caller that moves a different id list every time would only pile up registrations)
        cache = self.__dict__.get("_cell_sets", {})
        return 0 < ids.size <= self._SET_LIMIT and (len(cache) < 256 or ids.tobytes() in cache)

    def _cell_set(self, ids):
        """id of the registered cell set for this id list (include/swk.h: swk_register_cells)"""
        cache = self.__dict__.setdefault("_cell_sets", {})
        key = ids.tobytes()
        sid = cache.get(key)
        if sid is None:
            c = C.c_int(-1)
            _check(self.lib.swk_register_cells(self.h, _pi(ids), ids.size, C.byref(c)))
            sid = cache[key] = c.value
        return sid

    def gather_centroids(self, ids):
        """(n,4) {stage, xmomentum, ymomentum, elevation} of the listed triangles"""
        ids = _i64(ids)
        out = np.empty((ids.size, 4), dtype=np.float64)
        if self._use_set(ids):
            _check(self.lib.swk_gather_set(self.h, self._cell_set(ids), _pd(out)))
        else:
            _check(self.lib.swk_gather_centroids(self.h, _pi(ids), ids.size, _pd(out)))
        return out

    def scatter_centroids(self, ids, values):
        """new {stage, xmomentum, ymomentum} of the listed triangles; queued in stream order"""
        ids = _i64(ids)
        v = _f64(values).reshape(ids.size, 3)
        if self._use_set(ids):
            _check(self.lib.swk_scatter_set(self.h, self._cell_set(ids), _pd(v)))
        else:
            _check(self.lib.swk_scatter_centroids(self.h, _pi(ids), ids.size, _pd(v)))

    def scatter_bed(self, ids, values):
        ids = _i64(ids)
        v = _f64(values).reshape(ids.size)
        _check(self.lib.swk_scatter_bed(self.h, _pi(ids), ids.size, _pd(v)))

    def add_fractional_step_volume(self, volume):
        _check(self.lib.swk_add_fractional_step_volume(self.h, float(volume)))

    def set_local_ghost_copy(self, full_ids, ghost_ids):
        f, g = _i64(full_ids), _i64(ghost_ids)
        _check(self.lib.swk_set_local_ghost_copy(self.h, _pi(f), _pi(g), f.size))

    # -- individual passes ------------------------------------------------------
    def set_time(self, t):
        _check(self.lib.swk_set_time(self.h, float(t)))

    def protect(self):
        me = C.c_double(0.0)
        _check(self.lib.swk_protect(self.h, C.byref(me)))
        return me.value

    def extrapolate_second_order_edge_sw(self):
        _check(self.lib.swk_extrapolate_second_order_edge_sw(self.h))

    def distribute_to_vertices_and_edges(self):
        _check(self.lib.swk_distribute_to_vertices_and_edges(self.h))

    def update_boundary(self):
        _check(self.lib.swk_update_boundary(self.h))

    def compute_fluxes(self, substep=0):
        ft = C.c_double(0.0)
        _check(self.lib.swk_compute_fluxes(self.h, int(substep), C.byref(ft)))
        return ft.value

    def update_conserved_quantities(self, timestep):
        n = C.c_int64(0)
        _check(self.lib.swk_update_conserved_quantities(self.h, float(timestep), C.byref(n)))
        return n.value

    def backup_conserved_quantities(self):
        _check(self.lib.swk_backup_conserved_quantities(self.h))

    def saxpy_conserved_quantities(self, a, b, divide_by=1.0):
        _check(self.lib.swk_saxpy_conserved_quantities(self.h, float(a), float(b), float(divide_by)))

    def update_ghosts(self):
        _check(self.lib.swk_update_ghosts(self.h))

    def update_ghosts_async(self):
        _check(self.lib.swk_update_ghosts_async(self.h))

    def apply_fractional_steps(self, timestep):
        _check(self.lib.swk_apply_fractional_steps(self.h, float(timestep)))

    def get_statistics(self):
        r = SwkEvolveResult()
        _check(self.lib.swk_get_statistics(self.h, C.byref(r)))
        return r

    # -- time loop --------------------------------------------------------------------
    def evolve(self, relative_yieldtime, relative_finaltime=None, max_steps=0):
        r = SwkEvolveResult()
        ft = -1.0 if relative_finaltime is None else float(relative_finaltime)
        _check(self.lib.swk_evolve(self.h, float(relative_yieldtime), ft, int(max_steps), C.byref(r)))
        return r

    # -- host-paced time loop (time-dependent boundary values / rates) -------------------------
    def step_begin(self, relative_yieldtime, relative_finaltime=None):
        ft = -1.0 if relative_finaltime is None else float(relative_finaltime)
        _check(self.lib.swk_step_begin(self.h, float(relative_yieldtime), ft))

    def step_first(self):
        r = SwkEvolveResult()
        _check(self.lib.swk_step_first(self.h, C.byref(r)))
        return r

    def step_rest(self):
        _check(self.lib.swk_step_rest(self.h))

    def step_end(self):
        r = SwkEvolveResult()
        _check(self.lib.swk_step_end(self.h, C.byref(r)))
        return r

    def set_boundary_values_substep(self, segment, substep, values):
        v = (_D * 3)(*[float(x) for x in values])
        _check(self.lib.swk_set_boundary_values_substep(self.h, int(segment), int(substep), v))

    def set_boundary_table(self, segment, frames):
        f = np.ascontiguousarray(frames, dtype=np.float64)
        assert f.ndim == 3 and f.shape[2] == 3
        _check(self.lib.swk_set_boundary_table(self.h, int(segment), f.shape[0], f.shape[1], _pd(f)))

    def set_boundary_table_frame(self, segment, frame, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        _check(self.lib.swk_set_boundary_table_frame(self.h, int(segment), int(frame), _pd(v)))

    def set_rate_dynamic(self, op_id, dynamic=True):
        _check(self.lib.swk_set_rate_dynamic(self.h, int(op_id), int(bool(dynamic))))

    def run_steps(self, n_steps, per_kernel=False):
        """exactly n_steps timesteps, CUDA-event timed on the library stream -> elapsed ms"""
        ms = C.c_float(0.0)
        _check(self.lib.swk_run_steps(self.h, int(n_steps), int(bool(per_kernel)), C.byref(ms)))
        return float(ms.value)

    KERNEL_NAMES = ("extrapolate", "flux", "update", "flux_update")

    def kernel_timing(self):
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        _check(self.lib.swk_kernel_timing(self.h, ms, n))
        return {name: (ms[i], n[i]) for i, name in enumerate(self.KERNEL_NAMES)}

    def reset_yield_statistics(self):
        _check(self.lib.swk_reset_yield_statistics(self.h))

    def synchronize(self):
        _check(self.lib.swk_synchronize(self.h))

    def pin(self, array):
        """page-lock a long-lived numpy buffer (idempotent)"""
        key = array.ctypes.data
        if key not in self._pinned:
            _check(self.lib.swk_pin_host_buffer(self.h, C.c_void_p(key), array.nbytes))
            self._pinned[key] = array.nbytes

    def unpin_all(self):
        for key in list(self._pinned):
            self.lib.swk_unpin_host_buffer(self.h, C.c_void_p(key))
        self._pinned = {}

    def stream(self):
        s = C.c_void_p()
        _check(self.lib.swk_stream(self.h, C.byref(s)))
        return s.value

    def kernel_launch_count(self):
        n = C.c_int64(0)
        _check(self.lib.swk_kernel_launch_count(self.h, C.byref(n)))
        return n.value

    def bytes_per_triangle_step(self):
        a, b = C.c_double(0), C.c_double(0)
        _check(self.lib.swk_bytes_per_triangle_step(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- multi-GPU --------------------------------------------------------------------------
    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        _check(load_library().swk_nccl_unique_id(buf))
        return bytes(buf.raw)

    def comm_init(self, unique_id, rank, nranks):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _check(self.lib.swk_comm_init(self.h, buf, int(rank), int(nranks)))

    def comm_attach(self, nccl_comm):
        """Join the process-level communicator (NcclComm below)."""
        _check(self.lib.swk_comm_attach(self.h, nccl_comm.h))

    def set_halo(self, send, recv):
        """send / recv: {peer_rank: ids (caller numbering)}; both sides list ids sorted by
        global id so that no ids travel (distribute_mesh.py:1128-1170)."""
        peers = sorted(set(send) | set(recv))
        n = len(peers)
        ranks = (C.c_int * max(n, 1))(*peers)
        s_arr = [_i64(send.get(p, [])) for p in peers]
        r_arr = [_i64(recv.get(p, [])) for p in peers]
        sc = _i64([a.size for a in s_arr])
        rc = _i64([a.size for a in r_arr])
        sp = (_PI * max(n, 1))(*[_pi(a) for a in s_arr])
        rp = (_PI * max(n, 1))(*[_pi(a) for a in r_arr])
        _check(self.lib.swk_set_halo(self.h, n, ranks, _pi(sc), sp, _pi(rc), rp))


class NcclComm:
    """Process-level NCCL communicator of libswk (include/swk.h: swk_comm_*): the device time loops of
    the attached domains and the small host-level collectives share it; nothing but NCCL is needed."""
    SUM, MIN, MAX = 0, 1, 2

    def __init__(self, unique_id, rank, nranks, device):
        self.lib = load_library()
        self.rank, self.nranks, self.device = int(rank), int(nranks), int(device)
        buf = C.create_string_buffer(bytes(unique_id), 128)
        h = C.c_void_p()
        _check(self.lib.swk_comm_create(buf, self.rank, self.nranks, self.device, C.byref(h)))
        self.h = h

    def allreduce(self, array, op):
        """In-place all-reduce of a contiguous float64 or int64 numpy array."""
        a = array
        if a.dtype == np.float64:
            dt = 0
        elif a.dtype == np.int64:
            dt = 1
        else:
            raise TypeError("allreduce: float64 or int64 arrays only")
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("allreduce: array must be C-contiguous")
        _check(self.lib.swk_comm_allreduce(self.h, a.ctypes.data_as(C.c_void_p), a.size, dt, int(op)))
        return a

    def close(self):
        if self.h:
            self.lib.swk_comm_destroy(self.h)
            self.h = None


def build_neighbour_structure_native(triangles, number_of_nodes):
    """neighbours, neighbour_edges, number_of_boundaries through libswk's host-side helper
    (the reference also does this natively, neighbour_table.cpp); None when the library is absent."""
    try:
        lib = load_library()
    except (SwkError, OSError):
        return None
    tri = _i64(triangles)
    N = tri.shape[0]
    nb = np.empty((N, 3), dtype=np.int64)
    ne = np.empty((N, 3), dtype=np.int64)
    nob = np.empty(N, dtype=np.int64)
    code = lib.swk_build_neighbour_structure(N, int(number_of_nodes), _pi(tri), _pi(nb), _pi(ne), _pi(nob))
    if code != SWK_OK:
        raise Exception(lib.swk_last_error().decode())
    return nb, ne, nob


def mesh_geometry_native(nodes, triangles, use_inscribed_circle=False):
    """(vertex_coordinates, areas, normals, edgelengths, centroid_coordinates, radii,
    edge_midpoint_coordinates, first_degenerate) through libswk's host-side helper; None when the
    library is absent (the numpy formulation in mesh.py gives the same bits)."""
    try:
        lib = load_library()
    except (SwkError, OSError):
        return None
    nodes = _f64(nodes)
    tri = _i64(triangles)
    N = tri.shape[0]
    V = np.empty((3 * N, 2)); areas = np.empty(N); normals = np.empty((N, 6)); el = np.empty((N, 3))
    cc = np.empty((N, 2)); radii = np.empty(N); E = np.empty((3 * N, 2))
    bad = np.full(1, -1, dtype=np.int64)
    code = lib.swk_mesh_geometry(N, nodes.shape[0], _pd(nodes), _pi(tri), int(bool(use_inscribed_circle)),
                                 _pd(V), _pd(areas), _pd(normals), _pd(el), _pd(cc), _pd(radii), _pd(E), _pi(bad))
    if code != SWK_OK:
        raise Exception(lib.swk_last_error().decode())
    return V, areas, normals, el, cc, radii, E, int(bad[0])
