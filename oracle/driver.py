"""TEST INFRASTRUCTURE ONLY - never imported by the product.

Host time loop of the oracle: a numpy restatement of the reference's evolve
sequence around the C kernels of ``liboracle.so`` (the port, sw_oracle.c) or
``_ref/libanuga_ref*.so`` (the reference's own C sources, see Makefile).

Restated from /root/reference/anuga:
  abstract_2d_finite_volumes/generic_domain.py
      _evolve_base :1715-1912, evolve_one_euler_step :1914-1972,
      evolve_one_rk2_step :1974-2051, evolve_one_rk3_step :2053-2179,
      update_boundary :2288-2306, update_timestep :2349-2415,
      update_ghosts :2448-2469
  shallow_water/shallow_water_domain.py
      compute_fluxes :1828-1877, distribute_to_vertices_and_edges :1882-1915,
      update_conserved_quantities :2089-2163
  shallow_water/friction.py :22-73 (manning_friction_implicit)
  shallow_water/boundaries.py :235-288, 477-517, 616-635 and
  abstract_2d_finite_volumes/generic_boundary_conditions.py :173-264, 370-411
  operators/rate_operators.py :149-269, operators/boundary_flux_integral_operator.py :44-62

The input is a *scenario*: a plain dict of numpy arrays and scalars (see
``REQUIRED_ARRAYS``); the oracle has no dependency on the product package.
"""
import ctypes as C
import os

import math

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBS = {
    "port": os.path.join(HERE, "liboracle.so"),
    "ref": os.path.join(HERE, "_ref", "libanuga_ref.so"),
    "ref_fma": os.path.join(HERE, "_ref", "libanuga_ref_fma.so"),
}
PREFIX = {"port": "orc_", "ref": "ref_", "ref_fma": "ref_"}

_I = C.c_int64
_D = C.c_double
_PI = C.POINTER(C.c_int64)
_PD = C.POINTER(C.c_double)


class OrcDomain(C.Structure):
    """ctypes mirror of oracle/orc_domain.h"""
    _fields_ = [
        ("number_of_elements", _I), ("boundary_length", _I),
        ("number_of_riverwall_edges", _I), ("ncol_riverwall_hydraulic_properties", _I),
        ("extrapolate_velocity_second_order", _I), ("low_froude", _I),
        ("timestep_fluxcalls", _I), ("optimise_dry_cells", _I),
        ("epsilon", _D), ("H0", _D), ("g", _D), ("minimum_allowed_height", _D),
        ("maximum_allowed_speed", _D), ("evolve_max_timestep", _D),
        ("beta_w", _D), ("beta_w_dry", _D), ("beta_uh", _D), ("beta_uh_dry", _D),
        ("beta_vh", _D), ("beta_vh_dry", _D),
        ("neighbours", _PI), ("neighbour_edges", _PI), ("surrogate_neighbours", _PI),
        ("number_of_boundaries", _PI), ("tri_full_flag", _PI), ("edge_flux_type", _PI),
        ("edge_river_wall_counter", _PI),
        ("normals", _PD), ("edgelengths", _PD), ("radii", _PD), ("areas", _PD),
        ("centroid_coordinates", _PD), ("edge_coordinates", _PD), ("vertex_coordinates", _PD),
        ("riverwall_elevation", _PD), ("riverwall_rowIndex", _PI),
        ("riverwall_hydraulic_properties", _PD),
        ("stage_centroid_values", _PD), ("xmom_centroid_values", _PD), ("ymom_centroid_values", _PD),
        ("bed_centroid_values", _PD), ("height_centroid_values", _PD), ("friction_centroid_values", _PD),
        ("stage_edge_values", _PD), ("xmom_edge_values", _PD), ("ymom_edge_values", _PD),
        ("bed_edge_values", _PD), ("height_edge_values", _PD),
        ("stage_vertex_values", _PD), ("xmom_vertex_values", _PD), ("ymom_vertex_values", _PD),
        ("bed_vertex_values", _PD), ("height_vertex_values", _PD),
        ("stage_boundary_values", _PD), ("xmom_boundary_values", _PD), ("ymom_boundary_values", _PD),
        ("stage_explicit_update", _PD), ("xmom_explicit_update", _PD), ("ymom_explicit_update", _PD),
        ("stage_semi_implicit_update", _PD), ("xmom_semi_implicit_update", _PD),
        ("ymom_semi_implicit_update", _PD),
        ("max_speed", _PD), ("x_centroid_work", _PD), ("y_centroid_work", _PD),
        ("boundary_flux_sum", _PD),
    ]


REQUIRED_ARRAYS = [
    "neighbours", "neighbour_edges", "surrogate_neighbours", "number_of_boundaries",
    "tri_full_flag", "normals", "edgelengths", "radii", "areas",
    "centroid_coordinates", "edge_coordinates", "vertex_coordinates",
    "boundary_cells", "boundary_edges",
    "stage_centroid_values", "xmom_centroid_values", "ymom_centroid_values",
    "bed_centroid_values", "friction_centroid_values",
]

DEFAULT_PARAMS = dict(
    g=9.8, epsilon=1.0e-12, H0=1.0e-5, minimum_allowed_height=1.0e-5,
    maximum_allowed_speed=0.0, evolve_max_timestep=1000.0, evolve_min_timestep=1.0e-6,
    max_smallsteps=50, CFL=1.0, timestepping_method="rk2",
    beta_w=1.0, beta_w_dry=0.0, beta_uh=1.0, beta_uh_dry=0.0, beta_vh=1.0, beta_vh_dry=0.0,
    extrapolate_velocity_second_order=1, low_froude=0, optimise_dry_cells=0,
    sloped_mannings=False, fixed_flux_timestep=None, ghost_layer_width=2,
    centroid_transmissive_bc=False, default_order=2,
)

_FLUXCALLS = {"euler": 1, "rk2": 2, "rk3": 3}


def load(backend):
    path = LIBS[backend]
    if not os.path.exists(path):
        raise OSError("oracle library %s missing: run `make -C oracle`" % path)
    lib = C.CDLL(path)
    p = PREFIX[backend]
    PDOM = C.POINTER(OrcDomain)
    sig = {
        "compute_fluxes": (_D, [PDOM, _D, _I]),
        "protect": (_D, [PDOM]),
        "extrapolate": (_I, [PDOM]),
        "fix_negative_cells": (_I, [PDOM]),
        "manning_friction_flat": (None, [_D, _D, _I] + [_PD] * 7),
        "manning_friction_sloped": (None, [_D, _D, _I] + [_PD] * 8),
        "update": (_I, [_I, _D, _PD, _PD, _PD]),
        "backup_centroid_values": (None, [_I, _PD, _PD]),
        "saxpy_centroid_values": (None, [_I, _D, _D, _PD, _PD]),
    }
    fns = {}
    for name, (res, args) in sig.items():
        f = getattr(lib, p + name)
        f.restype = res
        f.argtypes = args
        fns[name] = f
    return fns


def _pd(a):
    return a.ctypes.data_as(_PD)


def _pi(a):
    return a.ctypes.data_as(_PI)


class OracleDomain:
    """CPU oracle of one domain.  ``backend`` selects the kernel library."""

    def __init__(self, scenario, backend="port"):
        self.backend = backend
        self.fn = load(backend)
        sc = scenario
        P = dict(DEFAULT_PARAMS)
        P.update(sc.get("params", {}))
        self.P = P
        for name in REQUIRED_ARRAYS:
            if name not in sc:
                raise KeyError("scenario lacks %r" % name)
        f64 = lambda a: np.array(a, dtype=np.float64, order="C", copy=True)
        i64 = lambda a: np.array(a, dtype=np.int64, order="C", copy=True)
        N = self.N = int(np.asarray(sc["areas"]).shape[0])
        M = self.M = int(np.asarray(sc["boundary_cells"]).shape[0])
        for name in ("neighbours", "neighbour_edges", "surrogate_neighbours",
                     "number_of_boundaries", "tri_full_flag", "boundary_cells", "boundary_edges"):
            setattr(self, name, i64(sc[name]))
        for name in ("normals", "edgelengths", "radii", "areas", "centroid_coordinates",
                     "edge_coordinates", "vertex_coordinates"):
            setattr(self, name, f64(sc[name]))
        self.stage_c = f64(sc["stage_centroid_values"])
        self.xmom_c = f64(sc["xmom_centroid_values"])
        self.ymom_c = f64(sc["ymom_centroid_values"])
        self.bed_c = f64(sc["bed_centroid_values"])
        self.friction_c = f64(sc["friction_centroid_values"])
        self.height_c = np.zeros(N)
        z3 = lambda: np.zeros((N, 3))
        self.stage_e, self.xmom_e, self.ymom_e, self.bed_e, self.height_e = z3(), z3(), z3(), z3(), z3()
        self.stage_v, self.xmom_v, self.ymom_v, self.height_v = z3(), z3(), z3(), z3()
        self.bed_v = f64(sc["bed_vertex_values"]) if "bed_vertex_values" in sc else \
            np.repeat(self.bed_c[:, None], 3, axis=1).copy()
        if "stage_vertex_values" in sc:
            self.stage_v = f64(sc["stage_vertex_values"])
        self.stage_b, self.xmom_b, self.ymom_b = np.zeros(M), np.zeros(M), np.zeros(M)
        self.stage_eu, self.xmom_eu, self.ymom_eu = np.zeros(N), np.zeros(N), np.zeros(N)
        self.stage_siu, self.xmom_siu, self.ymom_siu = np.zeros(N), np.zeros(N), np.zeros(N)
        self.stage_bk, self.xmom_bk, self.ymom_bk = np.zeros(N), np.zeros(N), np.zeros(N)
        self.max_speed = np.zeros(N)
        self.xwork, self.ywork = np.zeros(N), np.zeros(N)
        self.boundary_flux_sum = np.zeros(8)
        self.edge_flux_type = i64(sc["edge_flux_type"]) if "edge_flux_type" in sc else np.zeros(3 * N, np.int64)
        self.edge_river_wall_counter = i64(sc.get("edge_river_wall_counter", np.zeros(3 * N, np.int64)))
        self.riverwall_elevation = f64(sc.get("riverwall_elevation", np.zeros(1)))
        self.riverwall_rowIndex = i64(sc.get("riverwall_rowIndex", np.zeros(1, np.int64)))
        self.riverwall_hydraulic_properties = f64(sc.get("riverwall_hydraulic_properties", np.zeros(5)))
        self.ncol_rw = int(sc.get("ncol_riverwall_hydraulic_properties", 5))

        # boundary specification: tag -> spec tuple; tag_boundary_cells: tag -> ids
        self.boundary_map = dict(sc.get("boundary_map", {}))
        self.tag_boundary_cells = {t: np.asarray(v, dtype=np.int64)
                                   for t, v in sc.get("tag_boundary_cells", {}).items()}
        self.operators = list(sc.get("operators", []))
        self.forcing = list(sc.get("forcing", []))
        # single-process ghost copy (generic_domain.py:2448-2469)
        self.ghost_copy = sc.get("ghost_copy")      # (Idf, Idg) or None

        self.timestepping_method = P["timestepping_method"]
        self.CFL = P["CFL"]
        self.relative_time = 0.0
        self.starttime = 0.0
        self.timestep = 0.0
        self.flux_timestep = 0.0
        self.smallsteps = 0
        self._order_ = P["default_order"]
        self.number_of_steps = 0
        self.number_of_first_order_steps = 0
        self.recorded_min_timestep = P["evolve_max_timestep"]
        self.recorded_max_timestep = P["evolve_min_timestep"]
        self.evolved_called = False
        self.boundary_flux_integral = 0.0
        self.fractional_step_volume_integral = 0.0
        self.timestep_history = []
        self.mass_error = 0.0
        self.num_negative_cells = 0
        self._struct()

    # -- struct --------------------------------------------------------
    def _struct(self):
        P = self.P
        D = self.D = OrcDomain()
        D.number_of_elements = self.N
        D.boundary_length = self.M
        D.number_of_riverwall_edges = int((self.edge_flux_type == 1).sum())
        D.ncol_riverwall_hydraulic_properties = self.ncol_rw
        D.extrapolate_velocity_second_order = int(P["extrapolate_velocity_second_order"])
        D.low_froude = int(P["low_froude"])
        D.timestep_fluxcalls = _FLUXCALLS[self.timestepping_method]
        D.optimise_dry_cells = int(P["optimise_dry_cells"])
        for k in ("epsilon", "H0", "g", "minimum_allowed_height", "maximum_allowed_speed",
                  "evolve_max_timestep", "beta_w", "beta_w_dry", "beta_uh", "beta_uh_dry",
                  "beta_vh", "beta_vh_dry"):
            setattr(D, k, float(P[k]))
        for k in ("neighbours", "neighbour_edges", "surrogate_neighbours", "number_of_boundaries",
                  "tri_full_flag", "edge_flux_type", "edge_river_wall_counter", "riverwall_rowIndex"):
            setattr(D, k, _pi(getattr(self, k)))
        for k in ("normals", "edgelengths", "radii", "areas", "centroid_coordinates",
                  "edge_coordinates", "vertex_coordinates", "riverwall_elevation",
                  "riverwall_hydraulic_properties", "max_speed", "boundary_flux_sum"):
            setattr(D, k, _pd(getattr(self, k)))
        m = {
            "stage_centroid_values": self.stage_c, "xmom_centroid_values": self.xmom_c,
            "ymom_centroid_values": self.ymom_c, "bed_centroid_values": self.bed_c,
            "height_centroid_values": self.height_c, "friction_centroid_values": self.friction_c,
            "stage_edge_values": self.stage_e, "xmom_edge_values": self.xmom_e,
            "ymom_edge_values": self.ymom_e, "bed_edge_values": self.bed_e,
            "height_edge_values": self.height_e,
            "stage_vertex_values": self.stage_v, "xmom_vertex_values": self.xmom_v,
            "ymom_vertex_values": self.ymom_v, "bed_vertex_values": self.bed_v,
            "height_vertex_values": self.height_v,
            "stage_boundary_values": self.stage_b, "xmom_boundary_values": self.xmom_b,
            "ymom_boundary_values": self.ymom_b,
            "stage_explicit_update": self.stage_eu, "xmom_explicit_update": self.xmom_eu,
            "ymom_explicit_update": self.ymom_eu,
            "stage_semi_implicit_update": self.stage_siu, "xmom_semi_implicit_update": self.xmom_siu,
            "ymom_semi_implicit_update": self.ymom_siu,
            "x_centroid_work": self.xwork, "y_centroid_work": self.ywork,
        }
        for k, a in m.items():
            setattr(D, k, _pd(a))

    # -- kernels ---------------------------------------------------------
    def protect(self):
        me = self.fn["protect"](C.byref(self.D))
        self.mass_error += me
        return me

    def extrapolate(self):
        self.fn["extrapolate"](C.byref(self.D))

    def distribute_to_vertices_and_edges(self):
        self.protect()
        self.extrapolate()

    def compute_fluxes(self, substep):
        self.flux_timestep = self.fn["compute_fluxes"](C.byref(self.D),
                                                       self.P["evolve_max_timestep"], substep)

    def compute_forcing_terms(self):
        P = self.P
        if P["sloped_mannings"]:
            self.fn["manning_friction_sloped"](P["g"], P["minimum_allowed_height"], self.N,
                                               _pd(self.vertex_coordinates), _pd(self.stage_c),
                                               _pd(self.bed_v), _pd(self.xmom_c), _pd(self.ymom_c),
                                               _pd(self.friction_c), _pd(self.xmom_siu), _pd(self.ymom_siu))
        else:
            self.fn["manning_friction_flat"](P["g"], P["minimum_allowed_height"], self.N,
                                             _pd(self.stage_c), _pd(self.bed_c), _pd(self.xmom_c),
                                             _pd(self.ymom_c), _pd(self.friction_c),
                                             _pd(self.xmom_siu), _pd(self.ymom_siu))

        for spec in self.forcing:
            if spec[0] == "general_forcing":      # General_forcing.__call__ (forcing.py:400-452): update[k] += rate
                G = spec[1]
                rate = G.current_rate(self.get_time())
                upd = {"stage": self.stage_eu, "xmomentum": self.xmom_eu, "ymomentum": self.ymom_eu}[G.quantity_name]
                if G.exchange_indices is None:
                    upd[:] += rate
                else:
                    for k in G.exchange_indices:
                        upd[k] += rate
            if spec[0] == "wind":                 # Wind_stress.__call__ + assign_windfield_values (forcing.py:133-215)
                W = spec[1]
                t = self.get_time()
                xc = self.centroid_coordinates
                N = self.N

                def field(f):
                    if callable(f):
                        return np.asarray(f(t, xc[:, 0], xc[:, 1]), dtype=np.float64) * np.ones(N)
                    return f * np.ones(N, dtype=np.float64)
                s_vec, phi_vec = field(W.speed), field(W.phi)
                for k in range(N):
                    phi = phi_vec[k] * math.pi / 180.0
                    u = s_vec[k] * math.cos(phi)
                    v = s_vec[k] * math.sin(phi)
                    S = W.const * math.sqrt(u ** 2 + v ** 2)
                    self.xmom_eu[k] += S * u
                    self.ymom_eu[k] += S * v

    def update_conserved_quantities(self):
        dt = self.timestep
        for c, eu, siu in ((self.stage_c, self.stage_eu, self.stage_siu),
                           (self.xmom_c, self.xmom_eu, self.xmom_siu),
                           (self.ymom_c, self.ymom_eu, self.ymom_siu)):
            err = self.fn["update"](self.N, dt, _pd(c), _pd(eu), _pd(siu))
            if err != 0:
                raise RuntimeError("semi-implicit denominator <= 0 (quantity.c:806)")
        self.num_negative_cells += self.fn["fix_negative_cells"](C.byref(self.D))

    def backup_conserved_quantities(self):
        for c, b in ((self.stage_c, self.stage_bk), (self.xmom_c, self.xmom_bk), (self.ymom_c, self.ymom_bk)):
            self.fn["backup_centroid_values"](self.N, _pd(c), _pd(b))

    def saxpy_conserved_quantities(self, a, b):
        for c, bk in ((self.stage_c, self.stage_bk), (self.xmom_c, self.xmom_bk), (self.ymom_c, self.ymom_bk)):
            self.fn["saxpy_centroid_values"](self.N, a, b, _pd(c), _pd(bk))

    # -- boundaries --------------------------------------------------------
    def get_time(self):
        return self.starttime + self.relative_time

    def update_boundary(self):
        t = self.get_time()
        for tag, ids in self.tag_boundary_cells.items():
            spec = self.boundary_map.get(tag)
            if spec is None:
                continue
            kind = spec[0]
            vol = self.boundary_cells[ids]
            edge = self.boundary_edges[ids]
            n1 = self.normals[vol, 2 * edge]
            n2 = self.normals[vol, 2 * edge + 1]
            if kind == "reflective":
                self.stage_b[ids] = self.stage_e[vol, edge]
                q1 = self.xmom_e[vol, edge]
                q2 = self.ymom_e[vol, edge]
                r1 = -q1 * n1 - q2 * n2
                r2 = -q1 * n2 + q2 * n1
                self.xmom_b[ids] = n1 * r1 - n2 * r2
                self.ymom_b[ids] = n2 * r1 + n1 * r2
            elif kind == "dirichlet":
                w, uh, vh = spec[1]
                self.stage_b[ids] = w
                self.xmom_b[ids] = uh
                self.ymom_b[ids] = vh
            elif kind == "time":
                w, uh, vh = spec[1](t)[:3]
                self.stage_b[ids] = w
                self.xmom_b[ids] = uh
                self.ymom_b[ids] = vh
            elif kind == "transmissive":
                if self.P["centroid_transmissive_bc"]:
                    self.stage_b[ids] = self.stage_c[vol]
                    self.xmom_b[ids] = self.xmom_c[vol]
                    self.ymom_b[ids] = self.ymom_c[vol]
                else:
                    self.stage_b[ids] = self.stage_e[vol, edge]
                    self.xmom_b[ids] = self.xmom_e[vol, edge]
                    self.ymom_b[ids] = self.ymom_e[vol, edge]
            elif kind == "transmissive_n_zero_t_set_stage":
                x = float(spec[1](t))
                self.stage_b[ids] = x
                q1 = self.xmom_e[vol, edge]
                q2 = self.ymom_e[vol, edge]
                ndotq = n1 * q1 + n2 * q2
                self.xmom_b[ids] = ndotq * n1
                self.ymom_b[ids] = ndotq * n2
            elif kind == "transmissive_momentum_set_stage":
                self.stage_b[ids] = float(spec[1](t))
                self.xmom_b[ids] = self.xmom_e[vol, edge]
                self.ymom_b[ids] = self.ymom_e[vol, edge]
            elif kind == "transmissive_stage_zero_momentum":
                self.stage_b[ids] = self.stage_e[vol, edge]
                self.xmom_b[ids] = 0.0
                self.ymom_b[ids] = 0.0
            elif kind == "flather_external_stage_zero_velocity":
                # boundaries.py:1207-1266 (evaluate_segment), gravity = anuga.config.g
                value = spec[1](t)
                try:
                    so = float(value)
                except Exception:
                    so = float(value[0])
                sb = self.stage_e[vol, edge]
                xb = self.xmom_e[vol, edge]
                yb = self.ymom_e[vol, edge]
                eb = self.bed_e[vol, edge]
                bed = self.bed_c[vol]
                depth = np.maximum(sb - bed, 0.0)
                so = 0.0 * sb + so
                with np.errstate(divide="ignore", invalid="ignore"):
                    q0_dry = np.where(bed <= so, so, eb)
                    s = (9.8 / depth) ** 0.5
                    ndotq = n1 * xb + n2 * yb
                    w1 = 0.0 - s * so
                    w2 = np.where(ndotq > 0.0, (n2 * xb - n1 * yb) / depth, 0.0 * ndotq)
                    w3 = ndotq / depth + s * sb
                    q0_wet = (w3 - w1) / (2.0 * s)
                    qperp = (w3 + w1) / 2.0 * depth
                    qpar = w2 * depth
                    q1_wet = qperp * n1 + qpar * n2
                    q2_wet = qperp * n2 - qpar * n1
                dry = np.logical_or(depth == 0.0, so > bed)
                self.stage_b[ids] = np.where(dry, q0_dry, q0_wet)
                self.xmom_b[ids] = np.where(dry, 0.0 * xb, q1_wet)
                self.ymom_b[ids] = np.where(dry, 0.0 * yb, q2_wet)
            elif kind == "file":
                # File_boundary.evaluate / Field_boundary.evaluate per edge (generic_boundary_conditions.py:636-700,
                # boundaries.py:1072-1086): the interpolation function at the edge's point, stage + mean_stage
                B = spec[1]
                for m, v, e in zip(ids, vol, edge):
                    q = B.F(t, point_id=B.boundary_indices[(int(v), int(e))])
                    self.stage_b[m] = q[0] + B.mean_stage if B.device_kind == 10 else q[0]
                    self.xmom_b[m] = q[1]
                    self.ymom_b[m] = q[2]
            elif kind == "time_space":
                # Time_space_boundary.evaluate (generic_boundary_conditions.py:480-484) at the edge midpoints
                for m, v, e in zip(ids, vol, edge):
                    x, y = self.edge_coordinates[3 * v + e]
                    q = spec[1](t, x, y)
                    self.stage_b[m], self.xmom_b[m], self.ymom_b[m] = q[0], q[1], q[2]
            elif kind == "dirichlet_discharge":      # boundaries.py:881-885 (evaluate, edge by edge)
                self.stage_b[ids] = spec[1]
                self.xmom_b[ids] = -spec[2] * n1
                self.ymom_b[ids] = -spec[2] * n2
            elif kind == "characteristic_stage":
                # boundaries.py:760-843 (evaluate_segment), gravity = anuga.config.g
                value = spec[1](t)
                try:
                    w_outside = float(value)
                except Exception:
                    w_outside = float(value[0])
                sb = self.stage_e[vol, edge]
                xb = self.xmom_e[vol, edge]
                yb = self.ymom_e[vol, edge]
                eb = self.bed_e[vol, edge]
                w_outside = 0.0 * sb + w_outside
                with np.errstate(divide="ignore", invalid="ignore"):
                    sqrt_g = 9.8 ** 0.5
                    h_inside = np.maximum(sb - eb, 0)
                    uh_inside = n1 * xb + n2 * yb
                    vh_inside = n2 * xb - n1 * yb
                    u_inside = np.where(h_inside > 0.0, uh_inside / h_inside, 0.0)
                    h_outside = np.maximum(w_outside - eb, 0)
                    sqrt_h_inside = h_inside ** 0.5
                    sqrt_h_outside = h_outside ** 0.5
                    h_m = (0.5 * (sqrt_h_inside + sqrt_h_outside) + u_inside / 4.0 / sqrt_g) ** 2
                    u_m = 0.5 * u_inside + sqrt_g * (sqrt_h_inside - sqrt_h_outside)
                    uh_m = h_m * u_m
                    vh_m = np.where(uh_inside > 0.0, vh_inside, 0.0)
                    w_m = h_m + eb
                    dry_test = np.logical_or(h_inside == 0.0, h_outside == 0.0)
                    q1 = uh_m * n1 + vh_m * n2
                    q2 = uh_m * n2 - vh_m * n1
                self.stage_b[ids] = np.where(dry_test, w_outside, w_m)
                self.xmom_b[ids] = np.where(dry_test, 0.0, q1)
                self.ymom_b[ids] = np.where(dry_test, 0.0, q2)
            elif kind == "time_stage_zero_momentum":
                self.stage_b[ids] = float(spec[1](t))
                self.xmom_b[ids] = 0.0
                self.ymom_b[ids] = 0.0
            else:
                raise ValueError("unknown boundary kind %r" % (kind,))

    # -- time stepping -------------------------------------------------------
    def update_ghosts(self):
        if self.ghost_copy is not None:
            Idf, Idg = self.ghost_copy
            for c in (self.stage_c, self.xmom_c, self.ymom_c):
                c[Idg] = c[Idf]

    def update_timestep(self, yieldstep, finaltime):
        P = self.P
        if P["fixed_flux_timestep"] is not None:
            self.flux_timestep = P["fixed_flux_timestep"]
        timestep = min(self.CFL * self.flux_timestep, P["evolve_max_timestep"])
        self.recorded_max_timestep = max(timestep, self.recorded_max_timestep)
        self.recorded_min_timestep = min(timestep, self.recorded_min_timestep)
        if timestep < P["evolve_min_timestep"]:
            self.smallsteps += 1
            if self.smallsteps > P["max_smallsteps"]:
                self.smallsteps = 0
                if self._order_ == 1:
                    raise RuntimeError("Too small timestep %.16f reached even after %d steps of 1 order scheme"
                                       % (timestep, P["max_smallsteps"]))
                else:
                    self._order_ = 1
        else:
            self.smallsteps = 0
            if self._order_ == 1 and P["default_order"] == 2:
                self._order_ = 2
        if self.relative_finaltime is not None and self.relative_time + timestep > self.relative_finaltime:
            timestep = self.relative_finaltime - self.relative_time
        if self.relative_time + timestep > self.relative_yieldtime:
            timestep = self.relative_yieldtime - self.relative_time
        self.timestep = timestep

    def evolve_one_euler_step(self, yieldstep, finaltime):
        self.distribute_to_vertices_and_edges()
        self.update_boundary()
        self.compute_fluxes(0)
        self.compute_forcing_terms()
        self.update_timestep(yieldstep, finaltime)
        self.update_conserved_quantities()

    def evolve_one_rk2_step(self, yieldstep, finaltime):
        self.backup_conserved_quantities()
        self.distribute_to_vertices_and_edges()
        self.update_boundary()
        self.compute_fluxes(0)
        self.compute_forcing_terms()
        self.update_timestep(yieldstep, finaltime)
        self.update_conserved_quantities()
        self.relative_time = self.relative_time + self.timestep
        if self.P["ghost_layer_width"] < 4:
            self.update_ghosts()
        self.distribute_to_vertices_and_edges()
        self.update_boundary()
        self.compute_fluxes(1)
        self.compute_forcing_terms()
        self.update_conserved_quantities()
        self.saxpy_conserved_quantities(0.5, 0.5)

    def evolve_one_rk3_step(self, yieldstep, finaltime):
        self.backup_conserved_quantities()
        initial_time = self.relative_time
        self.distribute_to_vertices_and_edges()
        self.update_boundary()
        self.compute_fluxes(0)
        self.compute_forcing_terms()
        self.update_timestep(yieldstep, finaltime)
        self.update_conserved_quantities()
        self.relative_time = self.relative_time + self.timestep
        self.update_ghosts()
        self.distribute_to_vertices_and_edges()
        self.update_boundary()
        self.compute_fluxes(1)
        self.compute_forcing_terms()
        self.update_conserved_quantities()
        self.saxpy_conserved_quantities(0.25, 0.75)
        self.relative_time = initial_time + self.timestep * 0.5
        self.update_ghosts()
        self.distribute_to_vertices_and_edges()
        self.update_boundary()
        self.compute_fluxes(2)
        self.compute_forcing_terms()
        self.update_conserved_quantities()
        self.saxpy_conserved_quantities(2.0, 1.0)
        for c in (self.stage_c, self.xmom_c, self.ymom_c):
            c[:] = c / 3.0
        self.relative_time = initial_time + self.timestep

    # -- fractional-step operators ---------------------------------------------
    def apply_fractional_steps(self):
        # boundary_flux_integral_operator.py:44-62 (always registered)
        dt = self.timestep
        bfs = self.boundary_flux_sum
        m = self.timestepping_method
        if m == "euler":
            self.boundary_flux_integral += dt * bfs[0]
        elif m == "rk2":
            self.boundary_flux_integral += 0.5 * dt * (bfs[0] + bfs[1])
        else:
            self.boundary_flux_integral += 1.0 / 6.0 * dt * (bfs[0] + bfs[1] + 4.0 * bfs[2])
        bfs[:] = 0.0
        for op in self.operators:
            if op[0] == "rate":
                self._rate_operator(op[1])
            elif op[0] == "inlet":
                self._inlet_operator(op[1])
            elif op[0] == "boyd_box":
                self._boyd_box_operator(op[1])
            elif op[0] == "set_elevation":        # operators/set_elevation.py:116-150 (discontinuous elevation)
                o = op[1]
                ids = slice(None) if o["indices"] is None else np.asarray(o["indices"], dtype=np.int64)
                vt, v = o["value_type"], o["value"]
                x, y = self.centroid_coordinates[ids, 0], self.centroid_coordinates[ids, 1]
                value = v(self.get_time()) if vt == "t" else v(x, y) if vt == "x,y" else \
                    v(x, y, self.get_time()) if vt == "x,y,t" else float(v)
                height = self.stage_c[ids] - self.bed_c[ids]
                self.bed_c[ids] = value
                self.stage_c[ids] = self.bed_c[ids] + height
            elif op[0] == "set_quantity":         # operators/set_quantity.py:76-110
                o = op[1]
                ids = slice(None) if o["indices"] is None else np.asarray(o["indices"], dtype=np.int64)
                vt, v = o["value_type"], o["value"]
                x, y = self.centroid_coordinates[ids, 0], self.centroid_coordinates[ids, 1]
                value = v(self.get_time()) if vt == "t" else v(x, y) if vt == "x,y" else \
                    v(x, y, self.get_time()) if vt == "x,y,t" else float(v)
                {"stage": self.stage_c, "xmomentum": self.xmom_c, "ymomentum": self.ymom_c}[o["quantity"]][ids] = value
            elif op[0] == "boyd_pipe":
                self._boyd_box_operator(op[1], pipe=True)
            elif op[0] == "weir_orifice_trapezoid":
                self._boyd_box_operator(op[1], weir=True)
            else:
                raise ValueError("unknown operator %r" % (op[0],))

    def _rate_operator(self, o):
        """rate_operators.py:149-269.  o: dict(rate=scalar | (N,) array | f(t),
        factor=1.0, indices=None)"""
        dt = self.timestep
        factor = o.get("factor", 1.0)
        rate = o["rate"]
        if callable(rate):
            rate = rate(self.get_time())
        idx = o.get("indices")
        if o.get("rate_xyt") is not None:          # spatial-temporal rate: rate(x, y, t) at the centroids
            c = self.centroid_coordinates.reshape(-1, 2)
            rate = np.asarray(o["rate_xyt"](c[:, 0], c[:, 1], self.get_time()), dtype=np.float64) * np.ones(self.N)
        full = self.tri_full_flag == 1
        if idx is None:
            local_rates = factor * dt * rate * np.ones(self.N) if np.isscalar(rate) else factor * dt * np.asarray(rate)
            sl = slice(None)
        else:
            idx = np.asarray(idx, dtype=np.int64)
            local_rates = factor * dt * rate * np.ones(idx.size) if np.isscalar(rate) else factor * dt * np.asarray(rate)[idx]
            sl = idx
        if np.all(local_rates >= 0.0):
            self.stage_c[sl] = self.stage_c[sl] + local_rates
        else:
            height = self.stage_c[sl] - self.bed_c[sl]
            local_rates = np.maximum(local_rates, -height)
            f = np.where(local_rates < 0.0, (local_rates + height) / (height + 1.0e-10), 1.0)
            self.stage_c[sl] = self.stage_c[sl] + local_rates
            self.xmom_c[sl] = self.xmom_c[sl] * f
            self.ymom_c[sl] = self.ymom_c[sl] * f
        areas = self.areas[sl]
        fsel = full[sl]
        self.fractional_step_volume_integral += float(np.sum((local_rates * areas)[fsel]))

    def _inlet_operator(self, o):
        """structures/inlet_operator.py:78-157 with structures/inlet.py:107-227"""
        idx = np.asarray(o["indices"], dtype=np.int64)
        dt = self.timestep
        t = self.get_time()
        Qf = o["Q"]
        q = (lambda tt: float(Qf(tt))) if callable(Qf) else (lambda tt: float(Qf))
        areas = self.areas[idx]
        total_area = np.sum(areas)
        stages = self.stage_c[idx]
        elev = self.bed_c[idx]
        depths = stages - elev
        current_volume = np.sum(depths * areas)
        Q = 0.5 * (q(t) + q(t + dt))
        volume = Q * dt
        u = self.xmom_c[idx] * depths / (depths * depths + 1.0e-6)
        v = self.ymom_c[idx] * depths / (depths * depths + 1.0e-6)

        def set_momenta():
            d2 = self.stage_c[idx] - elev
            if o.get("velocity") is not None:
                self.xmom_c[idx] = d2 * o["velocity"][0]
                self.ymom_c[idx] = d2 * o["velocity"][1]
            else:
                self.xmom_c[idx] = d2 * u
                self.ymom_c[idx] = d2 * v
            if o.get("zero_velocity"):
                self.xmom_c[idx] = 0.0
                self.ymom_c[idx] = 0.0
        if volume >= 0.0:
            order = stages.argsort()
            summed_areas = np.cumsum(areas[order])
            summed_volume = np.zeros_like(areas)
            summed_volume[1:] = np.cumsum(summed_areas[:-1] * np.diff(stages[order]))
            index = np.nonzero(summed_volume <= volume)[0][-1]
            depth = (volume - summed_volume[index]) / summed_areas[index]
            new = stages.copy()
            new[order[0:index + 1]] = stages[order[index]] + depth
            self.stage_c[idx] = new
            self.fractional_step_volume_integral += volume
            set_momenta()
        elif current_volume + volume >= 0.0:
            self.stage_c[idx] = elev + (current_volume + volume) / total_area
            self.fractional_step_volume_integral += volume
            set_momenta()
        else:
            self.stage_c[idx] = elev + 0.0
            self.fractional_step_volume_integral -= current_volume
            self.xmom_c[idx] = 0.0                  # inlet_operator.py:159-160
            self.ymom_c[idx] = 0.0

    def _boyd_box_operator(self, o, pipe=False, weir=False):
        """structures/structure_operator.py:215-372 (the transfer) around
        structures/boyd_box_operator.py:150-441 (the rating) with the enquiry formulas of
        structures/inlet_enquiry.py:86-158.  `o` carries the resolved geometry (inlet triangle
        ids, enquiry triangle ids, outward unit vectors, barrel length) and the smoothing memory
        (smooth_delta_total_energy, smooth_Q), which this function updates in place."""
        G, VP = 9.8, 1.0e-6                       # anuga/config.py:45, :18
        dt = self.timestep
        stage, bed, xm, ym = self.stage_c, self.bed_c, self.xmom_c, self.ymom_c

        def enquiry(k):
            e = o["enquiry_indices"][k]
            inv = o["invert_elevations"][k]
            inv = bed[e] if inv is None else inv
            depth = max(stage[e] - inv, 0.0)
            wd = stage[e] - bed[e]
            u = wd * xm[e] / (wd ** 2 + VP)
            v = wd * ym[e] / (wd ** 2 + VP)
            if o.get("use_new_velocity_head"):
                n1, n2 = o["outward_vectors"][k]
                head = 0.5 * min(u * n1 + v * n2, 0.0) ** 2 / G
            else:
                head = 0.5 * math.sqrt(u ** 2 + v ** 2) ** 2 / G
            return dict(stage=stage[e], depth=depth, total=head + stage[e], specific=head + depth)

        # ---- discharge_routine ----
        flow_area = None
        if (o["diameter"] if pipe else o["height"]) <= 0.0:
            Q = speed = outlet_depth = 0.0
            i_in, i_out = 0, 1
        else:
            E = [enquiry(0), enquiry(1)]
            key = "total" if o["use_velocity_head"] else "stage"
            delta = E[0][key] - E[1][key]
            if dt > 0.0:
                ts = dt / max(dt, o["smoothing_timescale"], 1.0e-06)
            else:
                ts = 1.0
            o["smooth_delta_total_energy"] = o["smooth_delta_total_energy"] + ts * (delta - o["smooth_delta_total_energy"])
            sm = o["smooth_delta_total_energy"]
            if sm >= 0.0:
                i_in, i_out, delta = 0, 1, sm
            else:
                i_in, i_out, delta = 1, 0, -sm
            if E[i_in]["depth"] > 0.01:
                assert E[i_in]["specific"] >= 0.0
                drive = E[i_in]["specific"] if o["use_velocity_head"] else E[i_in]["depth"]
                rating = self._boyd_pipe_rating if pipe else (self._weir_rating if weir else self._boyd_box_rating)
                Q, speed, outlet_depth, flow_area = rating(o, drive, delta, E[i_out]["depth"])
                sign = np.sign(sm)
                o["smooth_Q"] = o["smooth_Q"] + ts * (Q * sign - o["smooth_Q"])
                if np.sign(o["smooth_Q"]) != sign:
                    Q = 0.0
                else:
                    Q = min(abs(o["smooth_Q"]), Q)
                speed = 0.0 if flow_area == 0 else Q / flow_area
            else:
                Q = speed = outlet_depth = 0.0
        if speed > o["max_velocity"]:
            speed = o["max_velocity"]
            Q = flow_area * speed

        # ---- Structure_operator.__call__ ----
        iin = np.asarray(o["inlet_indices"][i_in], dtype=np.int64)
        iout = np.asarray(o["inlet_indices"][i_out], dtype=np.int64)
        a_in, a_out = self.areas[iin], self.areas[iout]
        A_in, A_out = np.sum(a_in), np.sum(a_out)
        d_old = np.sum((stage[iin] - bed[iin]) * a_in) / A_in
        x_old = np.sum(xm[iin] * a_in) / A_in
        y_old = np.sum(ym[iin] * a_in) / A_in
        dtQd = dt * Q / d_old if d_old > 0.0 else 0.0
        adjust = o["always_use_Q_wetdry_adjustment"] or (d_old * A_in <= Q * dt)
        factor = 1.0 / (1.0 + dtQd / A_in)
        if adjust:
            d_new = d_old * factor
            dt_star = dt * d_new / d_old if d_old > 0.0 else 0.0
        else:
            d_new = d_old - dt * Q / A_in
            dt_star = dt
        if o["use_old_momentum_method"]:
            x_new, y_new = x_old * factor, y_old * factor
        else:
            if d_old > 0.0:
                if adjust:
                    f2 = 1.0 / (1.0 + dtQd * d_new / (d_old * A_in))
                else:
                    f2 = 1.0 / (1.0 + dt * Q / (d_old * A_in))
            else:
                f2 = 0.0
            x_new, y_new = x_old * f2, y_old * f2
        stage[iin] = bed[iin] + d_new
        xm[iin] = x_new
        ym[iin] = y_new
        x_loss = (x_old - x_new) * A_in
        y_loss = (y_old - y_new) * A_in
        extra = Q * dt_star / A_out
        out_dir = -np.asarray(o["outward_vectors"][i_out])
        d_out = np.sum((stage[iout] - bed[iout]) * a_out) / A_out + extra
        if o["use_momentum_jet"]:
            x_out = speed * d_out * out_dir[0]
            y_out = speed * d_out * out_dir[1]
        elif o["zero_outflow_momentum"]:
            x_out = y_out = 0.0
        else:
            x_out = np.sum(xm[iout] * a_out) / A_out + x_loss / A_out
            y_out = np.sum(ym[iout] * a_out) / A_out + y_loss / A_out
        stage[iout] = bed[iout] + d_out
        xm[iout] = x_out
        ym[iout] = y_out

    @staticmethod
    def _boyd_box_rating(o, drive, delta, tail_depth):
        """boyd_box_function (boyd_box_operator.py:265-400): inlet control = the smaller of the
        unsubmerged (weir, E^1.5) and submerged (orifice, D^0.89 E^0.61) ratings; outlet control
        (energy loss over the barrel) caps it when delta < drive."""
        G, VP = 9.8, 1.0e-6
        width, depth, barrels = o["width"], o["height"], o["barrels"]
        bf = 1 - o["blockage"]
        if o["blockage"] >= 1.0:
            return 0.0, 0.0, 0.0, 0.00001
        Qu = 0.544 * G ** 0.5 * bf * width * barrels * drive ** 1.50
        Qs = 0.702 * G ** 0.5 * bf * width * barrels * depth ** 0.89 * drive ** 0.61
        Q = Qu if Qu < Qs else Qs
        clear = bf * width * barrels
        dcrit = (Q ** 2 / G / clear ** 2) ** 0.333333
        if dcrit > depth:
            out_d, area, perim = depth, clear * depth, 2 * (clear + depth)
        else:
            out_d, area, perim = dcrit, clear * dcrit, clear + 2 * dcrit
        if delta < drive:
            if tail_depth > depth:
                out_d, area, perim = depth, clear * depth, 2.0 * (clear + depth)
            rh = area / perim
            vel = math.sqrt(delta / ((o["sum_loss"] / 2 / G) + (o["manning"] ** 2 * o["length"]) / rh ** 1.33333))
            Q = min(Q, area * vel)
        return Q, Q / (area + VP / area), out_d, area

    @staticmethod
    def _weir_rating(o, drive, delta, tail_depth):
        """weir_orifice_trapezoid_function (weir_orifice_trapezoid_operator.py:279-485): trapezoidal opening
        with side slopes z1, z2; critical depth by the reference's Newton iteration"""
        G, VP = 9.8, 1.0e-6
        width, depth, barrels, z1, z2 = o["width"], o["height"], o["barrels"], o["z1"], o["z2"]
        bf = 1 - o["blockage"]
        if o["blockage"] >= 1.0:
            return 0.0, 0.0, 0.0, 0.00001
        Qu = 1.7 * bf * barrels * ((2 * width + depth * (z1 + z2)) / 2) * drive ** 1.50
        Qs = 0.8 * bf * barrels * G ** 0.5 * (0.5 * depth * (2 * width + depth * (z1 + z2))) * drive ** 0.5
        Q = Qu if Qu < Qs else Qs

        def newton(Q):
            dcrit, dyc = 0.00001, 0.001
            while abs(dyc) > 0.00001:
                Tc = bf * barrels * width + (z1 + z2) * dcrit
                Ac = 0.5 * dcrit * (bf * barrels * width + Tc)
                fc = Ac ** 1.5 * Tc ** -0.5 - Q / (9.81 ** 0.5)
                ffc = Ac ** 1.5 * -0.5 * Tc ** -1.5 * (z1 + z2) + Tc ** -0.5 * 1.5 * Ac ** 0.5 * Tc
                dyc = -fc / ffc
                dcrit = dcrit + dyc
            return dcrit
        out_d = newton(Q)
        if out_d > depth:
            out_d = depth
        area = bf * barrels * width * out_d + 0.5 * (z1 + z2) * out_d ** 2
        perim = 2.0 * bf * barrels * width + (z1 + z2) * out_d + (out_d ** 2 + (z1 * out_d) ** 2) ** 0.5 \
            + (out_d ** 2 + (z2 * out_d) ** 2) ** 0.5
        rh = area / perim
        vel = math.sqrt(delta / ((o["sum_loss"] / 2 / G) + (o["manning"] ** 2 * o["length"]) / rh ** 1.33333))
        Qtail = area * vel
        if delta < drive:
            if tail_depth > depth:
                out_d = depth
            else:
                Q = min(Q, Qtail)
                out_d = newton(Q)
                if out_d > depth:
                    out_d = depth
            area = bf * barrels * width * out_d + 0.5 * (z1 + z2) * out_d ** 2
            perim = bf * barrels * width + (out_d ** 2 + (z1 * out_d) ** 2) ** 0.5 + (out_d ** 2 + (z2 * out_d) ** 2) ** 0.5
            rh = area / perim
            vel = math.sqrt(delta / ((o["sum_loss"] / 2 / G) + (o["manning"] ** 2 * o["length"]) / rh ** 1.33333))
            Q = min(Q, area * vel)
        return Q, Q / (area + VP / area), out_d, area

    @staticmethod
    def _boyd_pipe_rating(o, drive, delta, tail_depth):
        """boyd_pipe_function (boyd_pipe_operator.py:199-372): circular barrel; the energy-loss cap
        is applied in every case"""
        G, VP = 9.8, 1.0e-6
        diameter, barrels, blockage = o["diameter"], o["barrels"], o["blockage"]
        if blockage >= 1.0:
            return 0.0, 0.0, 0.0, 0.00001
        if blockage > 0.9:
            bf = 3.333 - 3.333 * blockage
        else:
            bf = 1.0 - 0.4012316798 * blockage - 0.3768350138 * (blockage ** 2)
        Qu = barrels * (0.421 * G ** 0.5 * ((bf * diameter) ** 0.87) * drive ** 1.63)
        Qs = barrels * (0.530 * G ** 0.5 * ((bf * diameter) ** 1.87) * drive ** 0.63)
        Q = min(Qu, Qs)
        dc1 = (bf * diameter) / 1.26 * (Q / G ** 0.5 * ((bf * diameter) ** 2.5)) ** (1 / 3.75)
        dc2 = (bf * diameter) / 0.95 * (Q / G ** 0.5 * (bf * diameter) ** 2.5) ** (1 / 1.95)
        out_d = dc2 if dc1 / (bf * diameter) > 0.85 else dc1
        if out_d >= (bf * diameter):
            out_d = bf * diameter
            area = barrels * (bf * diameter / 2) ** 2 * math.pi
            perim = barrels * bf * diameter * math.pi
        else:
            alpha = math.acos(1 - 2 * out_d / (bf * diameter)) * 2
            area = barrels * (bf * diameter) ** 2 / 8 * (alpha - math.sin(alpha))
            perim = barrels * (alpha * bf * diameter / 2.0)
        if delta < drive:
            if tail_depth > bf * diameter:
                out_d = bf * diameter
                area = barrels * (bf * diameter / 2) ** 2 * math.pi
                perim = barrels * bf * diameter * math.pi
            else:
                out_d = dc2 if dc1 / (bf * diameter) > 0.85 else dc1
                if out_d > bf * diameter:
                    out_d = bf * diameter
                    area = barrels * (bf * diameter / 2) ** 2 * math.pi
                    perim = barrels * bf * diameter * math.pi
                else:
                    alpha = math.acos(1 - 2 * out_d / (bf * diameter)) * 2
                    area = barrels * (bf * diameter) ** 2 / 8 * (alpha - math.sin(alpha))
                    perim = barrels * alpha * bf * diameter / 2.0
        rh = area / perim
        vel = math.sqrt(delta / ((o["sum_loss"] / 2 / G) + (o["manning"] ** 2 * o["length"]) / rh ** 1.33333))
        Q = min(Q, area * vel)
        return Q, Q / (area + VP / area), out_d, area

    # -- evolve ------------------------------------------------------------------
    def evolve(self, yieldstep=None, finaltime=None, duration=None, skip_initial_step=False):
        P = self.P
        epsilon = P["epsilon"]
        if self.evolved_called:
            skip_initial_step = True
        self.evolved_called = True
        if yieldstep is None:
            yieldstep = P["evolve_max_timestep"]
        yieldstep = float(yieldstep)
        self._order_ = P["default_order"]
        if finaltime is not None:
            self.relative_finaltime = float(finaltime) - self.starttime
        elif duration is not None:
            self.relative_finaltime = float(duration) + self.relative_time
        else:
            self.relative_finaltime = None
        self.relative_yieldtime = self.relative_time + yieldstep
        self.recorded_min_timestep = P["evolve_max_timestep"]
        self.recorded_max_timestep = P["evolve_min_timestep"]
        self.number_of_steps = 0
        self.number_of_first_order_steps = 0
        self.update_ghosts()
        if not skip_initial_step:
            self.distribute_to_vertices_and_edges()
            self.update_boundary()
            yield self.get_time()
        step = {"euler": self.evolve_one_euler_step, "rk2": self.evolve_one_rk2_step,
                "rk3": self.evolve_one_rk3_step}[self.timestepping_method]
        while True:
            initial_relative_time = self.relative_time
            step(yieldstep, finaltime)
            self.apply_fractional_steps()
            self.relative_time = initial_relative_time + self.timestep
            self.timestep_history.append(self.timestep)
            self.update_ghosts()
            self.number_of_steps += 1
            if self._order_ == 1:
                self.number_of_first_order_steps += 1
            if self.relative_finaltime is not None and \
                    self.relative_time >= self.relative_finaltime - epsilon:
                if self.relative_time > self.relative_finaltime:
                    raise RuntimeError("time overshot finaltime")
                self.relative_time = self.relative_finaltime
                self.distribute_to_vertices_and_edges()
                self.update_boundary()
                yield self.get_time()
                break
            if self.relative_time >= self.relative_yieldtime:
                self.distribute_to_vertices_and_edges()
                self.update_boundary()
                yield self.get_time()
                self.relative_yieldtime += yieldstep
                self.recorded_min_timestep = P["evolve_max_timestep"]
                self.recorded_max_timestep = P["evolve_min_timestep"]
                self.number_of_steps = 0
                self.number_of_first_order_steps = 0
                self.max_speed[:] = 0.0
