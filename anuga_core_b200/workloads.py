"""Synthetic inputs of the benchmark configurations (BASELINE.json
`configs`).  All are rectangular_cross_domain(m, n, len1=m*dx, len2=n*dx), dx = 1."""
import numpy as np

from .domain import rectangular_cross_domain
from .boundaries import Reflective_boundary
from .operators import Rate_operator


def sweep_elevation(x, y):
    return 0.01 * np.sin(2 * np.pi * x / 200.0) * np.cos(2 * np.pi * y / 200.0)


def sweep_stage(Lx, Ly):
    L = Lx
    return lambda x, y: sweep_elevation(x, y) + 0.5 + 0.1 * np.exp(-((x - Lx / 2) ** 2 + (y - Ly / 2) ** 2) / (0.1 * L) ** 2)


def roofline_sweep_domain(m, n=None, alg="DE1", rain=1.0e-4, **domain_kw):
    """configs[2]: everywhere-wet smooth fields, DE1, Reflective boundaries, scalar rain
    Rate_operator, Manning 0.03.  m = n = 2000 gives the 16M-triangle case."""
    n = m if n is None else n
    d = rectangular_cross_domain(m, n, len1=float(m), len2=float(n), **domain_kw)
    d.set_flow_algorithm(alg)
    d.set_store(False)
    d.set_quantity("elevation", sweep_elevation)
    d.set_quantity("stage", sweep_stage(float(m), float(n)), location="centroids")
    d.set_quantity("friction", 0.03)
    B = Reflective_boundary(d)
    d.set_boundary({t: B for t in d.get_boundary_tags()})
    if rain is not None:
        Rate_operator(d, rate=rain)
    return d


def dam_break_domain(m=100, n=None, alg="DE0", **domain_kw):
    """configs[0]: 100x100 (40k triangles) dam break over a sloping bed, Reflective boundaries."""
    n = m if n is None else n
    d = rectangular_cross_domain(m, n, len1=float(m), len2=float(n), **domain_kw)
    d.set_flow_algorithm(alg)
    d.set_store(False)
    d.set_quantity("elevation", lambda x, y: -x / (m / 2.0))
    d.set_quantity("stage", lambda x, y: np.where(x < m / 2.0, 1.0, 0.2), location="centroids")
    d.set_quantity("friction", 0.03)
    B = Reflective_boundary(d)
    d.set_boundary({t: B for t in d.get_boundary_tags()})
    return d


def tsunami_fields(d, L):
    """configs[1]: sloping beach with an island, still water, Manning 0.025;
    left: set-stage wave 0.5 sin(2 pi t / 60) with transmissive normal momentum, right: Transmissive,
    top / bottom: Reflective."""
    from .boundaries import Transmissive_boundary, Transmissive_n_momentum_zero_t_momentum_set_stage_boundary
    d.set_quantity("elevation", lambda x, y: -(10 - 9.9 * x / L)
                   + 0.5 * np.exp(-((x - 0.7 * L) ** 2 + (y - 0.5 * L) ** 2) / (0.05 * L) ** 2))
    d.set_quantity("stage", 0.0)
    d.set_quantity("friction", 0.025)
    Br = Reflective_boundary(d)
    Bl = Transmissive_n_momentum_zero_t_momentum_set_stage_boundary(d, tsunami_wave)
    Bt = Transmissive_boundary(d)
    bmap = {"left": Bl, "right": Bt, "top": Br, "bottom": Br}
    d.set_boundary({t: bmap.get(t) for t in d.get_boundary_tags()})


def tsunami_wave(t):
    import math
    return 0.5 * math.sin(2 * math.pi * t / 60.0)


def tsunami_domain(m, n=None, alg="DE1", **domain_kw):
    n = m if n is None else n
    d = rectangular_cross_domain(m, n, len1=float(m), len2=float(n), **domain_kw)
    d.set_flow_algorithm(alg)
    d.set_store(False)
    tsunami_fields(d, float(m))
    return d


def structures_fields(d, m, n, with_structures=True):
    """configs[4]: an embankment across the domain at x = L/2, water behind it,
    an Inlet_operator (Q = 100 m^3/s over a line) upstream and one Boyd box culvert through the embankment."""
    from .structures import Region, Inlet_operator, Boyd_box_operator
    L = float(m)
    xe = 0.5 * L
    d.set_quantity("elevation", lambda x, y: 2.0 * np.exp(-((x - xe) / 6.0) ** 2) + 0.001 * (L - x) / L)
    d.set_quantity("stage", lambda x, y: np.where(x < xe, 1.2, 0.4), location="centroids")
    d.set_quantity("friction", 0.03)
    Br = Reflective_boundary(d)
    d.set_boundary({t: (None if t == "ghost" else Br) for t in d.get_boundary_tags()})
    if with_structures:
        y0 = 0.5 * n
        Inlet_operator(d, Region(d, line=[[0.1 * L + 0.3, y0 - 20.2], [0.1 * L + 0.3, y0 + 20.3]]), Q=100.0)
        Boyd_box_operator(d, losses=1.5, width=3.0, height=1.5,
                          end_points=[[xe - 15.1, y0 + 0.3], [xe + 15.1, y0 + 0.3]],
                          apron=2.55, enquiry_gap=1.4, manning=0.013)


def structures_domain(m, n, rank=0, nranks=1, with_structures=True, comm=None, **domain_kw):
    """configs[4]: rectangular_cross m x n (4000 x 2000 = 32M triangles), DE1 with rk3 timestepping
    (set_flow_algorithm('DE2')).  nranks > 1: the rank's strip of the SAME global mesh (the 32M triangles
    are shared out, configs[4] names 8 GPUs); the structures are created on the sub-domain with their
    global geometry, as the reference's parallel scripts do after distribute()."""
    if nranks == 1:
        d = rectangular_cross_domain(m, n, len1=float(m), len2=float(n), **domain_kw)
    else:
        from . import parallel
        d = parallel.strip_partitioned_mesh_domain(m, n, rank, nranks, **domain_kw)
        if comm is not None:
            d.attach_communicator(comm)
    d.set_flow_algorithm("DE2")
    d.set_store(False)
    structures_fields(d, m, n, with_structures)
    return d


def domain_to_scenario(domain):
    """Plain-array snapshot of a Domain (inputs of the CPU oracle / reference arm).
    Only arrays and scalars: no dependency of the checker on this package."""
    q = domain.quantities
    m = domain.mesh
    sc = {}
    for name in ("neighbours", "neighbour_edges", "surrogate_neighbours", "number_of_boundaries",
                 "normals", "edgelengths", "radii", "areas", "centroid_coordinates",
                 "vertex_coordinates", "boundary_cells", "boundary_edges"):
        sc[name] = np.array(getattr(m, name), copy=True)
    sc["edge_coordinates"] = np.array(m.edge_midpoint_coordinates, copy=True)
    sc["tri_full_flag"] = np.array(domain.tri_full_flag, copy=True)
    sc["stage_centroid_values"] = q["stage"].centroid_values.copy()
    sc["xmom_centroid_values"] = q["xmomentum"].centroid_values.copy()
    sc["ymom_centroid_values"] = q["ymomentum"].centroid_values.copy()
    sc["bed_centroid_values"] = q["elevation"].centroid_values.copy()
    sc["friction_centroid_values"] = q["friction"].centroid_values.copy()
    sc["bed_vertex_values"] = q["elevation"].vertex_values.copy()
    sc["params"] = dict(
        g=domain.g, epsilon=domain.epsilon, H0=domain.H0,
        minimum_allowed_height=domain.minimum_allowed_height,
        maximum_allowed_speed=domain.maximum_allowed_speed,
        evolve_max_timestep=domain.evolve_max_timestep, evolve_min_timestep=domain.evolve_min_timestep,
        max_smallsteps=domain.max_smallsteps, CFL=domain.CFL,
        timestepping_method=domain.timestepping_method,
        beta_w=domain.beta_w, beta_w_dry=domain.beta_w_dry, beta_uh=domain.beta_uh,
        beta_uh_dry=domain.beta_uh_dry, beta_vh=domain.beta_vh, beta_vh_dry=domain.beta_vh_dry,
        extrapolate_velocity_second_order=int(domain.extrapolate_velocity_second_order),
        low_froude=int(domain.low_froude), sloped_mannings=bool(domain.use_sloped_mannings),
        fixed_flux_timestep=domain.fixed_flux_timestep, ghost_layer_width=domain.ghost_layer_width,
        centroid_transmissive_bc=bool(domain.centroid_transmissive_bc), default_order=domain.default_order,
    )
    if domain.boundary_map is not None:
        sc["boundary_map"] = {t: (None if B is None else B.oracle_spec()) for t, B in domain.boundary_map.items()}
    sc["tag_boundary_cells"] = {t: np.array(v, dtype=np.int64) for t, v in domain.tag_boundary_cells.items()}
    sc["operators"] = [op.oracle_spec() for op in domain.fractional_step_operators]
    sc["forcing"] = [f.oracle_spec() for f in getattr(domain, "forcing_terms", []) if hasattr(f, "oracle_spec")]
    if domain.processor in domain.full_send_dict and domain.processor in domain.ghost_recv_dict:
        sc["ghost_copy"] = (np.asarray(domain.full_send_dict[domain.processor][0], dtype=np.int64),
                            np.asarray(domain.ghost_recv_dict[domain.processor][0], dtype=np.int64))
    for k in ("edge_flux_type", "edge_river_wall_counter", "riverwall_elevation", "riverwall_rowIndex",
              "riverwall_hydraulic_properties", "ncol_riverwall_hydraulic_properties"):
        if hasattr(domain, k):
            sc[k] = getattr(domain, k)
    return sc
