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
    if domain.processor in domain.full_send_dict and domain.processor in domain.ghost_recv_dict:
        sc["ghost_copy"] = (np.asarray(domain.full_send_dict[domain.processor][0], dtype=np.int64),
                            np.asarray(domain.ghost_recv_dict[domain.processor][0], dtype=np.int64))
    for k in ("edge_flux_type", "edge_river_wall_counter", "riverwall_elevation", "riverwall_rowIndex",
              "riverwall_hydraulic_properties", "ncol_riverwall_hydraulic_properties"):
        if hasattr(domain, k):
            sc[k] = getattr(domain, k)
    return sc
