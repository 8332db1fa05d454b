"""Shared test inputs: the same seeded fields go to the CUDA path (through the product's
Domain API) and to the CPU oracle (as a plain scenario dict).  Fields follow SURVEY.md 8(d)."""
import numpy as np

import anuga_core_b200 as ab


def domain_to_scenario(domain):
    """Plain-array snapshot of a product Domain for oracle.driver.OracleDomain."""
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


def dam_break(n=20, alg="DE1", friction=0.03, boundary="reflective", **domain_kw):
    """Config 1 of SURVEY.md 8(d): sloping bed, partly shallow right half."""
    d = ab.rectangular_cross_domain(n, n, len1=float(n), len2=float(n), **domain_kw)
    d.set_flow_algorithm(alg)
    d.set_store(False)
    d.set_quantity("elevation", lambda x, y: -x / (n / 2.0))
    d.set_quantity("stage", lambda x, y: np.where(x < n / 2.0, 1.0, 0.2), location="centroids")
    d.set_quantity("friction", friction)
    _bind(d, boundary)
    return d


def wet_dry_beach(n=20, alg="DE1", **domain_kw):
    """Dam break running up a dry beach: exercises protect, dry-cell zeroing, fix-negative."""
    d = ab.rectangular_cross_domain(n, n, len1=float(n), len2=float(n), **domain_kw)
    d.set_flow_algorithm(alg)
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: 0.5 * (x - 0.4 * L) / (0.6 * L) * (x > 0.4 * L) + 0.05 * np.sin(y))
    d.set_quantity("stage", lambda x, y: np.where(x < 0.25 * L, 0.8, -1.0), location="centroids")
    d.set_quantity("friction", 0.03)
    _bind(d, "reflective")
    return d


def smooth_wet(n=20, alg="DE1", rain=None, **domain_kw):
    """Config 3: everywhere wet, smooth (the roofline-sweep fields)."""
    d = ab.rectangular_cross_domain(n, n, len1=float(n), len2=float(n), **domain_kw)
    d.set_flow_algorithm(alg)
    d.set_store(False)
    L = float(n)
    elev = lambda x, y: 0.01 * np.sin(2 * np.pi * x / 200.0) * np.cos(2 * np.pi * y / 200.0)
    d.set_quantity("elevation", elev)
    d.set_quantity("stage", lambda x, y: elev(x, y) + 0.5 + 0.1 * np.exp(-((x - L / 2) ** 2 + (y - L / 2) ** 2) / (0.1 * L) ** 2),
                   location="centroids")
    d.set_quantity("friction", 0.03)
    _bind(d, "reflective")
    if rain is not None:
        ab.Rate_operator(d, rate=rain)
    return d


def tsunami(n=24, alg="DE1", left="dirichlet", **domain_kw):
    """Config 2 style: sloping beach with an island, Dirichlet/set-stage inflow on the left,
    transmissive on the right, Manning 0.025."""
    d = ab.rectangular_cross_domain(n, n, len1=float(n), len2=float(n), **domain_kw)
    d.set_flow_algorithm(alg)
    d.set_store(False)
    L = float(n)
    d.set_quantity("elevation", lambda x, y: -(10 - 9.9 * x / L) + 0.5 * np.exp(-((x - 0.7 * L) ** 2 + (y - 0.5 * L) ** 2) / (0.05 * L) ** 2))
    d.set_quantity("stage", 0.0)
    d.set_quantity("friction", 0.025)
    Br = ab.Reflective_boundary(d)
    if left == "dirichlet":
        Bl = ab.Dirichlet_boundary([0.3, 0.0, 0.0])
    elif left == "set_stage":
        Bl = ab.Transmissive_n_momentum_zero_t_momentum_set_stage_boundary(d, lambda t: 0.5 * np.sin(2 * np.pi * t / 60.0))
    else:
        raise ValueError(left)
    d.set_boundary({"left": Bl, "right": ab.Transmissive_boundary(d), "top": Br, "bottom": Br})
    return d


def _bind(d, boundary):
    if boundary == "reflective":
        B = ab.Reflective_boundary(d)
        d.set_boundary({t: B for t in d.get_boundary_tags()})
    else:
        raise ValueError(boundary)


def rel_err(a, b):
    """|a-b| <= tol * max(|a|,|b|,floor), floor = 1e-12 * max|field| (SURVEY.md 8(d) parity gates)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    floor = 1e-12 * scale
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    return float(np.max(np.abs(a - b) / denom)) if a.size else 0.0


def scaled_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / scale
