"""TEST INFRASTRUCTURE ONLY.

Snapshot a *reference* ``anuga.shallow_water.Domain`` (scratch build, see
pyref.py) into an oracle scenario dict.  Used in the build container to pin the
oracle against the unmodified Python reference and to generate the golden
fixtures under tests/golden/.
"""
import numpy as np


def boundary_spec_from_anuga(B):
    """Map a reference boundary object to an oracle boundary spec tuple."""
    if B is None:
        return None
    name = type(B).__name__
    if name == "Reflective_boundary":
        return ("reflective",)
    if name == "Dirichlet_boundary":
        return ("dirichlet", [float(v) for v in B.dirichlet_values[:3]])
    if name == "Transmissive_boundary":
        return ("transmissive",)
    if name == "Time_boundary":
        return ("time", B.function)
    if name == "Transmissive_n_momentum_zero_t_momentum_set_stage_boundary":
        return ("transmissive_n_zero_t_set_stage", B.function)
    if name == "Transmissive_momentum_set_stage_boundary":
        return ("transmissive_momentum_set_stage", B.function)
    if name == "Transmissive_stage_zero_momentum_boundary":
        return ("transmissive_stage_zero_momentum",)
    if name == "Time_stage_zero_momentum_boundary":
        return ("time_stage_zero_momentum", B.f)
    raise NotImplementedError(name)


def scenario_from_anuga_domain(domain, operators=()):
    q = domain.quantities
    sc = {}
    mesh = domain.mesh
    for name in ("neighbours", "neighbour_edges", "surrogate_neighbours", "number_of_boundaries",
                 "normals", "edgelengths", "radii", "areas", "centroid_coordinates",
                 "vertex_coordinates", "boundary_cells", "boundary_edges"):
        sc[name] = np.array(getattr(mesh, name), copy=True)
    sc["edge_coordinates"] = np.array(mesh.edge_midpoint_coordinates, copy=True)
    sc["tri_full_flag"] = np.array(domain.tri_full_flag, copy=True)
    sc["stage_centroid_values"] = q["stage"].centroid_values.copy()
    sc["xmom_centroid_values"] = q["xmomentum"].centroid_values.copy()
    sc["ymom_centroid_values"] = q["ymomentum"].centroid_values.copy()
    sc["bed_centroid_values"] = q["elevation"].centroid_values.copy()
    sc["friction_centroid_values"] = q["friction"].centroid_values.copy()
    sc["bed_vertex_values"] = q["elevation"].vertex_values.copy()
    sc["stage_vertex_values"] = q["stage"].vertex_values.copy()
    sc["params"] = dict(
        g=domain.g, epsilon=domain.epsilon, H0=domain.H0,
        minimum_allowed_height=domain.minimum_allowed_height,
        maximum_allowed_speed=domain.maximum_allowed_speed,
        evolve_max_timestep=domain.evolve_max_timestep,
        evolve_min_timestep=domain.evolve_min_timestep,
        max_smallsteps=domain.max_smallsteps, CFL=domain.CFL,
        timestepping_method=domain.get_timestepping_method(),
        beta_w=domain.beta_w, beta_w_dry=domain.beta_w_dry,
        beta_uh=domain.beta_uh, beta_uh_dry=domain.beta_uh_dry,
        beta_vh=domain.beta_vh, beta_vh_dry=domain.beta_vh_dry,
        extrapolate_velocity_second_order=int(domain.extrapolate_velocity_second_order),
        low_froude=int(domain.low_froude), optimise_dry_cells=int(domain.optimise_dry_cells),
        sloped_mannings=bool(domain.use_sloped_mannings),
        fixed_flux_timestep=getattr(domain, "fixed_flux_timestep", None),
        ghost_layer_width=domain.ghost_layer_width,
        centroid_transmissive_bc=bool(domain.centroid_transmissive_bc),
        default_order=domain.default_order,
    )
    sc["boundary_map"] = {tag: boundary_spec_from_anuga(B) for tag, B in domain.boundary_map.items()}
    sc["tag_boundary_cells"] = {t: np.array(v, dtype=np.int64) for t, v in domain.tag_boundary_cells.items()}
    sc["operators"] = list(operators)
    return sc
