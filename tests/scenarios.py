"""Shared test inputs: the same seeded fields go to the CUDA path (through the product's
Domain API) and to the CPU oracle (as a plain scenario dict).  Fields follow SURVEY.md 8(d)."""
import numpy as np

import anuga_core_b200 as ab


from anuga_core_b200.workloads import domain_to_scenario  # noqa: E402,F401


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
