"""The Merimbula lake model (reference: examples/parallel/run_parallel_merimbula.py), written once against
the shared API: used by the golden generator (reference), the oracle test and the GPU test."""
import numpy as np

TSH = "/root/reference/examples/parallel/data/merimbula_10785_1.tsh"
EVOLVE = dict(yieldstep=5.0, finaltime=15.0)


def build(A, d):
    """d: a domain on the Merimbula mesh with elevation loaded"""
    d.set_flow_algorithm("DE1")
    d.set_store(False)
    x0, x1 = 756000.0, 756500.0
    d.set_quantity("stage", lambda x, y: 1.0 * ((x > x0) & (x < x1)), location="centroids")
    d.set_quantity("friction", 0.02)
    Br = A.Reflective_boundary(d)
    Bts = A.Transmissive_n_momentum_zero_t_momentum_set_stage_boundary(d, lambda t: 10 * np.sin(t / 60))
    d.set_boundary({"exterior": Br, "open": Bts})
    return d


def from_fixture(A, g):
    """the same domain rebuilt from the mesh arrays stored in the golden fixture (no reference tree needed)"""
    boundary = {(int(k[0]), int(k[1])): str(t) for k, t in zip(g["boundary_keys"], g["boundary_tags"])}
    d = A.Domain(g["nodes"], g["triangles"], boundary)
    d.set_quantity("elevation", g["elevation_vertex"], location="vertices")
    return build(A, d)
