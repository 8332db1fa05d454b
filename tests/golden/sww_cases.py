"""SWW output cases shared by the reference (make_golden_sww.py) and this repository's tests."""
import numpy as np

EVOLVE = dict(yieldstep=0.5, finaltime=2.0)


def static_domain(A, datadir, name):
    d = A.rectangular_cross_domain(5, 4, len1=5.0, len2=4.0)
    d.set_flow_algorithm("DE1")
    d.set_name(name)
    d.set_datadir(datadir)
    d.set_store(True)
    d.set_quantity("elevation", lambda x, y: 0.3 * np.sin(x) - 0.2 * y)
    d.set_quantity("friction", 0.025)
    d.set_quantity("stage", lambda x, y: np.maximum(0.3 * np.sin(x) - 0.2 * y, -0.3 + 0.05 * x))
    d.set_quantity("xmomentum", lambda x, y: 0.1 * x * y, location="centroids")
    d.set_quantity("ymomentum", lambda x, y: -0.05 * x)
    return d


def store_two_frames(d):
    """the storage calls Domain.evolve makes, without evolving (host arrays only)"""
    d.initialise_storage()
    d.store_timestep()
    q = d.quantities["stage"]
    q.vertex_values[:] = q.vertex_values + 0.07
    q.centroid_values[:] = q.centroid_values + 0.07
    d.relative_time = 0.5
    d.store_timestep()


def evolve_domain(A, cases, datadir, name):
    d = cases.beach_de1(A, n=10)
    d.set_store(True)
    d.set_name(name)
    d.set_datadir(datadir)
    return d


def distributed_static_files(ab, P, datadir, name, nparts=3):
    """per-rank SWW files (two frames, no evolve) of static_domain cut in `nparts` by an interleaved
    element partition; returns the global domain"""
    g = static_domain(ab, datadir, name)
    c = g.centroid_coordinates
    epart = ((np.floor(c[:, 0]) + 2 * np.floor(c[:, 1])) % nparts).astype(int)
    subs = P.distribute(g, nparts, epart=epart)
    for p in sorted(subs):
        store_two_frames(subs[p])
    return g, subs
