#!/usr/bin/env python
"""Golden partition indexing from the reference's own pipeline (distribute_mesh.py) for a
GIVEN element partition: a fake ``pymetis.part_graph`` injects the same epart that the test
later gives to anuga_core_b200.parallel (pymetis is absent here and its output is not pinned
by the reference's tests anyway - SURVEY.md 8(c)).

Run in the build container:  python tests/golden/make_golden_partition.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

CASES = {
    # name: (m, n, nparts, ghost_layer_width, epart rule)
    "partition_strips_8x5_p3": (8, 5, 3, 2, "strips"),
    "partition_checker_6x6_p4": (6, 6, 4, 2, "checker"),
    "partition_strips_10x4_p2_w4": (10, 4, 2, 4, "strips"),
}


def make_epart(rule, cx, cy, m, n, nparts):
    if rule == "strips":
        return np.minimum((cx / m * nparts).astype(np.int64), nparts - 1)
    if rule == "checker":   # 2x2 blocks, deliberately not contiguous in the original numbering
        return (cx >= m / 2).astype(np.int64) * 2 + (cy >= n / 2).astype(np.int64)
    raise ValueError(rule)


def main():
    holder = {}
    anuga = pyref.import_anuga(epart_fn=lambda nparts, adj: holder["epart"])
    import anuga.parallel.distribute_mesh as dm
    import pymetis
    dm.part_graph = pymetis.part_graph
    dm.metis_version = "5_part_graph"
    from anuga.parallel.sequential_distribute import Sequential_distribute
    for name, (m, n, P, width, rule) in CASES.items():
        d = anuga.rectangular_cross_domain(m, n, len1=float(m), len2=float(n))
        d.set_store(False)
        c = d.centroid_coordinates
        epart = make_epart(rule, c[:, 0], c[:, 1], m, n, P)
        holder["epart"] = epart.tolist()
        sd = Sequential_distribute(d, parameters={"ghost_layer_width": width})
        sd.distribute(P)
        out = {"epart": epart, "m": m, "n": n, "nparts": P, "width": width}
        for p in range(P):
            (points, vertices, boundary, quantities, ghost_recv, full_send, tri_map, node_map,
             tri_l2g, node_l2g, glw) = sd.extract_submesh(p)[:11] if False else \
                dm.extract_submesh(sd.submesh, sd.triangles_per_proc, sd.p2s_map, p)
            pre = "r%d_" % p
            out[pre + "points"] = points
            out[pre + "triangles"] = vertices
            keys = sorted(boundary.keys())
            out[pre + "boundary_keys"] = np.array(keys, dtype=np.int64).reshape(-1, 2)
            out[pre + "boundary_tags"] = np.array([boundary[k] for k in keys])
            out[pre + "tri_l2s"] = np.asarray(tri_l2g)
            out[pre + "node_l2g"] = np.asarray(node_l2g)
            out[pre + "nfull"] = np.array([len(sd.submesh["full_triangles"][p])])
            for q, v in full_send.items():
                out[pre + "send_%d_local" % q] = np.asarray(v[0])
                out[pre + "send_%d_global" % q] = np.asarray(v[1])
            for q, v in ghost_recv.items():
                out[pre + "recv_%d_local" % q] = np.asarray(v[0])
                out[pre + "recv_%d_global" % q] = np.asarray(v[1])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "ranks", P, "files", len(out))


if __name__ == "__main__":
    main()
