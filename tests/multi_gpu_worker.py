"""torchrun worker: evolve one golden case distributed over WORLD_SIZE GPUs and dump the full
cells of every rank.  usage: torchrun ... tests/multi_gpu_worker.py CASE OUTDIR [rule]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import anuga_core_b200 as ab  # noqa: E402
from anuga_core_b200 import parallel as P  # noqa: E402
import cases  # noqa: E402

case, outdir = sys.argv[1], sys.argv[2]
rule = sys.argv[3] if len(sys.argv) > 3 else "blocks"
comm = P.init_process_group()
rank, size = comm.rank, comm.size
builder, ev = cases.CASES[case]
g = builder(ab)
store = len(sys.argv) > 4 and sys.argv[4] == "store"
if store:
    g.set_store(True)
    g.set_name("multi_" + case)
    g.set_datadir(outdir)
N = g.number_of_triangles
if rule == "blocks":
    epart = (np.arange(N) * size) // N
else:   # interleaved quadrants: non-contiguous in the original numbering, several peers per rank
    c = g.centroid_coordinates
    L = c[:, 0].max()
    epart = ((c[:, 0] > L / 2).astype(int) + 2 * (c[:, 1] > L / 2).astype(int)) % size
d = P.distribute(g, size, epart=epart, ranks=[rank], domain_kw=dict(device=int(os.environ.get("LOCAL_RANK", "0"))))[rank]
d.attach_communicator(comm)
times, steps = [], 0
for t in d.evolve(**ev):
    times.append(t)
    steps += d.number_of_steps
if store:
    d.sww_merge(delete_old=True)
nf = d.number_of_full_triangles
q = d.quantities
np.savez(os.path.join(outdir, "rank%d.npz" % rank), ids=d.tri_l2s[:nf], gids=d.tri_l2g[:nf], stage=q["stage"].centroid_values[:nf],
         xmom=q["xmomentum"].centroid_values[:nf], ymom=q["ymomentum"].centroid_values[:nf],
         times=np.array(times), steps=np.array([steps]), dt=np.array([d.timestep]))
comm.barrier()
