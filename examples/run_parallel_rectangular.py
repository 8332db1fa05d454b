"""A parallel run written the way the reference's example scripts are (examples/parallel/ in
anuga_core: rank 0 builds the sequential domain, anuga.distribute hands out the sub-domains,
boundaries and operators are set on the distributed domain, every rank writes its SWW file and
sww_merge combines them) - with this package imported under the reference's name.

    python examples/run_parallel_rectangular.py                               # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \\
           --master-addr 127.0.0.1 --master-port 29555 examples/run_parallel_rectangular.py -sn 400
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import anuga_core_b200 as anuga
from anuga_core_b200 import Reflective_boundary, Set_stage, rectangular_cross_domain
from anuga_core_b200 import distribute, myid, numprocs, finalize, barrier

parser = argparse.ArgumentParser(description="rectangular dam break, one process per GPU")
parser.add_argument("-ft", "--finaltime", type=float, default=0.5)
parser.add_argument("-ys", "--yieldstep", type=float, default=0.1)
parser.add_argument("-sn", "--sqrtN", type=int, default=200, help="cells per side: 500 -> 1 000 000 triangles")
parser.add_argument("-gl", "--ghost_layer", type=int, default=2)
parser.add_argument("-o", "--outdir", default=".")
args = parser.parse_args()

t0 = time.time()
if myid == 0:
    domain = rectangular_cross_domain(args.sqrtN, args.sqrtN, len1=2.0, len2=2.0, origin=(-1.0, -1.0))
    domain.set_store(True)
    domain.set_quantity("elevation", lambda x, y: -1.0 - x)
    domain.set_quantity("stage", 1.0)
    domain.set_flow_algorithm("DE1")
    domain.set_name("sw_rectangle")
    domain.set_datadir(args.outdir)
    print("sequential domain: %d triangles in %.2f s" % (domain.number_of_triangles, time.time() - t0))
else:
    domain = None

domain = distribute(domain, parameters=dict(ghost_layer_width=args.ghost_layer))

R = Reflective_boundary(domain)
domain.set_boundary({"left": R, "right": R, "bottom": R, "top": R})
Set_stage(domain, center=(0.0, 0.0), radius=0.5, stage=2.0)()

barrier()
t0 = time.time()
steps = 0
for t in domain.evolve(yieldstep=args.yieldstep, finaltime=args.finaltime):
    steps += domain.number_of_steps
    if myid == 0:
        domain.print_timestepping_statistics()
evolve_time = time.time() - t0
volume = domain.get_water_volume()
if numprocs > 1:
    volume = domain._comm.allreduce_sum(volume)
if myid == 0:
    print("evolve: %.2f s, %d steps on %d GPU(s); water volume %.6f" % (evolve_time, steps, numprocs, volume))
domain.sww_merge(delete_old=True)
if myid == 0:
    print("output:", os.path.join(args.outdir, domain.get_global_name() + ".sww"))
finalize()
