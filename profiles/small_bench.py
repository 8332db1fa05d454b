"""config[0]-sized meshes: steps/s of the device time loop (launch-bound regime).
usage: python profiles/small_bench.py [cells=100] [alg=DE0] [steps=2000]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anuga_core_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
alg = sys.argv[2] if len(sys.argv) > 2 else "DE0"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
d = workloads.dam_break_domain(n, n, alg=alg)
it = d.evolve(yieldstep=1.0e9, finaltime=None)
next(it)
dev = d._dev
dev.run_steps(50)
ms = dev.run_steps(steps)
t0 = time.perf_counter()
r = dev.evolve(1e9, None, steps)
dev.synchronize()
wall = time.perf_counter() - t0
N = d.number_of_triangles
print("%s n=%d N=%d graph=%s: run_steps %.2f us/step (%.3e tri-steps/s) | swk_evolve %.2f us/step wall" % (
    alg, n, N, os.environ.get("SWK_NO_GRAPH", "0") != "1", ms / steps * 1e3, N * steps / ms * 1e3, wall / steps * 1e6))
