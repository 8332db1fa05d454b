"""Small driver for ncu: builds the roofline-sweep domain and runs a few DE1 steps.
usage: python profiles/run_profile.py [cells_per_side=2000] [steps=3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anuga_core_b200 import workloads

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = workloads.roofline_sweep_domain(size, size, alg="DE1", rain=1.0e-4)
it = d.evolve(yieldstep=1.0e9, finaltime=None)
next(it)
ms = d._dev.run_steps(steps)
print("N=%d steps=%d ms/step=%.3f" % (d.number_of_triangles, steps, ms / steps))
