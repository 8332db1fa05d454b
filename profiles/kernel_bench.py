"""Per-kernel device times of the DE1 step for the library named by $SWK_LIB (experiments).
usage: SWK_LIB=path python profiles/kernel_bench.py [cells_per_side=1000] [steps=30]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anuga_core_b200 import workloads

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
d = workloads.roofline_sweep_domain(size, size, alg="DE1", rain=1.0e-4)
it = d.evolve(yieldstep=1.0e9, finaltime=None)
next(it)
dev = d._dev
dev.run_steps(5)
ms = dev.run_steps(steps, per_kernel=True)
kt = dev.kernel_timing()
N = d.number_of_triangles
print("%-40s N=%d  %.3f ms/step  %.3e tri-steps/s | " % (os.path.basename(os.environ.get("SWK_LIB", "libswk.so")), N, ms / steps, N * steps / ms * 1e3)
      + "  ".join("%s %.3f" % (k, v[0] / max(v[1], 1)) for k, v in kt.items()), flush=True)
