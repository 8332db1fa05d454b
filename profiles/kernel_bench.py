"""Per-kernel device times of the DE1 step for the library named by $SWK_LIB (experiments).
usage: SWK_LIB=path python profiles/kernel_bench.py [cells_per_side=1000] [steps=30]
Prints one line; `sha` is a digest of the conserved centroid arrays after the run, so that variants of the
kernels can be checked for bit-identical results against each other."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anuga_core_b200 import workloads

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
alg = sys.argv[3] if len(sys.argv) > 3 else "DE1"
d = workloads.roofline_sweep_domain(size, size, alg=alg, rain=1.0e-4)
it = d.evolve(yieldstep=1.0e9, finaltime=None)
next(it)
dev = d._dev
dev.run_steps(5)
if os.environ.get("KB_GRAPH_FIRST"):
    ms2 = dev.run_steps(steps, per_kernel=False)    # graph replay, no per-kernel events
    ms = dev.run_steps(steps, per_kernel=True)
    kt = dev.kernel_timing()
else:
    ms = dev.run_steps(steps, per_kernel=True)
    kt = dev.kernel_timing()
    ms2 = dev.run_steps(steps, per_kernel=False)    # graph replay, no per-kernel events
N = d.number_of_triangles
d._mark_device_newer()
d.sync_to_host()
h = hashlib.sha1()
for name in ("stage", "xmomentum", "ymomentum"):
    h.update(d.quantities[name].centroid_values.tobytes())
print("%-28s N=%d %.3f ms/step (graph %.3f) %.3e tri-steps/s | " % (
    os.path.basename(os.environ.get("SWK_LIB", "libswk.so")), N, ms / steps, ms2 / steps, N * steps / ms2 * 1e3)
      + "  ".join("%s %.3f" % (k, v[0] / max(v[1], 1)) for k, v in kt.items()) + " | sha " + h.hexdigest()[:12], flush=True)
