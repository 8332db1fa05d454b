"""Build kernel variants of libswk.so into variants/ (git-ignored; they travel to the GPU box).
usage: python profiles/build_variants.py name="-DFLAG=1 -DOTHER=2" ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anuga_core_b200 import build as b

os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
procs = []
for arg in sys.argv[1:]:
    name, flags = arg.split("=", 1)
    out = os.path.join(ROOT, "variants", "libswk_%s.so" % name)
    cmd = ["nvcc"] + b.NVCC_FLAGS + flags.split() + ["-o", out, b.SRC]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    outp = p.communicate()[0]
    print(name, "ok" if p.returncode == 0 else "FAILED\n" + outp)
