#!/usr/bin/env python
"""Code-level evidence for DESIGN.md section 3/4 (no GPU needed): rebuilds libswk.so with `-Xptxas -v` and
disassembles it, then writes
  profiles/r2/ptxas_table.txt   registers / spills / shared memory of every kernel
  profiles/r2/sass_excerpt.txt  per hot kernel: instruction-class counts and every global-memory instruction
                                (the 256-bit record accesses LDG.E.256 / STG.E.256, their cache-policy
                                modifiers, and the bulk L2 prefetch UBLKPF.L2)
usage: python profiles/sass_evidence.py"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anuga_core_b200 import build as b

OUT = os.path.join(ROOT, "profiles", "r2")
os.makedirs(OUT, exist_ok=True)
HOT = ["k_extrapolate", "k_flux<false>", "k_update(", "k_flux_update<false>", "k_flux_update<true>", "k_boundary_values",
       "k_update_timestep", "k_finish_step"]

obj = "/tmp/swk_evidence.o"
res = subprocess.run(["nvcc"] + [f for f in b.NVCC_FLAGS if f not in ("-shared", "-ldl")] +
                     ["-Xptxas", "-v", "-c", "-o", obj, b.SRC], capture_output=True, text=True)
if res.returncode != 0:
    sys.exit(res.stderr)
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s
rows = []
cur = None
for line in res.stderr.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = dict(mangled=m.group(1), name=demangle(m.group(1)), regs=None, spill_st=0, spill_ld=0, stack=0, smem=0)
        rows.append(cur)
        props_of = None
        continue
    if cur is None:
        continue
    m = re.search(r"Function properties for (\S+)", line)
    if m:
        props_of = m.group(1)          # (called device functions report their own frames after the kernel's)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and props_of == cur["mangled"]:
        cur["stack"], cur["spill_st"], cur["spill_ld"] = map(int, m.groups())
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
        m2 = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = int(m2.group(1)) if m2 else 0
with open(os.path.join(OUT, "ptxas_table.txt"), "w") as fh:
    fh.write("nvcc %s\n\n" % " ".join(f for f in b.NVCC_FLAGS if f not in ("-shared", "-ldl")))
    fh.write("%-92s %5s %9s %9s %6s %6s\n" % ("kernel", "regs", "spill st", "spill ld", "stack", "smem"))
    for r in sorted(rows, key=lambda r: r["name"]):
        fh.write("%-92s %5s %9d %9d %6d %6d\n" % (r["name"][:92], r["regs"], r["spill_st"], r["spill_ld"], r["stack"], r["smem"]))

sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)[1:]
with open(os.path.join(OUT, "sass_excerpt.txt"), "w") as fh:
    fh.write("cuobjdump -sass of the object built with the flags in ptxas_table.txt (sm_100a)\n")
    for blk in blocks:
        mangled = blk.split("\n", 1)[0].strip()
        name = demangle(mangled)
        if not any(h in name for h in HOT):
            continue
        ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", blk)
        classes = collections.Counter(i.split(".")[0] for i in ins)
        mem = collections.Counter(i for i in ins if i.split(".")[0] in
                                  ("LDG", "STG", "LDL", "STL", "UBLKPF", "CCTL", "ATOMG", "RED", "ATOM", "LDS", "STS"))
        fh.write("\n== %s\n   %d instructions; FP64: DFMA %d DMUL %d DADD %d DSETP %d MUFU %d\n" % (
            name, len(ins), classes["DFMA"], classes["DMUL"], classes["DADD"], classes["DSETP"], classes["MUFU"]))
        for k, v in sorted(mem.items()):
            fh.write("   %4d x %s\n" % (v, k))
print("written", OUT)
