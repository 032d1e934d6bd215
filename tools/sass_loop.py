#!/usr/bin/env python3
"""Write the hot loop of a dumped kernel (tools/dump_kernel.py <config> first) to profiles/<out>: the smallest loop that holds
at least `min_fp64` FP64 instructions (default 60), or with `mma` the smallest loop that holds DMMA instructions.
usage: sass_loop.py <config> <out file> "<title>" [min_fp64 | mma]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
sel = sys.argv[4] if len(sys.argv) > 4 else "60"
lines = open(os.path.join(ROOT, "sde-sim-rs_b200", "build", cfg + ".sass")).read().splitlines()
ins = []
for l in lines:
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            loops.append((addr[tgt], i))


def op(t):
    return re.sub(r'^@!?U?P\d+\s+', '', t).split()[0]


def count(l, names):
    return sum(1 for _, t in ins[l[0]:l[1] + 1] if op(t).split('.')[0] in names)


if sel == "mma":
    cand = [l for l in loops if count(l, ("DMMA",)) > 0]
else:
    cand = [l for l in loops if count(l, ("DFMA", "DMUL", "DADD")) >= int(sel)]
lo, hi = min(cand, key=lambda l: l[1] - l[0])
mix = {}
for _, t in ins[lo:hi + 1]:
    k = op(t).split('.')[0]
    mix[k] = mix.get(k, 0) + 1
top = ", ".join(f"{v} {k}" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:12])
with open(os.path.join(ROOT, "profiles", out), "w") as f:
    f.write(f"// {title}\n// tools/dump_kernel.py {cfg} (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a); {hi - lo + 1} instructions: {top}\n")
    for a, t in ins[lo:hi + 1]:
        f.write("/*%04x*/  %s ;\n" % (a, t))
print(out, hi - lo + 1, top)
