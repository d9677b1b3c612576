#!/usr/bin/env python3
"""tools/sass_report.py [L C] -> profiles/r2_sass_fast_dp_<L>_<C>.txt

Evidence for the DPX claim: dumps the SASS of one instantiation of the packed DP kernel from the
shipped rabbitsalign_b200/librsa_ext.so (cuobjdump), finds its steady-state row loops (the two back-edges with the
SHFL.UP pair: the HASN=false and HASN=true variants), and writes the opcode histogram of the whole kernel, the
histogram and per-column averages of the shorter loop (reads without N), and that loop's body."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
C = int(sys.argv[2]) if len(sys.argv) > 2 else 19
lib = os.path.join(ROOT, "rabbitsalign_b200", "librsa_ext.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
# split into functions
funcs = re.split(r"\n\s+Function : ", sass)
want = f"fast_dp_kernelILi{L}ELi{C}E"
body = next((f for f in funcs if want in f.split("\n", 1)[0]), None)
if body is None:
    sys.exit(f"no kernel matching {want}")
name = body.split("\n", 1)[0].strip()
ins = []  # (address, text)
for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", body):
    ins.append((int(m.group(1), 16), m.group(2).strip()))


def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0]


def hist(seq):
    return collections.Counter(opcode(t).split(".")[0] for _, t in seq)


# loops: backward branches
loops = []
for a, t in ins:
    m = re.match(r"(?:@!?U?P\d+\s+)?BRA\s+(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        lo = int(m.group(1), 16)
        seq = [(x, y) for x, y in ins if lo <= x <= a]
        n_dpx = sum(1 for _, y in seq if "VIADDMNMX" in y)
        if sum(1 for _, y in seq if y.startswith("SHFL.UP")) >= 2 and 7 * C <= n_dpx < 8 * C:  # one row: 3 fused add+max and 4 clamp facts per column
            loops.append((lo, a, seq))
out = [f"kernel {name}", f"{len(ins)} instructions; row loops found: " + ", ".join(f"[{lo:#x}, {hi:#x}] {len(s)} instr" for lo, hi, s in loops), ""]
regs = re.search(r"REG:(\d+)", subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout.split(want, 1)[-1][:400] if want else "")
if regs:
    out.append(f"registers per thread: {regs.group(1)}")
out.append("opcode histogram, whole kernel:")
out += [f"  {n:5d} {op}" for op, n in hist(ins).most_common()]
if loops:
    lo, hi, seq = min(loops, key=lambda x: len(x[2]))
    h = hist(seq)
    alu = sum(h[o] for o in ("LOP3", "IADD3", "VIADDMNMX", "VIMNMX3", "VIMNMX", "PRMT", "SHF", "VIADD", "LEA", "SEL", "ISETP", "PLOP3", "HSET2"))
    fma = sum(h[o] for o in ("IMAD", "HFMA2", "FFMA"))
    hf = h["HFMA2"]
    out += ["", f"steady-state row loop without N in the reads [{lo:#x}, {hi:#x}]: {len(seq)} instructions per target row of {C} column "
            f"pairs = {len(seq) / C:.2f} per packed cell pair ({alu / C:.2f} on the ALU-pipe opcode list, {fma / C:.2f} on the FMA pipe: {(fma - hf) / C:.2f} IMAD + {hf / C:.2f} HFMA2)"]
    out += [f"  {n:5d} {op}  ({n / C:.2f}/column)" for op, n in h.most_common()]
    out += ["", "loop body:"] + [f"  {a:#06x}  {t}" for a, t in seq]
dst = os.path.join(ROOT, "profiles", f"r2_sass_fast_dp_{L}_{C}.txt")
open(dst, "w").write("\n".join(out) + "\n")
print(dst, len(out), "lines")
