#!/usr/bin/env python
"""Static instruction mix of an address range of one kernel:  python tools/sass_mix.py <obj> <mangled substring> <lo> <hi> [div]
(cuobjdump -sass; opcode counts grouped by issue pipe; div = steps the range covers)."""
import re
import subprocess
import sys

obj, func, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
div = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
on = False
ops = {}
for line in txt.splitlines():
    if "Function :" in line:
        on = func in line
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if on and m and lo <= int(m.group(1), 16) <= hi:
        op = m.group(2).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
FMA = {"IMAD", "FMUL", "FADD", "FFMA", "HFMA2"}
ALU = {"LOP3", "SHF", "PRMT", "IADD3", "ISETP", "SEL", "LEA", "VIADD", "VIMNMX", "IADD", "MOV", "FLO", "POPC", "FSETP", "PLOP3", "BMSK", "SGXT", "IABS", "VABSDIFF", "FMNMX"}
XU = {"MUFU", "I2FP", "F2I", "I2F", "F2F"}
LSU = {"LDS", "STS", "LDG", "STG", "LDGSTS", "LDGDEPBAR", "DEPBAR", "ATOMS", "RED", "LDC", "LDSM"}
tot = sum(ops.values())
grp = {"fma": 0, "alu": 0, "xu": 0, "lsu": 0, "other": 0}
for op, c in ops.items():
    g = "fma" if op in FMA else "alu" if op in ALU else "xu" if op in XU else "lsu" if op in LSU else "other"
    grp[g] += c
print(f"{tot} instructions ({tot / div:.1f} per step): " + ", ".join(f"{g} {c / div:.1f}" for g, c in grp.items()))
print(" ".join(f"{op}:{c / div:.1f}" for op, c in sorted(ops.items(), key=lambda kv: -kv[1])))
