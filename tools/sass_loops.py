#!/usr/bin/env python
"""Loops of a kernel's SASS with an opcode histogram each (which pipe does the hot loop load?).
usage: sass_loops.py MANGLED_KERNEL_NAME [min_instr]   (reads mdz_b200/libmdzcuda.so via cuobjdump)"""
import collections
import re
import subprocess
import sys

name = sys.argv[1]
min_n = int(sys.argv[2]) if len(sys.argv) > 2 else 150
lib = sys.argv[3] if len(sys.argv) > 3 else "mdz_b200/libmdzcuda.so"
sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, lib], capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
print("instructions:", len(ins))
loops = []
for addr, t in ins:
    m = re.search(r"BRA\S*\s+.*0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < addr:
            loops.append((tgt, addr, (addr - tgt) // 16 + 1))
ALU = ("LOP3", "SHF", "IADD3", "ISETP", "SEL", "FLO", "POPC", "PRMT", "VIADD", "IABS", "LEA", "MOV", "VIMNMX", "IMNMX", "PLOP3", "BREV", "SGXT", "VOTE", "R2P", "P2R", "CS2R")
FMA = ("IMAD", "FFMA", "FMUL", "FADD")
for lo, hi, n in loops:
    if n < min_n:
        continue
    c = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            c[t.split()[0].split(".")[0] + ("." + t.split()[0].split(".")[1] if t.split()[0].startswith("IMAD.") else "")] += 1
    tot = sum(c.values())
    alu = sum(v for k, v in c.items() if k.split(".")[0] in ALU)
    wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE"))
    fma = sum(v for k, v in c.items() if k.split(".")[0] in FMA)
    print("loop %#x..%#x: %d instr, ALU-class %d, IMAD* %d (WIDE %d)" % (lo, hi, tot, alu, fma, wide))
    print("    " + ", ".join("%s %d" % kv for kv in c.most_common(24)))
