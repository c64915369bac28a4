#!/usr/bin/env python
"""Registers, local memory, shared memory and resident blocks per SM for every limb count
(needs a GPU: the occupancy figures come from cudaOccupancyMaxActiveBlocksPerMultiprocessor
through mdzcuda_plan_kernel_info).  Prints a markdown table."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mdz_b200
from views import make_view

spills = {}
rep = open(os.path.join(ROOT, "mdz_b200", "csrc", "ptxas_report.txt")).read()
for m in re.finditer(r"Compiling entry function '_ZN3mdz18escape_mpfr_kernelILi(\d+)ELb(\d)E.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", rep):
    spills[(int(m.group(1)), int(m.group(2)))] = (int(m.group(3)), int(m.group(4)), int(m.group(5)))
for m in re.finditer(r"Compiling entry function '_ZN3mdz18escape_gmpf_kernelILi(\d+)E.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", rep):
    spills[("g", int(m.group(1)))] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))

print("| mode | precision (bits) | 32-bit limbs | registers/thread | stack frame (B) | spill stores / loads (B, static) | dynamic smem/block (B) | blocks/SM x 128 threads | warps/SM |")
print("|---|---|---|---|---|---|---|---|---|")
rows = [("long double", "ld", 64)] + [("MPFR", "mpfr", p) for p in range(96, 1025, 32)]
for label, mode, p in rows:
    v = make_view("-0.5", "0", "3", 32, 24, mode=mode, precision=max(p, 80) if mode == "ld" else p, depth=10)
    plan = mdz_b200.Plan(v, 0)
    plan.launch(); plan.wait()
    ki = plan.kernel_info()
    plan.close()
    st = spills.get((ki["limbs"], 0), ("?", "?", "?"))
    print("| %s | %d | %d | %d | %s | %s / %s | %d | %d | %d |" % (label, p, ki["limbs"], ki["regs_per_thread"], st[0], st[1], st[2],
                                                              ki["shared_bytes"], ki["blocks_per_sm"], ki["blocks_per_sm"] * 4))
for p in (128, 192, 256, 320, 384, 448, 512, 576, 640, 704, 768, 832, 896):
    v = make_view("-0.5", "0", "3", 32, 24, mode="gmp", precision=p, depth=10)
    plan = mdz_b200.Plan(v, 0)
    plan.launch(); plan.wait()
    ki = plan.kernel_info()
    plan.close()
    st = spills.get(("g", ki["limbs"]), ("?", "?", "?"))
    print("| GMP mpf | %d | %d | %d | %s | %s / %s | %d | %d | %d |" % (p, ki["limbs"], ki["regs_per_thread"], st[0], st[1], st[2],
                                                                    ki["shared_bytes"], ki["blocks_per_sm"], ki["blocks_per_sm"] * 4))
for mode, precs in (("mpfr", (2048, 4096, 6144, 8192)), ("gmp", (1024, 1856, 3904, 5952, 8000))):
    for p in precs:
        v = make_view("-0.5", "0", "3", 32, 24, mode=mode, precision=p, depth=10)
        plan = mdz_b200.Plan(v, 0)
        plan.launch(); plan.wait()
        ki = plan.kernel_info()
        plan.close()
        tt = ki["lanes_per_pixel"]
        kk = 6 if p in (6144, 5952) else 8
        m = re.search(r"Compiling entry function '_ZN3mdz18escape_coop_kernelILi%dELi%dELb%dEE.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads"
                      % (kk, tt, 1 if mode == "gmp" else 0), rep)
        st = m.groups() if m else ("?", "?", "?")
        print("| %s, %d lanes x %d words per pixel |" % ("GMP mpf" if mode == "gmp" else "MPFR", tt, kk)
              + " %d | %d | %d | %s | %s / %s | %d | %d | %d |" % (p, ki["limbs"], ki["regs_per_thread"], st[0], st[1], st[2],
                                                                 ki["shared_bytes"], ki["blocks_per_sm"], ki["blocks_per_sm"] * 4))
