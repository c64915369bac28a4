#!/usr/bin/env python
"""Registers / stack / spill bytes per kernel: an older ptxas report against the current build's
(mdz_b200/csrc/ptxas_report.txt, written by the Makefile).  usage: ptxas_diff.py OLD [filter...]"""
import re
import subprocess
import sys


def parse(path):
    out, cur, spill = {}, None, None
    for line in open(path).read().splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur:
            spill = tuple(int(x) for x in m.groups())
            continue
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            out[cur] = (int(m.group(1)), spill)
            cur = spill = None
    return out


old, new = parse(sys.argv[1]), parse("mdz_b200/csrc/ptxas_report.txt")
names = subprocess.run(["c++filt"], input="\n".join(sorted(new)), capture_output=True, text=True).stdout.split("\n")
for k, d in zip(sorted(new), names):
    d = d.replace("mdz::", "").replace("(EscapeParams)", "")
    if sys.argv[2:] and not any(f in d for f in sys.argv[2:]):
        continue
    if k not in old:
        print("%-50s new %s" % (d, new[k]))
    elif old[k] != new[k]:
        print("%-50s %s -> %s" % (d, old[k], new[k]))
