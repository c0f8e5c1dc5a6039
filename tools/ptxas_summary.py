#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: registers / spills per kernel.  usage: ptxas_summary.py LOG [name-filter]"""
import re
import subprocess
import sys

t = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", re.S)
for m in pat.finditer(t):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    if flt in name:
        print(f"{name:60s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill st/ld {m.group(3)}/{m.group(4)}")
