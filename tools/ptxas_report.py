#!/usr/bin/env python3
"""Registers / spills / stack per kernel from `nvcc -Xptxas -v` output on stdin."""
import re, sys
name = None
for line in sys.stdin:
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        name = m.group(1); continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and name:
        stack = m.groups(); continue
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        short = re.sub(r"^_ZN9cosmolike\d+", "", name)
        short = re.sub(r"ILi(\d)ELi(\d)ELi(\d)EE.*", r"<\1,\2,\3>", short)
        print(f"{short[:60]:60s} regs {m.group(1):>3s}  stack {stack[0]:>4s}  spill st/ld {stack[1]:>4s}/{stack[2]:>4s}")
        name = None
