#!/usr/bin/env python
"""Registers / spills / shared memory of every kernel in libvivim_b200.so (ptxas -v output of a forced rebuild).

    python scripts/ptxas_report.py [substring ...]     # filter by demangled-name substrings
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vivim_b200 import build as vb  # noqa: E402

cmd = [vb.find_nvcc()] + vb.NVCC_FLAGS + ["-Xptxas", "-v", "-ccbin", "/usr/bin/g++", "-o", vb.LIB_PATH,
                                          os.path.join(vb.CSRC, "vivim_b200.cu")]
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode != 0:
    sys.exit(res.stdout + res.stderr)
rows, name = [], None
for line in res.stderr.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void vv::", "").replace("__nv_bfloat16", "bf16").replace("__half", "f16")
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and name:
        spill = (int(m.group(2)), int(m.group(3)))
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        sm = re.search(r"(\d+) bytes smem", line)
        rows.append((name, int(m.group(1)), spill, int(sm.group(1)) if sm else 0))
        name = None
flt = sys.argv[1:]
for n, r, sp, sm in sorted(rows):
    if all(f in n for f in flt):
        print(f"{r:4d} regs  spill {sp[0]:4d}/{sp[1]:4d}  smem {sm:6d}  {n}")
