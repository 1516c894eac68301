#!/usr/bin/env python
"""Registers / stack (spills + local arrays) of every kernel in libnlkernels.so: `python tools/res_usage.py [filter]`."""
import re
import subprocess
import sys

so = "numpyllvm_b200/libnlkernels.so"
txt = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
names = re.findall(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", txt)
dem = subprocess.run(["c++filt"], input="\n".join(n[0] for n in names), capture_output=True, text=True).stdout.splitlines()
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for (n, reg, stack, sh), d in zip(names, dem):
    d = re.sub(r"\(nl::KParams\)$", "", d).replace("nl::", "").replace("(unsigned int)", "")
    d = re.sub(r"void ", "", d)
    if pat in d:
        print(f"{int(reg):4d} regs {int(stack):5d} stack {int(sh):6d} smem  {d[:200]}")
