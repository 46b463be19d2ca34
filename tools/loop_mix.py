#!/usr/bin/env python3
"""Instruction mix of a kernel's main loop from the built library (cuobjdump -sass): the loop is taken to be the span
of the longest backward branch.  usage: python tools/loop_mix.py <substring of the mangled kernel name> [lib]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else "kopek_b200/lib/libkopek_cuda.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0].strip()
    if sys.argv[1] not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", blk)]
    best = (0, 0, 0)
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > best[0]:
            best = (a - int(m.group(1), 16), int(m.group(1), 16), a)
    loop = [t for a, t in ins if best[1] <= a <= best[2]]
    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in loop)
    regs = re.search(r"REG:(\d+)", subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout.split(name)[1]) if False else None
    print(f"{name}\n  total {len(ins)}  loop {len(loop)}  " + " ".join(f"{k}:{v}" for k, v in c.most_common(40)))
