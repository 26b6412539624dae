#!/usr/bin/env python
"""Opcode histogram of the kernels in an object / shared library (cuobjdump -sass), per kernel, with the tensor-core and
tensor-memory mnemonics called out: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, HMMA = legacy mma.sync.

    python tools/sass_histogram.py vsrd_b200/libvsrd_b200.so [kernel regex] > profiles/rNN_sass_opcodes.txt
"""
import collections
import re
import subprocess
import sys

path = sys.argv[1]
pattern = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
kernels, current = collections.OrderedDict(), None
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        current = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kernels[current] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and current:
        kernels[current][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCQMMA", "UTCMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "HMMA", "UTMALDG", "SYNCS", "MUFU", "FFMA2", "FMUL2", "FADD2")
for name, hist in kernels.items():
    if pattern and not pattern.search(name):
        continue
    total = sum(hist.values())
    if total < 50:
        continue
    flags = ", ".join(f"{k} {hist[k]}" for k in KEY if hist.get(k))
    print(f"== {name}: {total} instructions; {flags}")
    print("   " + ", ".join(f"{op} {n}" for op, n in hist.most_common(14)))
