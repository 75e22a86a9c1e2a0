#!/usr/bin/env python
"""Static code size of k_detect by source line range (no GPU needed): disassembles the built library with line
info (nvdisasm -g) and counts SASS instructions per source file / line bucket, so that the instruction-cache
footprint of each part of the kernel can be seen before spending GPU time.
usage: sass_regions.py [function-substring] [bucket-lines]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rtl_433_b200", "csrc", "libr433b.so")
want = sys.argv[1] if len(sys.argv) > 1 else "k_detectILi2E"
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
fn, cur = None, None
count = collections.Counter()
inl = collections.Counter()
for line in out.split("\n"):
    if line.startswith(".text."):
        fn = line.strip()
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if fn and want in fn and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        if cur:
            count[cur] += 1
total = sum(count.values())
print(f"{want}: {total} SASS instructions = {total * 16 / 1024:.0f} KiB")
if bucket:
    b = collections.Counter()
    for (f, l), n in count.items():
        b[(f, l // bucket * bucket)] += n
    for (f, l), n in sorted(b.items()):
        if n >= 20:
            print(f"{f}:{l:5d}-{l + bucket - 1:5d}  {n:6d}")
else:
    # regions by the marker strings of tools/ncu_regions.py
    sys.argv = [sys.argv[0], "x", "detect"]
    src = open(os.path.join(ROOT, "tools", "ncu_regions.py")).read()
    marks_src = src[src.index("MARKS = {"):src.index("fname, marks = MARKS[which]")]
    ns = {}
    exec(marks_src, ns)
    fname, marks = ns["MARKS"]["detect"]
    lines = open(os.path.join(ROOT, "rtl_433_b200", "csrc", fname)).read().split("\n")
    pos = []
    for name, needle in marks:
        for i, l in enumerate(lines):
            if needle in l:
                pos.append((i + 1, name))
                break
    pos.sort()
    reg = collections.Counter()
    for (f, l), n in count.items():
        name = f
        if f == fname:
            name = "(before first marker)"
            for l0, nm in pos:
                if l >= l0:
                    name = nm
        reg[name] += n
    for k, n in sorted(reg.items(), key=lambda kv: -kv[1]):
        print(f"{n:6d}  {k}")
