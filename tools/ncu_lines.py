#!/usr/bin/env python
"""Per-source-line table of an .ncu-rep (captured with --import-source on): executed warp instructions
and stall samples per line of one file, in line order.
usage: ncu_lines.py REPORT.ncu-rep FILE.cuh FIRST LAST [N_SAMPLES]   (NCU_KERNEL=regex selects a kernel)"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KF = ["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
n_samples = float(sys.argv[5]) if len(sys.argv) > 5 else 4096.0 * (1 << 20)
out = subprocess.run(["ncu", "-i", rep] + KF + ["--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
src = open(os.path.join(ROOT, "rtl_433_b200", "csrc", fname)).read().split("\n")
cur, hd = None, None
acc = collections.defaultdict(lambda: [0.0, 0.0])
tot = [0.0, 0.0]
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hd = r
        continue
    if hd is None or cur is None:
        continue
    d = dict(zip(hd, r))
    try:
        ln = int(d["Line No"])
        s, i = float(d.get("# Samples", 0) or 0), float(d.get("Instructions Executed", 0) or 0)
    except ValueError:
        continue
    tot[0] += s
    tot[1] += i
    if cur == fname:
        acc[ln][0] += s
        acc[ln][1] += i
for ln in range(lo, hi + 1):
    if ln in acc and acc[ln][1] > 0:
        s, i = acc[ln]
        print(f"{ln:5d} {100 * s / tot[0]:5.1f}% {100 * i / tot[1]:5.1f}% {i / n_samples:6.3f}  {src[ln - 1].strip()[:110]}")
