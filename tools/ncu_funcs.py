#!/usr/bin/env python
"""Aggregate the SASS page of an .ncu-rep by device FUNCTION (the noinline phases of k_detect): executed warp
instructions, stall samples and the top stall reasons per function.  Function ranges come from the symbol table
of the cubin inside the library the report was taken from (same tree!).
usage: ncu_funcs.py REPORT.ncu-rep KERNEL_MANGLED_SUBSTRING [N_SAMPLES]"""
import collections
import csv
import io
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
n_samples = float(sys.argv[3]) if len(sys.argv) > 3 else 4096.0 * (1 << 20)
lib = os.path.join(ROOT, "rtl_433_b200", "csrc", "libr433b.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
syms = subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout
funcs = []  # (offset, size, name) inside the kernel's section
import re
rows_ = [re.sub(r"\[<other>: \w+\]", "", l).split() for l in syms.split("\n")]
rows_ = [f for f in rows_ if len(f) >= 8 and f[3] == "FUNC"]
ksec = [f[6] for f in rows_ if kern in f[7] and not f[7].startswith("$")][0]
for f in rows_:
    if f[6] == ksec and f[7].startswith("$"):
        short = f[7].split("$")[-1]
        short = subprocess.run(["c++filt", short], capture_output=True, text=True).stdout.strip().split("(")[0]
        funcs.append((int(f[1], 16), int(f[2], 0), short))
funcs.sort()
KF = ["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []
out = subprocess.run(["ncu", "-i", rep] + KF + ["--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ridx = {r: hdr.index(r) for r in reasons}
base = None
acc = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rows[2:]:
    try:
        addr = int(r[ia], 16)
    except (ValueError, IndexError):
        continue
    if base is None:
        base = addr
    off = addr - base
    name = "(kernel body)"
    for o, sz, nm in funcs:
        if nm != "(kernel body)" and o <= off < o + sz:
            name = nm
    a = acc[name]
    a["samples"] += float(r[isamp] or 0)
    a["inst"] += float(r[iex] or 0)
    a["static"] += 1
    for k, i in ridx.items():
        a[k] += float(r[i] or 0)
ts = sum(a["samples"] for a in acc.values())
ti = sum(a["inst"] for a in acc.values())
print("| function | stall samples | warp instructions | warp-instr / IQ sample | SASS (static) | top stall reasons |")
print("|---|---|---|---|---|---|")
for name, a in sorted(acc.items(), key=lambda kv: -kv[1]["inst"]):
    top = sorted(((a[k], k) for k in reasons), reverse=True)[:3]
    tops = ", ".join(f"{k[6:]} {100 * v / max(a['samples'], 1):.0f}%" for v, k in top)
    print(f"| {name} | {100 * a['samples'] / ts:.1f} % | {100 * a['inst'] / ti:.1f} % | {a['inst'] / n_samples:.3f} | {int(a['static'])} | {tops} |")
print(f"\ntotal: {ti / n_samples:.2f} warp instructions per IQ sample")
