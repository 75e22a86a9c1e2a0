#!/usr/bin/env python
"""Aggregate the source page of an .ncu-rep (captured with --import-source on) by code region:
share of stall samples, share of executed warp instructions, warp instructions per IQ sample and
average active lanes.  Regions are found by marker strings in the current sources, so run it on a
report taken from the same tree.   usage: ncu_regions.py REPORT.ncu-rep detect|slice [N_SAMPLES]"""
import collections
import csv
import io
import os
import subprocess
import sys
import os as _os
KF = ["-k", "regex:" + _os.environ["NCU_KERNEL"]] if _os.environ.get("NCU_KERNEL") else []

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, which = sys.argv[1], sys.argv[2]
n_samples = float(sys.argv[3]) if len(sys.argv) > 3 else 4096.0 * (1 << 20)

MARKS = {
    "detect": ("r433b_detect.cuh", [
        ("load_group / env_group", "void load_group"), ("fm: discriminator (disc_fill)", "void disc_fill("),
        ("fm: state rebuild (fm_cold)", "void fm_cold("), ("fm: window low-pass (fm_window)", "void fm_window("),
        ("fm: demand", "void fm_demand("), ("deferred carrier estimate (f1_evaluate)", "int f1_step("),
        ("prologue", "k_detect(DetectParams p)"), ("log_append", "auto log_append = [&]"), ("emit", "auto emit = [&]"),
        ("am_repair", "void am_repair("),
        ("tile loop top / AM tile load + hand-over check", "for (unsigned long long t0 = p.sample_begin"),
        ("tile FM pass / fm_need", "// ---- FM for the whole tile when it cannot be made on demand"),
        ("idle_tile", "auto idle_tile = [&]"), ("idle_fast", "auto idle_fast = [&]"), ("gap_fast", "auto gap_fast = [&]"),
        ("pulse_fast", "auto pulse_fast = [&]"), ("pulse0_fast (first pulse + FSK)", "auto pulse0_fast = [&]"),
        ("gapstart_fast", "auto gapstart_fast = [&]"), ("walk loop + det_step call", "for (int n = 0; n < nv_tile;) {"),
        ("flush / save", "// flush_sdr_flow()")]),
    "slice": ("r433b_slice.cuh", [
        ("EventWriter", "struct EventWriter"), ("slicer helpers", "struct PulseView"), ("slicer_begin / slicer_step (front end)", "slicer_begin("),
        ("slicer_apply (back end)", "slicer_apply("), ("slice_dispatch loop", "slice_dispatch(")]),
}
fname, marks = MARKS[which]
src_lines = open(os.path.join(ROOT, "rtl_433_b200", "csrc", fname)).read().split("\n")
pos = []
for name, needle in marks:
    for i, l in enumerate(src_lines):
        if needle in l:
            pos.append((i + 1, name))
            break
pos.sort()

out = subprocess.run(["ncu", "-i", rep] + KF + [ "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur, hd = None, None
reg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
last_region = None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hd = r
        continue
    if hd is None or cur is None or r[0] == "Function Name":
        continue
    d = dict(zip(hd, r))
    try:
        ln = int(d["Line No"])
    except ValueError:
        # a SASS row under the last source line: static code size of the region
        if last_region is not None and len(r) > 2 and r[2].startswith("0x"):
            reg[last_region][3] += 1
        continue

    def f(k):
        try:
            return float(d.get(k, 0) or 0)
        except ValueError:
            return 0.0
    name = cur
    if cur == fname:
        name = "(before first marker)"
        for l0, nm in pos:
            if ln >= l0:
                name = nm
        name = f"{fname}: {name}"
    v = reg[name]
    last_region = name
    v[0] += f("# Samples")
    v[1] += f("Instructions Executed")
    v[2] += f("Thread Instructions Executed")
ts, ti = sum(v[0] for v in reg.values()), sum(v[1] for v in reg.values())
print("| region | stall samples | warp instructions | warp-instr / IQ sample | avg active lanes | SASS instructions (static) |")
print("|---|---|---|---|---|---|")
for k, (s, i, t, c) in sorted(reg.items(), key=lambda kv: -kv[1][1]):
    if i / ti < 0.002 and c < 200:
        continue
    print(f"| {k} | {100 * s / ts:.1f} % | {100 * i / ti:.1f} % | {i / n_samples:.2f} | {t / max(i, 1):.1f} | {c} |")
print(f"\nstatic code: {sum(v[3] for v in reg.values())} SASS instructions = {sum(v[3] for v in reg.values()) * 16 / 1024:.0f} KiB")
print(f"\ntotal (source view): {ti / n_samples:.2f} warp instructions per IQ sample")
