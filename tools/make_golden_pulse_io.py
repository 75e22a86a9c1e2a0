#!/usr/bin/env python
"""Generate tests/golden/pulse_io.json from the UNMODIFIED reference (oracle/_ref): what pulse_data_load() /
rfraw_parse() read out of the texts of tests/test_pulse_io.py:ook_cases(), and what pulse_data_dump() /
pulse_data_print_vcd() print for seeded pulse_data_t structs.  Run in the build container."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_pulse_io as t  # noqa: E402
from oracle import refh  # noqa: E402

out = {"load": [], "dump": []}
for name, text in t.ook_cases().items():
    if name == "rfraw_full":  # 2600 characters of the same byte: covered by the live comparison only
        continue
    for rate in (250000, 1024000):
        out["load"].append({"name": name, "rate": rate, "text": text,
                            "packages": [t.pd_facts(p) for p in refh.load_ook(text, rate, cap=64)]})
rng = random.Random(23)
for i in range(6):
    pd = t.random_pulse_data(rng, fsk=i % 2 == 1)
    ook = refh.dump_ook(pd)
    n = int(pd["num_pulses"])
    rec = {k: int(pd[k]) for k in t.INT_FIELDS}
    rec.update({k: float(pd[k]) for k in t.FLOAT_FIELDS})
    rec["pulse"] = [int(v) for v in pd["pulse"][:n]]
    rec["gap"] = [int(v) for v in pd["gap"][:n]]
    out["dump"].append({"pd": rec, "ook": ook[ook.index("\n") + 1:], "vcd": refh.dump_vcd(pd, "'")})
with open(t.GOLDEN, "w") as f:
    json.dump(out, f, indent=0, separators=(",", ":"))
print(t.GOLDEN, os.path.getsize(t.GOLDEN), "bytes,", len(out["load"]), "load cases,", len(out["dump"]), "dump cases")

# ---- pulse analyzer: what pulse_analyzer() prints for seeded packages + the hashes of its trial demodulation
import numpy as np  # noqa: E402
import test_analyzer as ta  # noqa: E402
from rtl_433_b200 import lib  # noqa: E402

GOLDEN_AN = os.path.join(os.path.dirname(t.GOLDEN), "analyzer.json")
rng = random.Random(41)
ref = refh.Ref(store_bitbuffers=False)
cases = []
for rep in range(2):
    for kind in ta.KINDS:
        pd = ta.shaped_package(rng, kind, 250000 if rep == 0 else 1024000)
        pd["ook_low_estimate"] = pd["ook_high_estimate"] = 0  # loaded packages carry no level estimates
        typ = 2 if (rep == 1 and kind in ("pcm", "pwm_fixed_gap", "manchester")) else 1
        if typ == 2:
            pd["fsk_f2_est"], pd["fsk_f1_est"] = 2500, -1800
        text, hashes = ref.analyze(pd, typ)
        n = int(pd["num_pulses"])
        cases.append({"kind": kind, "rate": int(pd["sample_rate"]), "type": typ, "fsk_f1_est": int(pd["fsk_f1_est"]),
                      "fsk_f2_est": int(pd["fsk_f2_est"]), "pulse": [int(v) for v in pd["pulse"][:n]],
                      "gap": [int(v) for v in pd["gap"][:n]], "text": text, "hashes": [str(h) for h in hashes]})
with open(GOLDEN_AN, "w") as f:
    json.dump({"cases": cases}, f, indent=0, separators=(",", ":"))
print(GOLDEN_AN, os.path.getsize(GOLDEN_AN), "bytes,", len(cases), "analyzer cases")
