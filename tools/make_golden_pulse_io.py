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
