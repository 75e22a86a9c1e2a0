#!/usr/bin/env python
"""Write rtl_433_b200/data/devices_25.12.json: the slicer-relevant fields (modulation, six
timing floats, priority, disabled) of every r_device the reference registers, in DEVICES
order (include/rtl_433_devices.h; protocol_num = index + 1, src/r_api.c:133-142).

Run in the build container (needs oracle/_ref, i.e. /root/reference).  The table is protocol
data, not code; tests/test_abi.py re-checks it against the compiled reference when present.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refh  # noqa: E402

r = refh.Ref()
devs = r.protocols()
out = {"reference": "merbanan/rtl_433 25.12 (91b2ebdb)", "count": len(devs), "devices": devs}
path = os.path.join(ROOT, "rtl_433_b200", "data", "devices_25.12.json")
with open(path, "w") as f:
    json.dump(out, f, indent=0, separators=(",", ":"))
    f.write("\n")
print(path, len(devs), "protocols,", sum(1 for d in devs if d["disabled"] == 0), "enabled by default")
