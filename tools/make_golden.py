#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).  Run in the build
container (needs /root/reference compiled by oracle/Makefile).  Each fixture holds the input IQ
bytes, the run parameters and what the reference produced: package headers, pulse/gap widths,
every event as (device index, FNV-1a of the 6604-byte bitbuffer_t), CRC32 of the AM/FM stage
arrays, and the decoders' JSON output.
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refh  # noqa: E402
from rtl_433_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
PKG_KEYS = ["type", "block", "offset", "start_ago", "end_ago", "num_pulses", "ook_low_estimate", "ook_high_estimate",
            "fsk_f1_est", "fsk_f2_est", "pulse_count", "num_events"]
PKG_FLOATS = ["freq1_hz", "freq2_hz", "rssi_db", "snr_db", "noise_db", "sample_file_pos"]


def pad16(x, fill):
    n = (x.nbytes + 15) // 16 * 16
    if n == x.nbytes:
        return x
    extra = np.full((n - x.nbytes) // x.itemsize, fill, x.dtype)
    return np.concatenate([x, extra])


CASES = {
    "nice_flor_s": dict(iq=pad16(synth.nice_flor_s_file(), 128), fmt=2, rate=250000, freq=433920000, fpdm=2, protocols=[169]),
    "silvercrest_r1": dict(iq=pad16(synth.silvercrest_file(), 128), fmt=2, rate=250000, freq=433920000, fpdm=2, protocols=[1]),
    "ook_noisy": dict(iq=synth.ook_stream(5, n_samples=1 << 17, n_bursts=2, kinds=("silvercrest", "nexus"), decodable=True), fmt=2, rate=250000,
                      freq=433920000, fpdm=2, protocols=None),
    "fsk_minmax": dict(iq=synth.fsk_stream(3, n_samples=1 << 16, n_bursts=1), fmt=4, rate=1024000, freq=868000000, fpdm=2,
                       protocols=None),
    "fsk_classic": dict(iq=synth.fsk_stream(4, n_samples=1 << 16, n_bursts=1), fmt=4, rate=1024000, freq=433920000, fpdm=0,
                        protocols=None),
}

for name, c in CASES.items():
    def make(chain):
        r = refh.Ref(chain_decoders=chain, store_bitbuffers=False, store_stages=True)
        if c["protocols"] is None:
            r.register_defaults()
        else:
            for p in c["protocols"]:
                r.register(p)
        return r

    # events are recorded with decoders stubbed out (every priority class runs, src/r_api.c:444);
    # a second run with the real decoders chained yields the decoded JSON
    r = make(False)
    devs = r.registered()
    res = r.run(c["iq"], c["fmt"], c["rate"], c["freq"], c["fpdm"])
    r2 = make(True)
    res["json"] = r2.run(c["iq"], c["fmt"], c["rate"], c["freq"], c["fpdm"])["json"]
    r2.close()
    pk = res["packages"]
    arrays = {
        "iq": np.ascontiguousarray(c["iq"]).view(np.uint8),
        "params": np.array([c["fmt"], c["rate"], c["freq"], c["fpdm"]], np.int64),
        "protocol_nums": np.array([d["protocol_num"] for d in devs], np.int32),
        "pkg_int": np.array([[p[k] for k in PKG_KEYS] for p in pk], np.int64).reshape(len(pk), len(PKG_KEYS)),
        "pkg_float": np.array([[p[k] for k in PKG_FLOATS] for p in pk], np.float32).reshape(len(pk), len(PKG_FLOATS)),
        "pulses": np.concatenate([p["pulse"] for p in pk]) if pk else np.zeros(0, np.int32),
        "gaps": np.concatenate([p["gap"] for p in pk]) if pk else np.zeros(0, np.int32),
        "ev_package": np.array([e["package"] for e in res["events"]], np.uint32),
        "ev_dev": np.array([e["dev"] for e in res["events"]], np.uint32),
        "ev_hash": np.array([e["hash"] for e in res["events"]], np.uint64),
        "stage_crc": np.array([zlib.crc32(res["am"].tobytes()), zlib.crc32(res["fm"].tobytes())], np.uint64),
        "json": np.frombuffer("\n".join(res["json"]).encode(), np.uint8),
    }
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print(name, len(c["iq"]), "bytes iq,", len(pk), "packages,", len(res["events"]), "events,", len(res["json"]), "decoded:",
          res["json"][:2])
    r.close()
with open(os.path.join(OUT, "README.md"), "w") as f:
    f.write("Golden vectors produced by tools/make_golden.py from the unmodified reference\n"
            "(merbanan/rtl_433 25.12, compiled by oracle/Makefile into oracle/_ref).\n"
            "pkg_int columns: " + ", ".join(PKG_KEYS) + "\npkg_float columns: " + ", ".join(PKG_FLOATS) + "\n")
