"""CPU: pin the oracle (oracle/r433_oracle.c) and the product's device functions (run on the CPU
by tests/host_core.cpp) against the golden vectors produced by the unmodified reference
(tests/golden/, tools/make_golden.py).  These include the reference's own in-tree IQ vector
(tests/rtl_tcp_serve.py + tests/http-rtltcp-test.sh: Nice Flor-s, code e7a760b94372e)."""
import glob
import os
import zlib

import numpy as np
import pytest

import helpers
from oracle import orc, refh
from rtl_433_b200 import lib

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
PKG_KEYS = ["type", "block", "offset", "start_ago", "end_ago", "num_pulses", "ook_low_estimate", "ook_high_estimate",
            "fsk_f1_est", "fsk_f2_est", "pulse_count", "num_events"]
PKG_FLOATS = ["freq1_hz", "freq2_hz", "rssi_db", "snr_db", "noise_db", "sample_file_pos"]


def devices_of(g):
    table = {d["protocol_num"]: d for d in lib.default_device_table(include_disabled=True)}
    return [table[int(n)] for n in g["protocol_nums"]]


def check_against_golden(res, g, floats):
    pk = res["packages"]
    assert len(pk) == len(g["pkg_int"])
    got = np.array([[p[k] for k in PKG_KEYS] for p in pk], np.int64).reshape(len(pk), len(PKG_KEYS))
    assert np.array_equal(got, g["pkg_int"]), (got, g["pkg_int"])
    if floats:
        gf = np.array([[p[k] for k in PKG_FLOATS] for p in pk], np.float32).reshape(len(pk), len(PKG_FLOATS))
        assert np.array_equal(gf, g["pkg_float"])  # same libm on the same box; 1 ULP otherwise
    assert np.array_equal(np.concatenate([p["pulse"] for p in pk]) if pk else np.zeros(0, np.int32), g["pulses"])
    assert np.array_equal(np.concatenate([p["gap"] for p in pk]) if pk else np.zeros(0, np.int32), g["gaps"])
    assert [e["package"] for e in res["events"]] == list(g["ev_package"])
    assert [e["dev"] for e in res["events"]] == list(g["ev_dev"])
    assert [e["hash"] for e in res["events"]] == list(g["ev_hash"])
    assert zlib.crc32(res["am"].tobytes()) == int(g["stage_crc"][0])
    assert zlib.crc32(res["fm"].tobytes()) == int(g["stage_crc"][1])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(path):
    g = np.load(path)
    fmt, rate, freq, fpdm = (int(v) for v in g["params"])
    o = orc.Oracle(store_bitbuffers=False, store_stages=True)
    o.add_devices(devices_of(g))
    check_against_golden(o.run(g["iq"], fmt, rate, freq, fpdm), g, floats=True)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_device_functions_match_golden(path):
    g = np.load(path)
    fmt, rate, freq, fpdm = (int(v) for v in g["params"])
    hc = helpers.HostCore(store_bitbuffers=False, store_stages=True)
    hc.add_devices(devices_of(g))
    check_against_golden(hc.run(g["iq"], fmt, rate, freq, fpdm), g, floats=False)


def test_reference_known_answers():
    """The decoder-facing rows the reference documents: `{52}e7a760b94372e` (src/devices/nice_flor_s.c:134,
    tests/http-rtltcp-test.sh:35) and `{33}7c2600020` x4 (src/devices/silvercrest.c:27-36)."""
    table = {d["protocol_num"]: d for d in lib.default_device_table(include_disabled=True)}
    for name, proto, rows in (("nice_flor_s", 169, ["{52}e7a760b94372e"]), ("silvercrest_r1", 1, ["{33}7c2600020"] * 4)):
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
        for cls in (orc.Oracle, helpers.HostCore):
            o = cls(store_bitbuffers=True)
            o.add_devices([table[proto]])
            res = o.run(g["iq"], 2)
            bb = res["events"][0]["bitbuffer"]
            assert [refh.row_hex(bb, k) for k in range(int(bb["num_rows"]))][:len(rows)] == rows
        assert name in ("nice_flor_s", "silvercrest_r1")
        js = bytes(g["json"]).decode()
        assert ('"code":1139' in js) if proto == 169 else ('Silvercrest-Remote' in js)
