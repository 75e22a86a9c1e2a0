"""-m gpu: the drop-in claim end to end.  The reference's UNMODIFIED decoders (the r_device structs
registered inside oracle/_ref) are driven by the product's r433b_dispatch_r_devices() from GPU
results; decoded JSON and every decoder's decode_events/ok/messages/fails counters must equal a
pure-reference run of the same capture (real decoders chained, so priority gating is live)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import refh
from rtl_433_b200 import lib, synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refh.available(), reason="oracle/_ref not shipped")]


def test_reference_decoders_behind_gpu_path():
    x = synth.ook_stream(5, n_samples=1 << 19, n_bursts=4, kinds=("silvercrest", "nexus", "nice"), decodable=True)
    r = refh.Ref(chain_decoders=True, store_bitbuffers=False)
    n = r.register_defaults()
    devs = r.registered()
    # 1. the reference alone
    want = r.run(x, 2)
    want_stats = [r.device_stats(i) for i in range(n)]
    assert want["json"], "the capture should decode to something"
    # 2. GPU path + the same r_device structs
    ctx = lib.Context(0)
    ctx.set_devices(devs)
    ctx.process(x, np.array([0, x.nbytes], np.uint64), lib.FMT_CU8, 250000, 433920000)
    ctx.fetch()
    ptrs = r.L.refh_begin_external_dispatch(r.h)
    try:
        rc = ctx.L.r433b_dispatch_r_devices(ctx.h, C.byref(ctx._res), 0, ptrs, n)
        assert rc == 0, ctx.L.r433b_last_error(ctx.h)
        got_json = [l for l in r.L.refh_json(r.h).decode().split("\n") if l]
        got_stats = [r.device_stats(i) for i in range(n)]
    finally:
        r.L.refh_end_external_dispatch(r.h)
    assert got_json == want["json"]
    assert got_stats == want_stats
    assert sum(s[0] for s in got_stats) > 1000  # thousands of decode_fn calls took place
    ctx.close()


SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "shim_c99")


@pytest.mark.skipif(not os.path.exists(SHIM), reason="oracle/_ref/shim_c99 not built/shipped (make -C oracle shim)")
def test_c99_shim_of_integration_md(tmp_path):
    """INTEGRATION.md section 2, compiled verbatim as C99 against the reference's headers and linked with the
    reference's unmodified objects + libr433b.so (oracle/Makefile `shim`): the reference's decoders, its
    data_acquired_handler and its JSON printer run behind the GPU path.  Two capture files in one batch; the
    messages must be the ones a pure-reference run decodes, in order."""
    files = [synth.silvercrest_file(),
             synth.ook_stream(5, n_samples=1 << 19, n_bursts=4, kinds=("silvercrest", "nexus", "nice"), decodable=True)]
    want = []
    r = refh.Ref(chain_decoders=True, store_bitbuffers=False)
    r.register_defaults()
    for x in files:
        want += [json.loads(l) for l in r.run(x, 2)["json"] if l]
    assert any(m.get("model") == "Silvercrest-Remote" for m in want)
    paths = []
    for i, x in enumerate(files):
        p = tmp_path / f"g{i:03d}_433.92M_250k.cu8"
        x.tofile(p)
        paths.append(str(p))
    out = subprocess.run([SHIM, "250000"] + paths, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = [json.loads(l) for l in out.stdout.split("\n") if l.strip()]
    assert got == want
