"""cf32 captures: the reference converts them to cs16 while reading (src/rtl_433.c:1811-1825:
"clamp float to [-1,1] and scale to Q0.15"); everything downstream is the cs16 path."""
import numpy as np
import pytest

import helpers
from oracle import orc
from rtl_433_b200 import lib, synth


def test_oracle_conversion_known_answers():
    o = orc.Oracle(store_bitbuffers=False)
    x = np.array([0.0, 0.5, 1.0, -1.0, 1.5, -2.0, 3.0518e-5, -3.0518e-5, 0.99999, np.nan, 1e20, -1e20, np.inf], np.float32)
    got = o.cf32_to_cs16(x)
    # truncation toward zero, clamp to +-32767; NaN / beyond-int32 products become INT_MIN on the
    # reference's x86-64 builds and therefore -32767 after the clamp
    want = [0, 16383, 32767, -32767, 32767, -32767, 0, 0, 32766, -32767, -32767, -32767, -32767]
    assert got.tolist() == want
    rng = np.random.default_rng(3)
    r = (rng.standard_normal(100000) * 0.6).astype(np.float32)
    t = np.trunc(r.astype(np.float32) * np.float32(32767)).clip(-32767, 32767).astype(np.int16)
    assert np.array_equal(o.cf32_to_cs16(r), t)


@pytest.mark.gpu
def test_cf32_input_matches_converted_cs16():
    devices = lib.default_device_table()
    o = orc.Oracle(store_bitbuffers=False, store_stages=True)
    o.add_devices(devices)
    rng = np.random.default_rng(5)
    streams = []
    for seed, gain in ((61, 1.0), (62, 2.5)):  # the second one clips
        cs = synth.fsk_stream(seed, n_samples=1 << 19)
        f = (cs.astype(np.float32) / np.float32(32768.0)) * np.float32(gain)
        f[rng.integers(0, 2000, 16)] = np.array([np.nan, 1e20, -1e20, np.inf, -np.inf, 7.0, -7.0, 1.0, -1.0, 0.0, 1e-9, -1e-9,
                                                 0.99999, -0.99999, 65540.0, -65540.0], np.float32)
        streams.append(f)
    lens = [s.nbytes for s in streams]
    assert all(n % 32 == 0 for n in lens)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ctx = lib.Context(0)
    try:
        ctx.set_devices(devices)
        for stages in (False, True):
            ctx.process(np.concatenate(streams), offsets, lib.FMT_CF32, 1024000, 868000000, want_stages=stages)
            ctx.fetch()
            for i, f in enumerate(streams):
                cs = o.cf32_to_cs16(f)
                ref = o.run(cs, 4, 1024000, 868000000)
                assert any(p["type"] == 2 for p in ref["packages"])
                got = helpers.gpu_stream_results(ctx, i)
                if stages:
                    got["am"], got["fm"] = ctx.copy_stage(i, len(cs) // 2)
                d = helpers.compare_results(ref, got, f"cf32 stream {i}", stages=stages)
                assert not d, "\n".join(d[:20])
    finally:
        ctx.close()
