"""CPU checks (on the product's own __host__ __device__ functions, through tests/host_core.cpp) of
the two invariants the warp-level shortcuts of k_detect rest on; the shortcuts themselves are
covered bit-for-bit by the -m gpu parity tests."""
import ctypes as C

import numpy as np

import helpers
from rtl_433_b200 import lib


def _lib():
    L = C.CDLL(helpers.build_hostcore())
    L.hc_check_pulse_bound.argtypes = [C.c_uint, C.c_int]
    L.hc_check_bracket_rebuild.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int]
    return L


def test_pulse_threshold_bound():
    """PULSE fast path: the chunk maximum bounds the high estimate and every "below" threshold."""
    assert _lib().hc_check_pulse_bound(1, 200000) == 0


def test_fm_filter_state_rebuild_is_sound_and_usually_converges():
    """FM on demand: monotone filter => once the two extreme start states meet, all states have met.
    A live (noisy) discriminator makes them meet within one 1024/512-sample tile; an exactly
    constant one does not -- the case the kernel handles by walking forward from the last exact state."""
    L = _lib()
    for cs16, rate, tile in ((0, 250000, 1024), (0, 1024000, 1024), (1, 1024000, 512), (1, 250000, 512)):
        n = L.hc_check_bracket_rebuild(7, 400, cs16, rate, tile, 300)
        assert n >= 0, "a start state escaped the bracket"
        assert n >= 396, f"only {n}/400 noisy sequences converged within one tile (cs16={cs16}, rate={rate})"
        n = L.hc_check_bracket_rebuild(7, 50, cs16, rate, tile, 0)
        assert n >= 0
    # constant input: the floor map keeps distinct fixed points apart for at least some levels
    assert L.hc_check_bracket_rebuild(9, 200, 0, 250000, 1024, 0) < 200


def test_slice_work_lists_cover_every_device_once_and_align_large_modulations():
    """k_slice's work items are 32 consecutive slots of these lists: every device that takes the
    package type appears exactly once, holes are padding only, a modulation with >= 16 devices
    starts on a warp boundary, and devices of one modulation stay together."""
    devs = lib.default_device_table(include_disabled=True)
    hc = helpers.HostCore()
    hc.add_devices(devs)
    hc.L.hc_slice_list.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    for ptype, takes in ((1, lambda m: 3 <= m <= 13 and m != 7), (2, lambda m: 16 <= m <= 18)):
        buf = np.zeros(4096, np.uint32)
        n = hc.L.hc_slice_list(hc.h, ptype, buf.ctypes.data, len(buf))
        lst = buf[:n]
        real = [int(x) for x in lst if x != 0xFFFFFFFF]
        want = [i for i, d in enumerate(devs) if takes(d["modulation"])]
        assert sorted(real) == want
        mods = [devs[i]["modulation"] for i in real]
        assert mods == sorted(mods), "devices of one modulation are contiguous, modulations ascending"
        count = {m: mods.count(m) for m in set(mods)}
        seen = set()
        for slot, x in enumerate(lst):
            if x == 0xFFFFFFFF:
                continue
            m = devs[int(x)]["modulation"]
            if m not in seen:
                seen.add(m)
                if count[m] >= 16:
                    assert slot % 32 == 0, f"modulation {m} ({count[m]} devices) starts at slot {slot}"
        assert n - len(real) < 32 * len(count)
