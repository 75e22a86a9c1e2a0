"""CPU checks (on the product's own __host__ __device__ functions, through tests/host_core.cpp) of
the two invariants the warp-level shortcuts of k_detect rest on; the shortcuts themselves are
covered bit-for-bit by the -m gpu parity tests."""
import ctypes as C

import helpers


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
