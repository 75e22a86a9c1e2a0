"""ctypes view of oracle/_ref/libr433ref.so -- TEST INFRASTRUCTURE ONLY.

The shared object holds the unmodified reference sources plus oracle/ref_harness.c.
Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline legs may
import this module; the product (rtl_433_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libr433ref.so")

BITBUF_ROWS, BITBUF_COLS = 50, 128


class Package(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("block", C.c_int32), ("offset", C.c_uint64),
        ("sample_rate", C.c_uint32), ("depth_bits", C.c_uint32), ("start_ago", C.c_uint32),
        ("end_ago", C.c_uint32), ("num_pulses", C.c_uint32),
        ("ook_low_estimate", C.c_int32), ("ook_high_estimate", C.c_int32),
        ("fsk_f1_est", C.c_int32), ("fsk_f2_est", C.c_int32),
        ("freq1_hz", C.c_float), ("freq2_hz", C.c_float), ("centerfreq_hz", C.c_float),
        ("range_db", C.c_float), ("rssi_db", C.c_float), ("snr_db", C.c_float), ("noise_db", C.c_float),
        ("sample_file_pos", C.c_float),
        ("pulse_off", C.c_uint32), ("pulse_count", C.c_uint32),
        ("first_event", C.c_uint32), ("num_events", C.c_uint32),
    ]


class Event(C.Structure):
    _fields_ = [("package", C.c_uint32), ("dev", C.c_uint32), ("ret", C.c_int32),
                ("bb_idx", C.c_uint32), ("hash", C.c_uint64)]


class DevInfo(C.Structure):
    _fields_ = [("protocol_num", C.c_uint32), ("modulation", C.c_uint32),
                ("short_width", C.c_float), ("long_width", C.c_float), ("reset_limit", C.c_float),
                ("gap_limit", C.c_float), ("sync_width", C.c_float), ("tolerance", C.c_float),
                ("priority", C.c_uint32), ("disabled", C.c_uint32), ("name", C.c_char * 96)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "name"}
        d["name"] = self.name.decode("utf-8", "replace")
        return d


class BitBuffer(C.Structure):
    _fields_ = [("num_rows", C.c_uint16), ("free_row", C.c_uint16),
                ("bits_per_row", C.c_uint16 * BITBUF_ROWS), ("syncs_before_row", C.c_uint16 * BITBUF_ROWS),
                ("bb", (C.c_uint8 * BITBUF_COLS) * BITBUF_ROWS)]


BITBUFFER_DTYPE = np.dtype([("num_rows", "<u2"), ("free_row", "<u2"), ("bits_per_row", "<u2", (BITBUF_ROWS,)),
                            ("syncs_before_row", "<u2", (BITBUF_ROWS,)), ("bb", "u1", (BITBUF_ROWS, BITBUF_COLS))])
assert BITBUFFER_DTYPE.itemsize == 6604 == C.sizeof(BitBuffer)


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.refh_create.restype = C.c_void_p
        for name, res in [("refh_packages", C.POINTER(Package)), ("refh_events", C.POINTER(Event)),
                          ("refh_bitbuffers", C.c_void_p), ("refh_pulse_pool", C.POINTER(C.c_int32)),
                          ("refh_gap_pool", C.POINTER(C.c_int32)), ("refh_am", C.POINTER(C.c_int16)),
                          ("refh_fm", C.POINTER(C.c_int16)), ("refh_json", C.c_char_p)]:
            getattr(L, name).restype = res
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ["refh_num_packages", "refh_num_events", "refh_num_bitbuffers", "refh_num_stage"]:
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_void_p]
        L.refh_num_decoded.restype = C.c_uint64
        L.refh_num_decoded.argtypes = [C.c_void_p]
        L.refh_destroy.argtypes = [C.c_void_p]
        L.refh_set_capture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.refh_set_timing_mode.argtypes = [C.c_void_p, C.c_int]
        L.refh_set_levels.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
        L.refh_set_fm_low_pass.argtypes = [C.c_void_p, C.c_float]
        L.refh_num_protocols.argtypes = [C.c_void_p]
        L.refh_get_protocol.argtypes = [C.c_void_p, C.c_int, C.POINTER(DevInfo)]
        L.refh_register.argtypes = [C.c_void_p, C.c_int]
        L.refh_register_defaults.argtypes = [C.c_void_p]
        L.refh_register_custom.argtypes = [C.c_void_p, C.c_uint] + [C.c_float] * 6 + [C.c_uint]
        L.refh_num_registered.argtypes = [C.c_void_p]
        L.refh_get_registered.argtypes = [C.c_void_p, C.c_int, C.POINTER(DevInfo)]
        L.refh_run_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint32,
                                      C.c_int, C.c_uint32]
        L.refh_device_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32)]
        L.refh_abi_facts.argtypes = [C.POINTER(C.c_uint32)]
        L.refh_slice.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.refh_slice_all.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.refh_begin_external_dispatch.restype = C.c_void_p
        L.refh_begin_external_dispatch.argtypes = [C.c_void_p]
        L.refh_end_external_dispatch.argtypes = [C.c_void_p]
        L.refh_reset_stats.argtypes = [C.c_void_p]
        L.refh_parse_filename.argtypes = [C.c_char_p, C.POINTER(C.c_uint32)]
        L.refh_envelope_detect.restype = C.c_float
        L.refh_envelope_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.refh_magnitude_est_cu8.restype = C.c_float
        L.refh_magnitude_est_cu8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.refh_magnitude_est_cs16.restype = C.c_float
        L.refh_magnitude_est_cs16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.refh_low_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.refh_demod_fm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_ulong, C.c_uint32, C.c_float, C.c_void_p]
        L.refh_load_ook.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_int]
        L.refh_rfraw.argtypes = [C.c_char_p, C.c_void_p]
        L.refh_dump_ook.restype = C.c_size_t
        L.refh_dump_ook.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.refh_dump_vcd.restype = C.c_size_t
        L.refh_dump_vcd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
        L.refh_dump_raw.argtypes = [C.c_void_p, C.c_uint, C.c_uint64, C.c_void_p, C.c_uint8]
        L.refh_slice_pulse_data.argtypes = [C.c_void_p, C.c_void_p]
        L.refh_sigmf_open.argtypes = [C.c_char_p, C.POINTER(C.c_uint64)]
        L.refh_analyze.restype = C.c_size_t
        L.refh_analyze.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
        _lib = L
    return _lib


def abi_facts():
    out = (C.c_uint32 * 16)()
    lib().refh_abi_facts(out)
    keys = ["sizeof_r_device", "sizeof_bitbuffer", "sizeof_pulse_data", "off_modulation", "off_short_width",
            "off_decode_fn", "off_priority", "off_decode_events", "off_decode_ctx", "off_bits_per_row",
            "off_syncs_before_row", "off_bb", "off_pulse", "off_gap", "off_ook_low_estimate", "off_freq1_hz"]
    return dict(zip(keys, list(out)))


def _pkg_dict(p, pulses, gaps):
    d = {k: getattr(p, k) for k, _ in Package._fields_}
    d["pulse"] = pulses[p.pulse_off:p.pulse_off + p.pulse_count].copy()
    d["gap"] = gaps[p.pulse_off:p.pulse_off + p.pulse_count].copy()
    return d


class Ref:
    """One reference configuration (device set + detector levels); run streams through it."""

    def __init__(self, chain_decoders=False, store_bitbuffers=True, store_stages=False):
        self.L = lib()
        self.h = C.c_void_p(self.L.refh_create())
        self.L.refh_set_capture(self.h, int(chain_decoders), int(store_bitbuffers), int(store_stages))

    def close(self):
        if self.h:
            self.L.refh_destroy(self.h)
            self.h = None

    def protocols(self):
        out = []
        info = DevInfo()
        for i in range(self.L.refh_num_protocols(self.h)):
            self.L.refh_get_protocol(self.h, i, C.byref(info))
            out.append(info.as_dict())
        return out

    def registered(self):
        out = []
        info = DevInfo()
        for i in range(self.L.refh_num_registered(self.h)):
            self.L.refh_get_registered(self.h, i, C.byref(info))
            out.append(info.as_dict())
        return out

    def register(self, protocol_num):
        """protocol_num is the 1-based `-R n` number."""
        r = self.L.refh_register(self.h, protocol_num - 1)
        assert r >= 0
        return r

    def register_defaults(self):
        return self.L.refh_register_defaults(self.h)

    def register_custom(self, modulation, short_width, long_width, reset_limit, gap_limit=0.0, sync_width=0.0,
                        tolerance=0.0, priority=0):
        return self.L.refh_register_custom(self.h, modulation, short_width, long_width, reset_limit, gap_limit,
                                           sync_width, tolerance, priority)

    def set_levels(self, use_mag_est=0, level_limit=0.0, min_level=-12.1442, min_snr=9.0):
        self.L.refh_set_levels(self.h, use_mag_est, level_limit, min_level, min_snr)

    def set_fm_low_pass(self, v):
        self.L.refh_set_fm_low_pass(self.h, v)

    def set_timing_mode(self, on=True):
        self.L.refh_set_timing_mode(self.h, int(on))

    def run_raw(self, iq, sample_size, samp_rate=250000, center_freq=433920000, fpdm=2, block_bytes=0):
        iq = np.ascontiguousarray(iq)
        return self.L.refh_run_stream(self.h, iq.ctypes.data, iq.nbytes, sample_size, samp_rate, center_freq,
                                      fpdm, block_bytes)

    def run(self, iq, sample_size, samp_rate=250000, center_freq=433920000, fpdm=2, block_bytes=0):
        """Returns dict(packages=[...], events=[...], am=..., fm=..., json=[...])."""
        self.run_raw(iq, sample_size, samp_rate, center_freq, fpdm, block_bytes)
        L, h = self.L, self.h
        npk, nev, nbb, nst = (L.refh_num_packages(h), L.refh_num_events(h), L.refh_num_bitbuffers(h),
                              L.refh_num_stage(h))
        pk = L.refh_packages(h)
        npool = sum(pk[i].pulse_count for i in range(npk))
        pulses = np.ctypeslib.as_array(L.refh_pulse_pool(h), (max(npool, 1),))[:npool].copy() if npool else np.zeros(0, np.int32)
        gaps = np.ctypeslib.as_array(L.refh_gap_pool(h), (max(npool, 1),))[:npool].copy() if npool else np.zeros(0, np.int32)
        packages = [_pkg_dict(pk[i], pulses, gaps) for i in range(npk)]
        ev = L.refh_events(h)
        bbs = None
        if nbb:
            addr = L.refh_bitbuffers(h)
            bbs = np.frombuffer((C.c_uint8 * (nbb * 6604)).from_address(addr), dtype=BITBUFFER_DTYPE).copy()
        events = []
        for i in range(nev):
            e = ev[i]
            events.append({"package": e.package, "dev": e.dev, "ret": e.ret, "hash": e.hash,
                           "bitbuffer": bbs[e.bb_idx] if (bbs is not None and e.bb_idx != 0xFFFFFFFF) else None})
        res = {"packages": packages, "events": events,
               "json": [l for l in L.refh_json(h).decode("utf-8", "replace").split("\n") if l],
               "decoded": L.refh_num_decoded(h)}
        if nst:
            res["am"] = np.ctypeslib.as_array(L.refh_am(h), (nst,)).copy()
            res["fm"] = np.ctypeslib.as_array(L.refh_fm(h), (nst,)).copy()
        return res

    def device_stats(self, idx):
        out = (C.c_uint32 * 8)()
        self.L.refh_device_stats(self.h, idx, out)
        return list(out)

    def slice(self, dev_idx, fsk, sample_rate, pulse, gap):
        pulse = np.ascontiguousarray(pulse, np.int32)
        gap = np.ascontiguousarray(gap, np.int32)
        n = self.L.refh_slice(self.h, dev_idx, int(fsk), sample_rate, len(pulse), pulse.ctypes.data, gap.ctypes.data)
        L, h = self.L, self.h
        nbb = L.refh_num_bitbuffers(h)
        if not nbb:
            return []
        addr = L.refh_bitbuffers(h)
        return list(np.frombuffer((C.c_uint8 * (nbb * 6604)).from_address(addr), dtype=BITBUFFER_DTYPE).copy())[:n]


def _slice_all(self, fsk, sample_rate, pulse, gap):
    """All registered devices on one pulse train -> [(dev, bitbuffer record)] in dispatch order."""
    pulse = np.ascontiguousarray(pulse, np.int32)
    gap = np.ascontiguousarray(gap, np.int32)
    L, h = self.L, self.h
    n = L.refh_slice_all(h, int(fsk), sample_rate, len(pulse), pulse.ctypes.data, gap.ctypes.data)
    if not n:
        return []
    ev = L.refh_events(h)
    bbs = np.frombuffer((C.c_uint8 * (n * 6604)).from_address(L.refh_bitbuffers(h)), dtype=BITBUFFER_DTYPE).copy()
    return [(ev[i].dev, bbs[ev[i].bb_idx]) for i in range(n)]


Ref.slice_all = _slice_all


# compound file types of include/fileformat.h that the GPU path understands
FILE_FORMATS = {0x210820: "cu8", 0x210821: "cs8", 0x211021: "cs16", 0x212023: "cf32"}


def parse_filename(name):
    out = (C.c_uint32 * 3)()
    lib().refh_parse_filename(name.encode(), out)
    return {"format_code": out[0], "format": FILE_FORMATS.get(out[0]), "sample_rate": out[1], "center_frequency": out[2]}


def row_hex(bb, row):
    """'{len}hex' code of one bitbuffer row, as rtl_433 prints it."""
    n = int(bb["bits_per_row"][row])
    nbytes = (n + 7) // 8
    data = bytes(bb["bb"][row][:nbytes]) if nbytes <= BITBUF_COLS else bytes(bb["bb"].reshape(-1)[row * BITBUF_COLS:row * BITBUF_COLS + nbytes])
    hx = data.hex()
    return "{%d}%s" % (n, hx[:(n + 3) // 4])


# ---- pulse-level I/O of the reference (src/pulse_data.c, src/rfraw.c) ------------------------------------

PULSE_DATA_DTYPE = np.dtype([("offset", "<u8"), ("sample_rate", "<u4"), ("depth_bits", "<u4"), ("start_ago", "<u4"),
                             ("end_ago", "<u4"), ("num_pulses", "<u4"), ("pulse", "<i4", (1200,)), ("gap", "<i4", (1200,)),
                             ("ook_low_estimate", "<i4"), ("ook_high_estimate", "<i4"), ("fsk_f1_est", "<i4"),
                             ("fsk_f2_est", "<i4"), ("freq1_hz", "<f4"), ("freq2_hz", "<f4"), ("centerfreq_hz", "<f4"),
                             ("range_db", "<f4"), ("rssi_db", "<f4"), ("snr_db", "<f4"), ("noise_db", "<f4")], align=True)
assert PULSE_DATA_DTYPE.itemsize == 9672


def load_ook(text, samp_rate, cap=64):
    """pulse_data_load() until exhausted -> array of pulse_data_t records."""
    if isinstance(text, str):
        text = text.encode()
    out = np.zeros(cap + 1, PULSE_DATA_DTYPE)
    n = lib().refh_load_ook(text, len(text), samp_rate, out.ctypes.data, cap)
    assert n >= 0
    return out[:n].copy()


def rfraw(line):
    """rfraw_check() + rfraw_parse() into a zeroed pulse_data_t -> record or None."""
    out = np.zeros(1, PULSE_DATA_DTYPE)
    if not lib().refh_rfraw(line.encode() if isinstance(line, str) else line, out.ctypes.data):
        return None
    return out[0]


def dump_ook(pd):
    """pulse_data_dump() text (with its ';received <wall clock>' first line)."""
    pd = np.ascontiguousarray(pd)
    buf = C.create_string_buffer(1 << 16)
    n = lib().refh_dump_ook(pd.ctypes.data, buf, len(buf))
    return buf.raw[:n].decode()


def dump_vcd(pd, ch_id="'", header=False):
    pd = np.ascontiguousarray(pd)
    buf = C.create_string_buffer(1 << 17)
    n = lib().refh_dump_vcd(pd.ctypes.data, ord(ch_id), int(header), buf, len(buf))
    return buf.raw[:n].decode()


def dump_raw(pd, length, buf_offset, bits):
    pd = np.ascontiguousarray(pd)
    out = np.zeros(length, np.uint8)
    lib().refh_dump_raw(out.ctypes.data, length, buf_offset, pd.ctypes.data, bits)
    return out


def _slice_pulse_data(self, pd):
    """run_ook_demods / run_fsk_demods of all registered devices on a pulse_data_t record
    -> [(dev, hash, bitbuffer or None)] in dispatch order."""
    pd = np.ascontiguousarray(pd).copy()
    L, h = self.L, self.h
    n = L.refh_slice_pulse_data(h, pd.ctypes.data)
    if not n:
        return []
    ev = L.refh_events(h)
    nbb = L.refh_num_bitbuffers(h)
    bbs = np.frombuffer((C.c_uint8 * (nbb * 6604)).from_address(L.refh_bitbuffers(h)), dtype=BITBUFFER_DTYPE).copy() if nbb else None
    return [(ev[i].dev, ev[i].hash, bbs[ev[i].bb_idx] if bbs is not None and ev[i].bb_idx != 0xffffffff else None) for i in range(n)]


Ref.slice_pulse_data = _slice_pulse_data


def _analyze(self, pd, package_type=1):
    """pulse_analyzer() (src/pulse_analyzer.c:279) on a pulse_data_t record -> (stderr text, [bitbuffer hashes of
    the trial demodulation's events])."""
    pd = np.ascontiguousarray(pd).copy()
    buf = C.create_string_buffer(1 << 18)
    n = self.L.refh_analyze(self.h, pd.ctypes.data, package_type, buf, len(buf))
    ev = self.L.refh_events(self.h)
    return buf.raw[:n].decode(), [ev[i].hash for i in range(self.L.refh_num_events(self.h))]


Ref.analyze = _analyze


def sigmf_open(path):
    """sigmf_reader_open() -> dict(rc, sample_rate, center_frequency, data_offset)."""
    out = (C.c_uint64 * 4)()
    lib().refh_sigmf_open(path.encode(), out)
    rc = out[0] if out[0] < (1 << 63) else out[0] - (1 << 64)
    return {"rc": rc, "sample_rate": out[1], "center_frequency": out[2], "data_offset": out[3]}
