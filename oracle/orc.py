"""ctypes view of oracle/libr433oracle.so (the CPU restatement) -- TEST INFRASTRUCTURE ONLY.

Never imported by the product (rtl_433_b200/); see oracle/r433_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .refh import BITBUFFER_DTYPE, Event, Package, _pkg_dict

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libr433oracle.so")


class Device(C.Structure):
    _fields_ = [("modulation", C.c_uint32), ("short_width", C.c_float), ("long_width", C.c_float),
                ("reset_limit", C.c_float), ("gap_limit", C.c_float), ("sync_width", C.c_float),
                ("tolerance", C.c_float), ("priority", C.c_uint32)]


def build(force=False):
    src = os.path.join(HERE, "r433_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_capture.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_set_levels.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_set_fm_low_pass.argtypes = [C.c_void_p, C.c_float]
        L.orc_add_device.argtypes = [C.c_void_p, C.POINTER(Device)]
        L.orc_num_devices.argtypes = [C.c_void_p]
        L.orc_cf32_to_cs16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_run_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint32, C.c_int,
                                     C.c_uint32]
        for name, res in [("orc_packages", C.POINTER(Package)), ("orc_events", C.POINTER(Event)),
                          ("orc_bitbuffers", C.c_void_p), ("orc_pulse_pool", C.POINTER(C.c_int32)),
                          ("orc_gap_pool", C.POINTER(C.c_int32)), ("orc_am", C.POINTER(C.c_int16)),
                          ("orc_fm", C.POINTER(C.c_int16))]:
            getattr(L, name).restype = res
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ["orc_num_packages", "orc_num_events", "orc_num_bitbuffers", "orc_num_stage"]:
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_envelope_cu8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_magnitude_cu8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_magnitude_cs16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_low_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_fm_coeffs.argtypes = [C.c_int, C.c_uint32, C.c_float, C.c_void_p]
        L.orc_demod_fm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_ulong, C.c_uint32, C.c_float, C.c_void_p]
        L.orc_detector_levels.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_slice.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class Oracle:
    """Same surface as oracle.refh.Ref so tests can drive either."""

    def __init__(self, store_bitbuffers=True, store_stages=False):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())
        self.L.orc_set_capture(self.h, int(store_bitbuffers), int(store_stages))

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def add_device(self, modulation, short_width, long_width, reset_limit, gap_limit=0.0, sync_width=0.0,
                   tolerance=0.0, priority=0, **_):
        d = Device(modulation, short_width, long_width, reset_limit, gap_limit, sync_width, tolerance, priority)
        return self.L.orc_add_device(self.h, C.byref(d))

    def add_devices(self, devs):
        for d in devs:
            self.add_device(**{k: d[k] for k in ("modulation", "short_width", "long_width", "reset_limit",
                                                  "gap_limit", "sync_width", "tolerance", "priority")})

    def set_levels(self, use_mag_est=0, level_limit=0.0, min_level=-12.1442, min_snr=9.0):
        self.L.orc_set_levels(self.h, use_mag_est, level_limit, min_level, min_snr)

    def set_fm_low_pass(self, v):
        self.L.orc_set_fm_low_pass(self.h, v)

    def cf32_to_cs16(self, floats):
        """src/rtl_433.c:1811-1825 on a float32 array -> int16 array of the same length."""
        x = np.ascontiguousarray(floats, np.float32)
        out = np.empty(len(x), np.int16)
        self.L.orc_cf32_to_cs16(x.ctypes.data, out.ctypes.data, len(x))
        return out

    def run_raw(self, iq, sample_size, samp_rate=250000, center_freq=433920000, fpdm=2, block_bytes=0):
        iq = np.ascontiguousarray(iq)
        return self.L.orc_run_stream(self.h, iq.ctypes.data, iq.nbytes, sample_size, samp_rate, center_freq, fpdm,
                                     block_bytes)

    def run(self, iq, sample_size, samp_rate=250000, center_freq=433920000, fpdm=2, block_bytes=0):
        self.run_raw(iq, sample_size, samp_rate, center_freq, fpdm, block_bytes)
        L, h = self.L, self.h
        npk, nev, nbb, nst = L.orc_num_packages(h), L.orc_num_events(h), L.orc_num_bitbuffers(h), L.orc_num_stage(h)
        pk = L.orc_packages(h)
        npool = sum(pk[i].pulse_count for i in range(npk))
        if npool:
            pulses = np.ctypeslib.as_array(L.orc_pulse_pool(h), (npool,)).copy()
            gaps = np.ctypeslib.as_array(L.orc_gap_pool(h), (npool,)).copy()
        else:
            pulses = gaps = np.zeros(0, np.int32)
        packages = [_pkg_dict(pk[i], pulses, gaps) for i in range(npk)]
        ev = L.orc_events(h)
        bbs = None
        if nbb:
            addr = L.orc_bitbuffers(h)
            bbs = np.frombuffer((C.c_uint8 * (nbb * 6604)).from_address(addr), dtype=BITBUFFER_DTYPE).copy()
        events = [{"package": ev[i].package, "dev": ev[i].dev, "ret": ev[i].ret, "hash": ev[i].hash,
                   "bitbuffer": bbs[ev[i].bb_idx] if (bbs is not None and ev[i].bb_idx != 0xFFFFFFFF) else None}
                  for i in range(nev)]
        res = {"packages": packages, "events": events}
        if nst:
            res["am"] = np.ctypeslib.as_array(L.orc_am(h), (nst,)).copy()
            res["fm"] = np.ctypeslib.as_array(L.orc_fm(h), (nst,)).copy()
        return res

    def slice(self, dev_idx, sample_rate, pulse, gap):
        pulse = np.ascontiguousarray(pulse, np.int32)
        gap = np.ascontiguousarray(gap, np.int32)
        n = self.L.orc_slice(self.h, dev_idx, sample_rate, len(pulse), pulse.ctypes.data, gap.ctypes.data)
        nbb = self.L.orc_num_bitbuffers(self.h)
        if not nbb:
            return []
        addr = self.L.orc_bitbuffers(self.h)
        return list(np.frombuffer((C.c_uint8 * (nbb * 6604)).from_address(addr), dtype=BITBUFFER_DTYPE).copy())[:n]
