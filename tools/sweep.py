#!/usr/bin/env python
"""BASELINE configs[4]: mixed-rate sweep {250k, 1.024M, 2.048M} x {cu8 OOK, cs16 FSK} x batch size,
inputs resident in HBM, kernels timed by the library's CUDA events; algorithmic GB/s against the
measured HBM peak.  Writes a markdown table to stdout."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtl_433_b200 import lib, synth  # noqa: E402

peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
n = 1 << 18
distinct = 8
devs = lib.default_device_table()
ctx = lib.Context(0)
ctx.set_devices(devs)
print(f"| format | rate | streams x samples | k_front ms | front GB/s (in + out) | frac of {peak:.0f} GB/s | k_detect ms | detect GS/s | detect GB/s (AM in) | frac | k_slice2 ms | packages | events | total GS/s |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for fmt_name, fmt in (("cu8 OOK", lib.FMT_CU8), ("cs16 FSK", lib.FMT_CS16)):
    for rate in (250000, 1024000, 2048000):
        if fmt == lib.FMT_CU8:
            base = [synth.ook_stream(s, n_samples=n, rate=rate, n_bursts=2 if rate < 1000000 else 1,
                                     kinds=None if rate < 1000000 else ("nice",)) for s in range(distinct)]
            freq = 433920000
        else:
            base = [synth.fsk_stream(s, n_samples=n, rate=rate, n_bursts=2).view(np.uint8) for s in range(distinct)]
            freq = 868000000
        per = base[0].nbytes
        small = torch.from_numpy(np.concatenate(base)).cuda()
        for batch in (256, 1024, 4096, 16384, 32768):
            if batch * per > 40e9:
                continue
            dev = small.repeat((batch + distinct - 1) // distinct)[: batch * per].contiguous()
            offsets = np.arange(batch + 1, dtype=np.uint64) * np.uint64(per)
            best = None
            for it in range(3):
                ctx.process(dev.data_ptr(), offsets, fmt, rate, freq, data_on_device=True)
                tm = ctx.timing()
                if best is None or tm["front_ms"] + tm["detect_ms"] + tm["slice_ms"] < best["front_ms"] + best["detect_ms"] + best["slice_ms"]:
                    best = tm
            c = ctx.counts()
            gsps = c["samples"] / best["detect_ms"] / 1e6
            gbps = gsps * (2 + 1 / 16)           # k_detect reads the 16-bit AM and the chunk bounds
            fgb = c["samples"] * (fmt + 2 + 1 / 16) / best["front_ms"] / 1e6  # k_front: IQ in, AM + chunk bounds out
            tot = c["samples"] / (best["front_ms"] + best["detect_ms"] + best["slice_ms"]) / 1e6
            print(f"| {fmt_name} | {rate / 1e3:.0f}k | {batch} x 2^18 | {best['front_ms']:.2f} | {fgb:.0f} | {fgb / peak:.3f} | {best['detect_ms']:.2f} | {gsps:.1f} | {gbps:.1f} | {gbps / peak:.4f} | "
                  f"{best['slice_ms']:.2f} | {c['packages']} | {c['events']} | {tot:.1f} |", flush=True)
            del dev
        del small
