#!/usr/bin/env python
"""e2e (host buffers) time vs pipeline group count."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtl_433_b200 import lib, synth
streams, distinct, n = 4096, 32, 1 << 20
base = [synth.ook_stream(s, n_samples=n) for s in range(distinct)]
per = base[0].nbytes
host = torch.empty(streams * per, dtype=torch.uint8, pin_memory=True)
hv = host.numpy()
for i in range(streams):
    hv[i * per:(i + 1) * per] = base[i % distinct]
offsets = np.arange(streams + 1, dtype=np.uint64) * np.uint64(per)
ctx = lib.Context(0)
ctx.set_devices(lib.default_device_table())
for G in [8, 12, 16]:
    ctx.set_pipeline(G)
    for it in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        ctx.process(hv, offsets, lib.FMT_CU8, 250000, 433920000)
        t1 = time.perf_counter()
        ctx.fetch()
        t2 = time.perf_counter()
        tm = ctx.timing()
    print(f"G={G}: process {1e3*(t1-t):.1f} ms fetch {1e3*(t2-t1):.1f} ms total {1e3*(t2-t):.1f} ms -> {streams*n/(t2-t)/1e6:.0f} MS/s | kernels: detect {tm['detect_ms']:.1f} slice {tm['slice_ms']:.1f}")
