#!/usr/bin/env python
"""bench.py -- IQ MS/s through the demod + pulse-detect + slice hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path, all host cores

One "step" = one pass of the hot path over one batch of synthetic captures.  Workload at N=1 is
BASELINE.json configs[1]: 4096 batched 250 kS/s cu8 streams of 2^20 samples, OOK envelope ->
pulse_detect -> every slicer of the 335 default-enabled r_devices.  Under torchrun every rank
owns its own 4096 streams (independent capture files, no exchange step: weak scaling); NCCL is
used only to agree on the slowest rank's time and to sum the sample counters.

`value`     : inputs already resident in HBM, device-timed (CUDA events), max over ranks.
`e2e`       : the same batch through the C ABI with HOST buffers: pinned host IQ -> H2D ->
              kernels -> D2H of the compact packages/events, every step.
`roofline`  : k_detect (the dominant kernel): algorithmic bytes (2 B per cu8 IQ sample read +
              package records written) / its CUDA-event duration, against the measured HBM peak.
`cpu_baseline`: the unmodified reference (oracle/_ref) on a bounded sample of the same streams,
              all host cores, decoders stubbed so both arms end at the bitbuffer.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (format bytes/sample, rate, centre frequency, log2 samples per stream, streams per GPU)
    "ook_cu8_250k": dict(fmt=2, rate=250000, freq=433920000, log2n=20, streams=4096, kind="ook",
                         desc="BASELINE configs[1]: 4096 x 2^20-sample 250 kS/s cu8 streams, OOK+FSK detect, "
                              "all slicers of the 335 default r_devices"),
    "fsk_cs16_1024k": dict(fmt=4, rate=1024000, freq=868000000, log2n=20, streams=1024, kind="fsk",
                           desc="BASELINE configs[2]: 1024 x 2^20-sample 1.024 MS/s cs16 2-FSK streams, minmax detector"),
}


def _gen_one(args):
    kind, seed, n = args
    from rtl_433_b200 import synth
    if kind == "ook":
        return synth.ook_stream(seed, n_samples=n)
    return synth.fsk_stream(seed, n_samples=n).view(np.uint8)


def generate(kind, seeds, n):
    jobs = [(kind, s, n) for s in seeds]
    procs = min(len(jobs), max(1, (os.cpu_count() or 2) // 2), 32)
    if procs > 1:
        with mp.get_context("spawn").Pool(procs) as pool:
            return pool.map(_gen_one, jobs)
    return [_gen_one(j) for j in jobs]


# ------------------------------------------------------------------ reference / CPU arm -------

def _ref_worker(args):
    """One host core: the unmodified reference flow over `streams` (decoders stubbed out)."""
    kind, fmt, rate, freq, seeds, n, repeats = args
    from oracle import refh
    r = refh.Ref(store_bitbuffers=False)
    r.register_defaults()
    r.set_timing_mode(True)
    data = [_gen_one((kind, s, n)) for s in seeds]
    t = time.perf_counter()
    samples = 0
    for _ in range(repeats):
        for x in data:
            r.run_raw(x, fmt, rate, freq, 2)
            samples += x.nbytes // fmt
    return samples, time.perf_counter() - t


REF_DISTINCT_PER_CORE = 8


def reference_pass(w, cores, streams_per_core, seed0=0):
    """Every host core runs the unmodified reference over `streams_per_core` stream passes
    (REF_DISTINCT_PER_CORE seeded streams, repeated; detector state is reset per stream as for
    separate files).  -> (aggregate MS/s, samples, slowest core's processing seconds, wall)."""
    n = 1 << w["log2n"]
    distinct = min(REF_DISTINCT_PER_CORE, streams_per_core)
    repeats = max(1, streams_per_core // distinct)
    jobs = [(w["kind"], w["fmt"], w["rate"], w["freq"], [seed0 + c * distinct + i for i in range(distinct)], n, repeats)
            for c in range(cores)]
    t = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t
    samples = sum(r[0] for r in res)
    busy = max(r[1] for r in res)  # generation excluded: the slowest core's processing time
    return samples / busy / 1e6, samples, busy, wall


def run_reference(a, w):
    from oracle import refh
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not refh.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libr433ref.so was not built/shipped"}))
        return
    cores = os.cpu_count() or 1
    spc = a.ref_streams_per_core
    vals = []
    for i in range(min(a.warmup, 1)):
        reference_pass(w, cores, 1)
    tot_samples, tot_busy = 0, 0.0
    for i in range(a.steps):
        v, samples, busy, _ = reference_pass(w, cores, spc)
        vals.append(v)
        tot_samples += samples
        tot_busy += busy
    value = tot_samples / tot_busy / 1e6
    sample = f"{cores} processes x {spc} streams x 2^{w['log2n']} samples per step (seeded like the GPU arm), decoders stubbed"
    line = {
        "impl": "reference", "metric": "IQ MS/s end-to-end demod+slice", "value": round(value, 2), "unit": "MS/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(tot_busy / a.steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32 fixed point", "data": "synthetic",
        "config": {"workload": a.workload, "desc": w["desc"], "streams_per_step": cores * spc, "samples_per_stream": 1 << w["log2n"]},
        "cpu_baseline": {"value": round(value, 2), "unit": "MS/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------- GPU arm ---------

class ClockSampler(threading.Thread):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        busy = [v for v in sm if v > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    p = os.path.join(ROOT, "profiles", "detect_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def run_gpu(a, w):
    import torch
    import torch.distributed as dist
    from rtl_433_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    fmt, n = w["fmt"], 1 << w["log2n"]
    streams = a.streams or w["streams"]
    distinct = min(a.distinct, streams)
    # file i of the whole job goes to rank i mod world (round-robin shard, BASELINE configs[3]);
    # file i carries the synthetic capture with seed i mod (distinct * world)
    from rtl_433_b200 import shard
    my_files = shard.files_for_rank(streams * world, rank, world)
    seeds = sorted({f % (distinct * world) for f in my_files})
    base_by_seed = dict(zip(seeds, generate(w["kind"], seeds, n)))
    base = [base_by_seed[f % (distinct * world)] for f in my_files[:distinct]]
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per = base[0].nbytes
    reps = (streams + distinct - 1) // distinct
    host = torch.empty(streams * per, dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    for i in range(streams):
        hv[i * per:(i + 1) * per] = base[i % distinct]
    dev = host.cuda(non_blocking=True)
    torch.cuda.synchronize()
    offsets = np.arange(streams + 1, dtype=np.uint64) * np.uint64(per)
    devs = lib.default_device_table()
    ctx = lib.Context(local)
    ctx.set_devices(devs)
    ctx.set_pipeline(a.pipeline)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps (value, roofline) ------------------------------------------
    for _ in range(a.warmup):
        ctx.process(dev.data_ptr(), offsets, fmt, w["rate"], w["freq"], data_on_device=True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    detect_ms, slice_ms, launches = [], [], 0
    e0.record()
    for _ in range(a.steps):
        ctx.process(dev.data_ptr(), offsets, fmt, w["rate"], w["freq"], data_on_device=True)
        tm = ctx.timing()
        detect_ms.append(tm["detect_ms"])
        slice_ms.append(tm["slice_ms"])
        launches += tm["detect_launches"] + tm["slice_launches"]
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    counts = ctx.counts()

    # ---- end to end through the C ABI with host buffers -------------------------------------
    res = None
    for _ in range(min(a.warmup, 1)):
        ctx.process(hv, offsets, fmt, w["rate"], w["freq"])
        res = ctx.fetch()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        ctx.process(hv, offsets, fmt, w["rate"], w["freq"])
        res = ctx.fetch()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    tm_e2e = ctx.timing()
    clocks = sampler.stop()
    d2h_bytes = int(res["n_packages"] * 72 + 2 * res["pulse_pool"].nbytes + res["pairs"].nbytes + res["event_bytes"])

    samples_step = streams * n
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(samples_step)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max = t.tolist()
    total_samples_step = tot.item()

    if rank == 0:
        value = total_samples_step * a.steps / (dev_ms_max * 1e-3) / 1e6
        e2e_value = total_samples_step * a.e2e_steps / (e2e_ms_max * 1e-3) / 1e6
        peak, peak_src = measured_peak()
        det = float(np.mean(detect_ms))
        pkg_bytes = counts["packages"] * 72 + 8 * (counts["packages"] * 140)  # headers + ~pulse/gap widths
        alg_bytes = samples_step * fmt + pkg_bytes
        achieved = alg_bytes / (det * 1e-3) / 1e9
        # the recorded ncu capture is of the default OOK workload only
        traffic = recorded_traffic() if (a.workload == "ook_cu8_250k" and streams == w["streams"]) else None
        line = {
            "metric": "IQ MS/s end-to-end demod+slice", "value": round(value, 1), "unit": "MS/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(dev_ms_max / a.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16/int32 fixed point" if fmt == 2 else "int16/int32/int64 fixed point",
            "data": "synthetic",
            "config": {"workload": a.workload, "desc": w["desc"], "streams_per_gpu": streams, "samples_per_stream": n,
                       "distinct_streams_per_gpu": distinct, "replication": f"{distinct} seeded streams tiled {reps}x per GPU",
                       "devices": len(devs), "sharding": f"round-robin files over {world} ranks, no data-path collective",
                       "l2": f"inputs {streams * per / 2**20:.0f} MiB per GPU, larger than L2 (126 MB): no flush needed",
                       "packages_per_step": counts["packages"], "events_per_step": counts["events"]},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "MS/s", "h2d_bytes_per_step": int(streams * per),
                    "d2h_bytes_per_step": d2h_bytes, "steps": a.e2e_steps,
                    "breakdown_ms": {k: round(tm_e2e[k], 2) for k in ("h2d_ms", "detect_ms", "slice_ms", "d2h_ms")}},
            "gpu_launches": launches,
            "kernel_ms": {"k_detect": round(det, 3), "k_slice": round(float(np.mean(slice_ms)), 3)},
            "roofline": {"bound": "hbm", "kernel": "k_detect", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "peak_source": peak_src,
                         "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                         "algorithmic_bytes_per_launch": int(alg_bytes)},
        }
        if world == 1 and not a.no_cpu_baseline:
            from oracle import refh
            if refh.available():
                cores = os.cpu_count() or 1
                spc = a.ref_streams_per_core
                v, s, busy, wall = reference_pass(w, cores, spc)
                line["cpu_baseline"] = {"value": round(v, 2), "unit": "MS/s", "cores": cores, "kind": "reference",
                                        "sample": f"{cores} processes x {spc} streams x 2^{w['log2n']} samples of the same workload, "
                                                  f"{busy:.1f} s, decoders stubbed (ends at the bitbuffer like the GPU arm)"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "MS/s", "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref not shipped"}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ook_cu8_250k", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (default: the workload's)")
    ap.add_argument("--distinct", type=int, default=256, help="distinct seeded streams per GPU, tiled to --streams")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-streams-per-core", type=int, default=96,
                    help="stream passes per host core in the reference/cpu_baseline leg (~22 ms each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline", type=int, default=0, help="host-input pipeline groups (0 auto, 1 off)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 0)
    w = WORKLOADS[a.workload]
    if a.impl == "reference":
        run_reference(a, w)
    else:
        run_gpu(a, w)


if __name__ == "__main__":
    main()
