"""Capture-file ingest: the part of `rtl_433 -r FILE ...` in front of the hot path.

* `parse_capture_name()` restates file_info_parse_filename() / file_type() (src/fileformat.c:173-328):
  sample rate, centre frequency and sample format are read from tags in the file NAME
  ("g001_433.92M_250k.cu8", "am:s16:path", "868M_1024k.cs16", ...).
* `load_batches()` groups files by (format, rate, frequency) -- one r433b batch each -- and packs
  them at 16-byte aligned starts with their true lengths (r433b_batch.lengths).
* `python -m rtl_433_b200.captures FILES...` replays capture files through the GPU path and
  prints, per file, the detected packages and the bitbuffer rows every requested device's slicer
  produced, in rtl_433's "{len}hex" notation (what `rtl_433 -R n:vv` logs before decoding).

Decoding itself stays with the reference's decoders; this module stops at the
bitbuffer like the rest of the package.
"""
import argparse
import os
import re

import numpy as np

from . import lib

DEFAULT_RATE = 250000        # include/rtl_433.h:13
DEFAULT_FREQ = 433920000     # include/rtl_433.h:14

# format tags of file_type(), src/fileformat.c:222-252 (only what the hot path can take is mapped)
_FORMAT_TAGS = {"cu8": "cu8", "data": "cu8", "complex16u": "cu8", "cs8": "cs8", "complex16s": "cs8", "cs16": "cs16",
                "cf32": "cf32", "cfile": "cf32", "complex": "cf32", "s16": "s16", "u8": "u8", "s8": "s8", "u16": "u16",
                "u32": "u32", "s32": "s32", "f32": "f32", "cs32": "cs32"}
_CONTENT_TAGS = {"i", "q", "iq", "am", "fm", "vcd", "ook", "logic", "sigmf"}


def _scan(text, info):
    """One pass of file_type() over `text`, updating info in place (later tags win)."""
    p, n = 0, len(text)
    while p < n:
        ch = text[p]
        if ch.isdigit() and ch.isascii():
            start = p
            while p < n and text[p] in "0123456789":
                p += 1
            if p < n and text[p] == ".":
                p += 1
                if not (p < n and text[p] in "0123456789"):
                    continue  # "if not [0-9] after '.' abort": the number is dropped
                while p < n and text[p] in "0123456789":
                    p += 1
            s = p
            while p < n and text[p].isascii() and text[p].isalpha():
                p += 1
            num = float(text[start:s])
            unit = text[s:p]
            scale = {"k": 1e3, "m": 1e6, "g": 1e9}.get(unit[:1].lower(), 1.0) if unit else 1.0
            low = unit.lower()
            if low == "m":
                info["center_frequency"] = int(num * 1e6)
            elif low == "k":
                info["sample_rate"] = int(num * 1e3)
            elif low == "hz":
                info["center_frequency"] = int(num)
            elif low == "sps":
                info["sample_rate"] = int(num)
            elif len(unit) == 3 and low[1:] == "hz" and scale > 1.0:
                info["center_frequency"] = int(num * scale)
            elif len(unit) == 4 and low[1:] == "sps" and scale > 1.0:
                info["sample_rate"] = int(num * scale)
        elif ch.isascii() and ch.isalpha():
            start = p
            while p < n and text[p].isascii() and text[p].isalnum():
                p += 1
            tag = text[start:p].lower()
            if tag in _FORMAT_TAGS:
                info["format"] = _FORMAT_TAGS[tag]
            elif tag in _CONTENT_TAGS:
                info["content"] = tag
        else:
            p += 1


def parse_capture_name(spec):
    """-> dict(path, format, content, sample_rate, center_frequency); 0 = not given in the name.
    A prefix up to the last ':' (not followed by a backslash) is an override, parsed last."""
    info = {"format": None, "content": None, "sample_rate": 0, "center_frequency": 0}
    cut = None
    for m in re.finditer(":", spec):
        if spec[m.start() + 1:m.start() + 2] == "\\":
            break
        cut = m.start()
    if cut is not None and cut < 64:
        path = spec[cut + 1:]
        _scan(path, info)
        _scan(spec[:cut], info)
    else:
        path = spec
        _scan(spec, info)
    # file_type_guess_auto_format(): nothing (or just "iq") means cu8
    if info["format"] is None and info["content"] in (None, "iq"):
        info["format"] = "cu8"
    info["path"] = path
    return info


def read_sigmf(path):
    """A SigMF archive as `rtl_433 -r x.sigmf` reads it (sigmf_reader_open(), src/sigmf.c:336-434, and the file loop
    of src/rtl_433.c:1712-1723) -> dict(data, sample_rate, center_frequency, datatype, data_offset).

    The first `.sigmf-meta` member names the stream; `global."core:sample_rate"` and the LAST capture's
    `"core:frequency"` (src/sigmf.c:127-287 overwrites first_frequency per capture) become rate and centre
    frequency; the samples are the member named like the meta file with `-data`.  Two properties of the reference
    are kept because results depend on them: the payload is demodulated as cu8 whatever "core:datatype" says
    (src/rtl_433.c:1719), and the block loop reads from the start of the data member to the END OF THE ARCHIVE,
    i.e. including the tar padding and end-of-archive blocks behind the samples."""
    import json

    def members(raw):
        # ustar walk like microtar's: 512-byte headers, octal size at 124, type flag at 156, all-zero block ends
        pos = 0
        while pos + 512 <= len(raw):
            h = raw[pos:pos + 512]
            if h[0] == 0:
                break
            name = h[:100].split(b"\0", 1)[0].decode("latin-1")
            size = int(h[124:136].split(b"\0", 1)[0].strip() or b"0", 8)
            kind = h[156:157]
            yield name, size, kind, pos + 512
            pos += 512 + (size + 511) // 512 * 512

    raw = np.fromfile(path, dtype=np.uint8).tobytes()
    stream, meta = None, None
    for name, size, kind, at in members(raw):
        if kind not in (b"0", b"\0"):
            continue
        if name.lower().endswith(".sigmf-meta"):
            if stream is not None and name != stream:  # "updated meta file": the later stream name wins
                stream, meta = None, None
            if stream is None:
                stream, meta = name, json.loads(raw[at:at + size].decode("utf-8", "replace") or "{}")
    if stream is None:
        raise ValueError(f"{path}: SigMF input file with no streams")
    info = {"sample_rate": 0, "center_frequency": 0, "datatype": None}
    g = meta.get("global", {}) if isinstance(meta, dict) else {}
    if "core:sample_rate" in g:
        info["sample_rate"] = int(float(g["core:sample_rate"])) & 0xffffffff
    info["datatype"] = g.get("core:datatype")
    for cap in meta.get("captures", []) if isinstance(meta, dict) else []:
        if "core:frequency" in cap:
            info["center_frequency"] = int(float(cap["core:frequency"])) & 0xffffffff
    want = stream[:-4] + "data"
    for name, size, kind, at in members(raw):
        if name == want:
            info["data_offset"] = at
            info["data"] = np.frombuffer(raw, dtype=np.uint8)[at:].copy()
            return info
    raise ValueError(f"{path}: SigMF input file with no stream data")


_ABI_FORMAT = {"cu8": lib.FMT_CU8, "cs8": lib.FMT_CS8, "cs16": lib.FMT_CS16, "cf32": lib.FMT_CF32}


def load_batches(specs, default_rate=DEFAULT_RATE, default_freq=DEFAULT_FREQ, uniform="auto"):
    """-> list of dict(format, sample_rate, center_frequency, files, data, offsets, lengths), one per
    (format, rate, frequency) group, files in command-line order inside a group.

    `uniform`: put the files of a group on ONE stride (the longest file, rounded up) so that r433b_process() can
    overlap the host-to-device copy with the kernels, time slice by time slice (one strided copy per slice);
    "auto" does it when the padding costs less than half again the bytes, False packs the files back to back."""
    groups = {}
    preloaded = {}
    for spec in specs:
        info = parse_capture_name(spec)
        if info["content"] == "sigmf":
            # container: format, rate and frequency come from the archive's metadata, not from the name
            sm = read_sigmf(info["path"])
            preloaded[info["path"]] = sm["data"]
            groups.setdefault(("cu8", sm["sample_rate"], sm["center_frequency"]), []).append(info["path"])
            continue
        if info["format"] not in _ABI_FORMAT:
            raise ValueError(f"{spec}: format {info['format']!r} is not on the GPU path (cu8, cs8, cs16, cf32 are)")
        key = (info["format"], info["sample_rate"] or default_rate, info["center_frequency"] or default_freq)
        groups.setdefault(key, []).append(info["path"])
    out = []
    for (fmt, rate, freq), paths in groups.items():
        ss = {"cu8": 2, "cs8": 2, "cs16": 4, "cf32": 8}[fmt]
        align = 32 if fmt == "cf32" else 16
        bufs = [preloaded[p] if p in preloaded else np.fromfile(p, dtype=np.uint8) for p in paths]
        lengths = np.array([len(b) // ss * ss for b in bufs], np.uint64)  # a trailing partial sample is dropped
        offsets = np.zeros(len(bufs) + 1, np.uint64)
        longest = (int(lengths.max()) + 4095) // 4096 * 4096 if len(bufs) else 0
        if uniform is True or (uniform == "auto" and len(bufs) > 1 and longest * len(bufs) <= 1.5 * float(lengths.sum())):
            offsets = np.arange(len(bufs) + 1, dtype=np.uint64) * np.uint64(longest)
        else:
            for i, n in enumerate(lengths):
                offsets[i + 1] = offsets[i] + (int(n) + align - 1) // align * align
        data = np.zeros(int(offsets[-1]), np.uint8)
        for i, b in enumerate(bufs):
            data[int(offsets[i]):int(offsets[i]) + int(lengths[i])] = b[:int(lengths[i])]
        out.append({"format": fmt, "abi_format": _ABI_FORMAT[fmt], "sample_rate": rate, "center_frequency": freq,
                    "files": paths, "data": data, "offsets": offsets, "lengths": lengths})
    return out


def row_code(bb, row):
    """rtl_433's '{len}hex' row notation (src/decoder_util.c:61-90) of a re-inflated bitbuffer record."""
    n = int(bb["bits_per_row"][row])
    nbytes = (n + 7) // 8
    flat = bb["bb"].reshape(-1)
    data = bytes(flat[row * 128:row * 128 + nbytes])
    return "{%d}%s" % (n, data.hex()[:(n + 3) // 4])


def replay(specs, protocols=None, cuda_device=0, max_rows=8, out=print):
    """Run capture files through the GPU path; report packages and slicer output per file."""
    table = lib.default_device_table(include_disabled=True)
    if protocols:
        devs = [d for d in table if d["protocol_num"] in set(protocols)]
    else:
        devs = [d for d in table if d["disabled"] == 0]
    ctx = lib.Context(cuda_device)
    ctx.set_devices(devs)
    summary = []
    try:
        for batch in load_batches(specs):
            ctx.process(batch["data"], batch["offsets"], batch["abi_format"], batch["sample_rate"],
                        batch["center_frequency"], lengths=batch["lengths"])
            res = ctx.fetch()
            for i, path in enumerate(batch["files"]):
                pk = res["packages"][res["packages"]["stream"] == i]
                out(f"{path}: {batch['format']} {batch['sample_rate']} S/s {batch['center_frequency']} Hz, "
                    f"{int(batch['lengths'][i]) // {lib.FMT_CU8: 2, lib.FMT_CS8: 2, lib.FMT_CS16: 4, lib.FMT_CF32: 8}[batch['abi_format']]} samples, {len(pk)} package(s)")
                events = []

                def on_event(pkg, dev, pd, bb, events=events):
                    events.append((pkg, dev, bb.copy()))
                    return 0

                ctx.dispatch(i, on_event)
                for k in pk:
                    kind = "OOK" if k["type"] == lib.PACKAGE_OOK else "FSK"
                    out(f"  {kind} package @{int(k['offset'])}: {int(k['num_pulses'])} pulses, "
                        f"levels low {int(k['ook_low_estimate'])} high {int(k['ook_high_estimate'])}")
                shown = 0
                for pkg, dev, bb in events:
                    rows = [row_code(bb, r) for r in range(min(int(bb["num_rows"]), max_rows))]
                    if protocols or any(int(b) > 16 for b in bb["bits_per_row"][:int(bb["num_rows"])]):
                        if shown < 40:
                            out(f"    [{devs[dev]['protocol_num']}] {devs[dev]['name']}: " + " ".join(rows))
                        shown += 1
                summary.append({"file": path, "packages": len(pk), "events": len(events)})
    finally:
        ctx.close()
    return summary


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("files", nargs="+", help="capture files; rate/frequency/format come from the name as in rtl_433 -r")
    ap.add_argument("-R", dest="protocols", type=int, action="append", help="protocol number(s); default: all enabled")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    replay(a.files, a.protocols, a.device)


if __name__ == "__main__":
    main()
