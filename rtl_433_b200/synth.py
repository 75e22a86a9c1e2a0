"""Seeded synthetic IQ captures for parity tests and the benchmark.

The OOK generator is modelled on the reference's only in-tree IQ test vector,
tests/rtl_tcp_serve.py:46-71 (an on/off keyed tone), plus a Gaussian noise floor so the
integer IIR never sits on one of its constant-input fixed points.  Everything here is
input data; nothing in this file is on the measured path.
"""
import numpy as np

OOK_RATE = 250000
FSK_RATE = 1024000

# --- burst shapes: lists of (duration_us, on) ------------------------------------------


def _pwm_frame(bits, short_us, long_us, period_us=None, gap_us=None):
    seg = []
    for b in bits:
        w = short_us if b else long_us
        seg.append((w, 1))
        seg.append(((period_us - w) if period_us else gap_us, 0))
    return seg


def silvercrest_burst(rng=None, payload=None):
    """PWM 264/744 us, 1000 us period, 33 bits x 4 rows, 6000 us between rows
    (src/devices/silvercrest.c:56-65).  Default payload decodes: 7c 26 00 02 + '0'."""
    if payload is None:
        if rng is None:
            payload = [int(c) for c in "".join(f"{b:08b}" for b in (0x7C, 0x26, 0x00, 0x02)) + "0"]
        else:
            payload = list(rng.integers(0, 2, 33))
    seg = []
    for r in range(4):
        row = _pwm_frame(payload, 264, 744, period_us=1000)
        row[-1] = (6000 if r < 3 else 0, 0)
        seg += row
    return seg


def nice_flor_s_burst(rng=None, code="e7a760b94372e"):
    """PWM 500/1000 us with a 1500 us sync, 52 bits (src/devices/nice_flor_s.c:134,
    tests/http-rtltcp-test.sh:31-36)."""
    if rng is None:
        bits = [int(c) for c in "".join(f"{int(ch, 16):04b}" for ch in code)]
    else:
        bits = list(rng.integers(0, 2, 52))
    seg = _pwm_frame(bits, 500, 1000, gap_us=500)
    seg.append((1500, 1))
    seg.append((0, 0))
    return seg


def nexus_burst(rng):
    """PPM: 500 us pulse, 1000/2000 us gaps, 36 bits x 6 rows, 4000 us row gap
    (src/devices/nexus.c timing)."""
    bits = list(rng.integers(0, 2, 36))
    seg = []
    for r in range(6):
        for b in bits:
            seg.append((500, 1))
            seg.append((2000 if b else 1000, 0))
        seg.append((500, 1))
        seg.append((4000 if r < 5 else 0, 0))
    return seg


def manchester_burst(rng, half_us=500, nbits=64):
    """Manchester coded (Oregon-shaped), half-bit 500 us: 1 -> on,off ; 0 -> off,on."""
    bits = [1, 1, 1, 1] + list(rng.integers(0, 2, nbits))
    level = []
    for b in bits:
        level += [1, 0] if b else [0, 1]
    seg = []
    cur, width = level[0], 0
    for lv in level:
        if lv == cur:
            width += half_us
        else:
            seg.append((width, cur))
            cur, width = lv, half_us
    seg.append((width, cur))
    if seg[0][1] == 0:
        seg = seg[1:]
    if seg[-1][1] == 1:
        seg.append((0, 0))
    return seg


def _render_ook(seg, rate):
    """-> on/off mask (uint8) for a segment list."""
    spu = rate / 1e6
    parts = [np.full(int(round(us * spu)), on, np.uint8) for us, on in seg]
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def ook_stream(seed, n_samples=1 << 20, rate=OOK_RATE, n_bursts=8, sigma=2.0, kinds=None, decodable=False):
    """cu8 IQ (uint8[2*n_samples]) with `n_bursts` OOK bursts over a sigma-LSB noise floor."""
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal(2 * n_samples, dtype=np.float32)
    x = noise * np.float32(sigma) + np.float32(127.5)
    kinds = kinds or ("silvercrest", "nice", "nexus", "manchester")
    masks = []
    for _ in range(n_bursts):
        k = kinds[int(rng.integers(0, len(kinds)))]
        if k == "silvercrest":
            seg = silvercrest_burst(None if decodable else rng)
        elif k == "nice":
            seg = nice_flor_s_burst(None if decodable else rng)
        elif k == "nexus":
            seg = nexus_burst(rng)
        else:
            seg = manchester_burst(rng)
        masks.append(_render_ook(seg, rate))
    min_gap = int(0.008 * rate)
    lead_in = int(0.008 * rate)
    total = sum(len(m) for m in masks) + min_gap * (len(masks) + 1) + lead_in
    slack = n_samples - total
    if slack < 0:
        raise ValueError("stream too short for the requested bursts")
    cuts = np.sort(rng.integers(0, slack + 1, len(masks)))
    pos = lead_in
    prev = 0
    for m, c in zip(masks, cuts):
        pos += min_gap + int(c - prev)
        prev = c
        amp = np.float32(rng.uniform(40, 110))
        f = rng.uniform(-60e3, 60e3)
        ph = rng.uniform(0, 2 * np.pi) + 2 * np.pi * f / rate * np.arange(len(m), dtype=np.float64)
        on = m.astype(np.float32) * amp
        x[2 * pos:2 * (pos + len(m)):2] += on * np.cos(ph).astype(np.float32)
        x[2 * pos + 1:2 * (pos + len(m)) + 1:2] += on * np.sin(ph).astype(np.float32)
        pos += len(m)
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def silvercrest_file(rate=OOK_RATE, noise_sigma=0.0, seed=1):
    """BASELINE config 1: four Silvercrest rows, 8 ms lead-in, 50 kHz tone, amplitude 100.
    noise_sigma == 0 reproduces the constant 128/128 silence of tests/rtl_tcp_serve.py."""
    seg = [(8000, 0)] + silvercrest_burst() + [(20000, 0)]
    m = _render_ook(seg, rate)
    n = len(m)
    ph = 2 * np.pi * 50000.0 / rate * np.cumsum(m, dtype=np.float64)
    if noise_sigma > 0:
        rng = np.random.default_rng(seed)
        base = rng.standard_normal(2 * n) * noise_sigma + 127.5
    else:
        base = np.full(2 * n, 128.0)
    base[0::2] += m * 100.0 * np.cos(ph)
    base[1::2] += m * 100.0 * np.sin(ph)
    if noise_sigma > 0:
        return np.clip(np.rint(base), 0, 255).astype(np.uint8)
    return np.clip(np.trunc(base), 0, 255).astype(np.uint8)


def nice_flor_s_file(rate=OOK_RATE):
    """The reference's own vector: tests/rtl_tcp_serve.py synth_cu8 with the arguments of
    tests/http-rtltcp-test.sh (bits of e7a760b94372e, 500/1000/1500 us, gap 500, lead-in 8 ms,
    reset 20 ms), restated: int() truncation towards zero, silence = 128/128."""
    bits = "".join(f"{int(ch, 16):04b}" for ch in "e7a760b94372e")
    out = []
    phase = 0.0
    spu = rate / 1e6

    def emit(us, on):
        nonlocal phase
        for _ in range(int(round(us * spu))):
            if on:
                i = 128 + int(100 * np.cos(phase))
                q = 128 + int(100 * np.sin(phase))
                phase += 2 * np.pi * 50000.0 / rate
            else:
                i = q = 128
            out.append(max(0, min(255, i)))
            out.append(max(0, min(255, q)))

    emit(8000, False)
    for b in bits:
        emit(500 if b == "1" else 1000, True)
        emit(500, False)
    emit(1500, True)
    emit(20000, False)
    return np.array(out, np.uint8)


def fsk_stream(seed, n_samples=1 << 20, rate=FSK_RATE, n_bursts=4, sigma=40.0, bit_us=100.0, dev_hz=40e3):
    """cs16 IQ (int16[2*n_samples]): 2-FSK NRZ PCM bursts (0xAAAAAAAA preamble + 96 random bits)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(2 * n_samples, dtype=np.float32) * np.float32(sigma)
    spb = bit_us * rate / 1e6
    nbits = 32 + 96
    blen = int(np.ceil(nbits * spb))
    min_gap = int(0.02 * rate)
    lead_in = int(0.008 * rate)
    total = n_bursts * (blen + min_gap) + lead_in + min_gap
    slack = n_samples - total
    if slack < 0:
        raise ValueError("stream too short for the requested bursts")
    cuts = np.sort(rng.integers(0, slack + 1, n_bursts))
    pos = lead_in
    prev = 0
    for c in cuts:
        pos += min_gap + int(c - prev)
        prev = c
        bits = np.concatenate([np.tile([1, 0], 16), rng.integers(0, 2, 96)])
        idx = np.minimum((np.arange(blen) / spb).astype(np.int64), nbits - 1)
        freq = np.where(bits[idx] > 0, dev_hz, -dev_hz) + rng.uniform(-10e3, 10e3)
        ph = rng.uniform(0, 2 * np.pi) + 2 * np.pi * np.cumsum(freq) / rate
        amp = np.float32(rng.uniform(4000, 14000))
        x[2 * pos:2 * (pos + blen):2] += amp * np.cos(ph).astype(np.float32)
        x[2 * pos + 1:2 * (pos + blen) + 1:2] += amp * np.sin(ph).astype(np.float32)
        pos += blen
    return np.clip(np.rint(x), -32767, 32767).astype(np.int16)


def ook_train_stream(seed, n_pulses, on_us=200.0, off_us=200.0, n_samples=1 << 19, rate=OOK_RATE, sigma=2.0,
                     jitter_us=0.0, lead_us=8000.0):
    """cu8 IQ with ONE long on/off keyed train of `n_pulses` pulses (more than PD_MAX_PULSES = 1200 makes the
    detector return a full package mid-train, src/pulse_detect.c:429-441) over a noise floor."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(2 * n_samples, dtype=np.float32) * np.float32(sigma) + np.float32(127.5)
    seg = []
    for _ in range(n_pulses):
        seg.append((on_us + (rng.uniform(-jitter_us, jitter_us) if jitter_us else 0.0), 1))
        seg.append((off_us + (rng.uniform(-jitter_us, jitter_us) if jitter_us else 0.0), 0))
    m = _render_ook(seg, rate)
    pos = int(lead_us * rate / 1e6)
    if pos + len(m) + int(0.02 * rate) > n_samples:
        raise ValueError("stream too short for the requested train")
    amp = np.float32(rng.uniform(60, 110))
    f = rng.uniform(-40e3, 40e3)
    ph = rng.uniform(0, 2 * np.pi) + 2 * np.pi * f / rate * np.arange(len(m), dtype=np.float64)
    on = m.astype(np.float32) * amp
    x[2 * pos:2 * (pos + len(m)):2] += on * np.cos(ph).astype(np.float32)
    x[2 * pos + 1:2 * (pos + len(m)) + 1:2] += on * np.sin(ph).astype(np.float32)
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def cu8_to_cs16(x, gain=96):
    """The same capture as a cs16 file would hold it: (v - 128) * gain (|gain| <= 256 keeps it inside int16)."""
    return ((x.astype(np.int16) - 128) * np.int16(gain)).astype(np.int16)


def fsk_burst_stream(seed, n_bits, bit_us=100.0, n_samples=1 << 19, rate=FSK_RATE, sigma=40.0, dev_hz=40e3, cu8=False,
                     alternate=True, lead_us=8000.0):
    """ONE 2-FSK burst of `n_bits` bits (alternating by default: one frequency transition per bit, so more than
    2 x 1200 bits overflow the FSK pulse train and exercise pulse_data_shift, src/pulse_detect_fsk.c:114,205).
    cs16 by default; cu8=True renders the same burst as unsigned 8-bit IQ (amplitude 100 LSB, sigma 2)."""
    rng = np.random.default_rng(seed)
    spb = bit_us * rate / 1e6
    blen = int(np.ceil(n_bits * spb))
    pos = int(lead_us * rate / 1e6)
    if pos + blen + int(0.02 * rate) > n_samples:
        raise ValueError("stream too short for the requested burst")
    bits = np.tile([1, 0], (n_bits + 1) // 2)[:n_bits] if alternate else rng.integers(0, 2, n_bits)
    idx = np.minimum((np.arange(blen) / spb).astype(np.int64), n_bits - 1)
    freq = np.where(bits[idx] > 0, dev_hz, -dev_hz) + rng.uniform(-10e3, 10e3)
    ph = rng.uniform(0, 2 * np.pi) + 2 * np.pi * np.cumsum(freq) / rate
    if cu8:
        x = rng.standard_normal(2 * n_samples, dtype=np.float32) * np.float32(2.0) + np.float32(127.5)
        amp = np.float32(100.0)
    else:
        x = rng.standard_normal(2 * n_samples, dtype=np.float32) * np.float32(sigma)
        amp = np.float32(rng.uniform(4000, 14000))
    x[2 * pos:2 * (pos + blen):2] += amp * np.cos(ph).astype(np.float32)
    x[2 * pos + 1:2 * (pos + blen) + 1:2] += amp * np.sin(ph).astype(np.float32)
    if cu8:
        return np.clip(np.rint(x), 0, 255).astype(np.uint8)
    return np.clip(np.rint(x), -32767, 32767).astype(np.int16)
