// r433b_slice.cuh -- the pulse slicers (pulse train -> bitbuffer rows) and the compact event
// encoder they write through.  __host__ __device__ inlines: device code in r433b_kernels.cu,
// CPU-side unit tests through tests/host_core.cpp.
//
// One slicer run = one (package, device) pair, executed by one thread.  Instead of building a
// 6604-byte bitbuffer_t (include/bitbuffer.h:34-40) per event, rows are streamed straight into
// the pair's private output region in the wire format below; the host re-inflates a real
// bitbuffer_t only when it calls a decoder (r433b_host.cpp).  Every run is executed twice by
// the kernel: once counting (size), once storing.
//
//   event   := u32 { num_rows:7, dirty:1, free_row:8, words:16 }   words = whole event incl. this
//              row*  [u32 last_row_words  -- only if dirty]
//   row     := u32 { bits:16, syncs:16 }  then ceil(bits/32) data words (physical row order;
//              bit i of the row is bit (7 - i%8) of byte i/8, exactly bitbuffer_t's bb layout)
//
// `dirty` marks the reference's 50-row overflow path, which zeroes the last row's length but
// keeps its bytes (src/bitbuffer.c:118-121): the number of data words then comes from the trailer.
#pragma once
#include <stdint.h>
#include "r433b_core.cuh"

namespace r433b {

constexpr int kBbRows = 50;   // include/bitbuffer.h:28
constexpr int kBbCols = 128;  // include/bitbuffer.h:27

// float helpers: no fused multiply-add anywhere (the reference is built without contraction)
#ifdef __CUDA_ARCH__
R4_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
R4_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
R4_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
R4_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
R4_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
#else
R4_HD float fmul(float a, float b) { volatile float r = a * b; return r; }
R4_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
R4_HD float fdiv(float a, float b) { volatile float r = a / b; return r; }
R4_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
R4_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
#endif

// Integer timing of one device at one sample rate.  Derived ON THE HOST (r433b_host.cpp) with
// the reference's float expressions (src/pulse_slicer.c:70-91) and shipped as integers.
struct SlicerParams {
    int modulation;
    int ok; // 0: "sample rate too low" -> the slicer returns without events
    int s_short, s_long, s_reset, s_gap, s_sync, s_tol;
    float f_short, f_long; // 1/(width*samples_per_us) or 0
    unsigned priority;
    int pad;
};

// ----------------------------------------------------------------------------- writer ----

// One writer type serves both passes: out == nullptr only counts.  All offsets are in 32-bit
// words; every store is one aligned word.
struct EventWriter {
    uint32_t *out;      // pair region, or nullptr while counting
    unsigned limit;     // words the counting pass committed: rows of a trailing, never-emitted
                        // event lie beyond it and must not be written
    unsigned pos;       // words used by finished events + the current event so far
    unsigned committed; // words up to the end of the last emitted event
    unsigned events;
    // current event
    unsigned ev_start;
    unsigned num_rows, free_row;
    unsigned row0_bits; // bits_per_row[0] once row 0 is closed
    // current (last) row
    unsigned row_hdr;   // word offset of its header
    unsigned bits, syncs;
    unsigned row_hw;    // data words that hold something
    unsigned acc;       // word being assembled
    bool dirty;         // row length was reset while its bytes stayed (50-row overflow path)

    R4_HD void init(uint32_t *o, unsigned region_words = 0)
    {
        out = o;
        limit = region_words;
        pos = committed = 0;
        events = 0;
        reset_event();
    }

    R4_HD void reset_event() // bitbuffer_clear(), src/bitbuffer.c:17
    {
        pos = committed;
        ev_start = committed;
        num_rows = free_row = 0;
        row0_bits = 0;
        row_hdr = 0;
        bits = syncs = row_hw = 0;
        acc = 0;
        dirty = false;
    }

    R4_HD void put(unsigned at, uint32_t v)
    {
        if (out && at < limit) out[at] = v;
    }

    R4_HD unsigned first_row_bits() const { return num_rows <= 1 ? bits : row0_bits; }
    R4_HD unsigned last_row_bits() const { return bits; }

    R4_HD void open_row()
    {
        row_hdr = pos;
        pos += 1;
        bits = syncs = row_hw = 0;
        acc = 0;
    }

    R4_HD void flush_word()
    {
        unsigned w = (bits - 1) >> 5; // the word holding the newest bit
        unsigned at = row_hdr + 1 + w;
        if (out && at < limit) {
            if (w < row_hw)
                out[at] |= acc;
            else
                out[at] = acc;
        }
        if (w + 1 > row_hw) row_hw = w + 1;
        acc = 0;
    }

    R4_HD void close_row()
    {
        if (bits & 31) flush_word();
        put(row_hdr, bits | (syncs << 16));
        pos = row_hdr + 1 + row_hw;
    }

    R4_HD void first_row() // "Add first row automatically", src/bitbuffer.c:24-26
    {
        if (num_rows == 0) {
            ev_start = pos;
            pos += 1;
            num_rows = free_row = 1;
            open_row();
        }
    }

    // src/bitbuffer.c:22-56; bit i of a row lives in byte i/8 at (7 - i%8): within a
    // little-endian word that is shift (i & 31) ^ 7
    R4_HD void add_bit(int bit)
    {
        first_row();
        if (bits == 65535u) return;
        if (bits > 0 && (bits & (kBbCols * 8 - 1)) == 0) { // spill into the next physical row
            if (free_row < (unsigned)kBbRows)
                free_row++;
            else
                return;
        }
        acc |= (uint32_t)bit << ((bits & 31) ^ 7);
        bits++;
        if ((bits & 31) == 0) flush_word();
    }

    // src/bitbuffer.c:106-122
    R4_HD void add_row()
    {
        first_row();
        if (free_row < (unsigned)kBbRows) {
            if (num_rows == 1) row0_bits = bits;
            close_row();
            free_row++;
            // physical rows taken by spill-over sit between the old and the new last row
            for (unsigned r = num_rows; r + 1 < free_row; ++r) {
                put(pos, 0);
                pos += 1;
            }
            num_rows = free_row;
            open_row();
        } else {
            // row count exhausted: length forgotten, bytes (and syncs) stay
            if (bits & 31) flush_word();
            bits = 0;
            dirty = true;
            acc = 0;
        }
    }

    // src/bitbuffer.c:124-133
    R4_HD void add_sync()
    {
        first_row();
        if (bits) add_row();
        syncs++;
    }

    // account_event(): hand the buffer to the decoder, then clear it (src/pulse_slicer.c:26-66)
    R4_HD void emit()
    {
        if (num_rows == 0) { // an empty buffer is still an event (e.g. nrzs): header only
            ev_start = pos;
            pos += 1;
        } else {
            close_row();
            if (dirty) { // only the last row can be: its data length travels in a trailer word
                put(pos, row_hw);
                pos += 1;
            }
        }
        put(ev_start, num_rows | (dirty ? 0x80u : 0u) | (free_row << 8) | ((pos - ev_start) << 16));
        committed = pos;
        events++;
        reset_event();
    }
};

// ---------------------------------------------------------------------------- slicers ----

struct PulseView {
    int const *pulse;
    int const *gap;
    unsigned n;
};

R4_HD int iabs(int v) { return v < 0 ? -v : v; }
R4_HD bool within(int v, int nominal, int tol) { return v >= nominal - tol && v <= nominal + tol; }

// src/pulse_slicer.c:68-259
template <class W>
R4_HD void slice_pcm(PulseView const &p, SlicerParams const &t, W &w)
{
    if (!t.ok || t.s_long == 0) return;
    float f_short = t.f_short, f_long = t.f_long;
    int const gap_limit = t.s_gap ? t.s_gap : t.s_reset;
    int const max_zeros = gap_limit / t.s_long;
    int tol = t.s_tol;
    if (tol <= 0) tol = t.s_long / 4;
    bool const rz = t.s_short != t.s_long;
    int need = rz ? 4 : 12;
    int preamble = 0;
    unsigned const N = p.n;
    if (rz) {
        for (unsigned n = 0; n < N; ++n) { // :105-132
            int sw = 0, lw = 0, cnt = 0;
            while (n < N && within(p.pulse[n], t.s_short, tol) && within(p.pulse[n] + p.gap[n], t.s_long, tol)) {
                sw += p.pulse[n];
                lw += p.pulse[n] + p.gap[n];
                cnt++;
                n++;
            }
            if (cnt >= need) {
                f_long = fdiv((float)cnt, (float)lw);
                f_short = fdiv((float)cnt, (float)sw);
                need = cnt;
                preamble = cnt;
            }
        }
        if (preamble == 0) { // :134-157
            int sw = 0, lw = 0, cnt = 0;
            for (unsigned n = 0; n < N; ++n) {
                if (within(p.pulse[n], t.s_short, tol) && within(p.pulse[n] + p.gap[n], t.s_long, tol)) {
                    sw += p.pulse[n];
                    lw += p.pulse[n] + p.gap[n];
                    cnt++;
                }
            }
            if (cnt > 8) {
                f_long = fdiv((float)cnt, (float)lw);
                f_short = fdiv((float)cnt, (float)sw);
            }
        }
    } else {
        for (unsigned n = 0; n < N; ++n) { // :159-180, float product then DOUBLE +0.5
            int wsum = 0, cnt = 0;
            while (n < N && (int)dadd((double)fmul((float)p.pulse[n], f_short), 0.5) == 1
                    && (int)dadd((double)fmul((float)p.gap[n], f_long), 0.5) == 1) {
                wsum += p.pulse[n] + p.gap[n];
                cnt += 2;
                n++;
            }
            if (cnt >= need) {
                f_short = f_long = fdiv((float)cnt, (float)wsum);
                need = cnt;
                preamble = cnt;
            }
        }
        if (preamble == 0) { // :182-214
            int wsum = 0, cnt = 0;
            for (unsigned n = 0; n < N; ++n) {
                if (within(p.pulse[n], t.s_short, tol)) { wsum += p.pulse[n]; cnt += 1; }
                if (within(p.pulse[n], 2 * t.s_short, tol)) { wsum += p.pulse[n]; cnt += 2; }
                if (within(p.gap[n], t.s_long, tol)) { wsum += p.gap[n]; cnt += 1; }
                if (within(p.gap[n], 2 * t.s_long, tol)) { wsum += p.gap[n]; cnt += 2; }
            }
            if (cnt > 20) f_short = f_long = fdiv((float)cnt, (float)wsum);
        }
    }
    for (unsigned n = 0; n < N; ++n) { // :216-257
        int highs = (int)fadd(fmul((float)p.pulse[n], f_short), 0.5f);
        int lows = (int)fadd(fmul((float)(p.gap[n] + t.s_short - t.s_long), f_long), 0.5f);
        for (int i = 0; i < highs; ++i) w.add_bit(1);
        if (lows > max_zeros) lows = max_zeros;
        for (int i = 0; i < lows; ++i) w.add_bit(0);
        if (rz && iabs(p.pulse[n] - t.s_short) > tol)
            w.reset_event();
        else if (p.gap[n] > gap_limit && p.gap[n] <= t.s_reset)
            w.add_row();
        if ((n == N - 1 || p.gap[n] > t.s_reset) && (w.first_row_bits() > 0 || w.num_rows > 1)) w.emit();
    }
}

// src/pulse_slicer.c:261-337
template <class W>
R4_HD void slice_ppm(PulseView const &p, SlicerParams const &t, W &w)
{
    if (!t.ok) return;
    int z_lo, z_hi, o_lo, o_hi, s_lo = 0, s_hi = 0;
    if (t.s_tol > 0) {
        z_lo = t.s_short - t.s_tol; z_hi = t.s_short + t.s_tol;
        o_lo = t.s_long - t.s_tol;  o_hi = t.s_long + t.s_tol;
        if (t.s_sync > 0) { s_lo = t.s_sync - t.s_tol; s_hi = t.s_sync + t.s_tol; }
    } else {
        z_lo = 0;
        z_hi = (t.s_short + t.s_long) / 2 + 1;
        o_lo = z_hi - 1;
        o_hi = t.s_gap ? t.s_gap : t.s_reset;
    }
    for (unsigned n = 0; n < p.n; ++n) {
        int g = p.gap[n];
        if (g > z_lo && g < z_hi) w.add_bit(0);
        else if (g > o_lo && g < o_hi) w.add_bit(1);
        else if (g > s_lo && g < s_hi) w.add_sync();
        else if (g < t.s_reset) w.add_row();
        if ((n == p.n - 1 || g >= t.s_reset) && (w.first_row_bits() > 0 || w.num_rows > 1)) w.emit();
    }
}

// src/pulse_slicer.c:339-449
template <class W>
R4_HD void slice_pwm(PulseView const &p, SlicerParams const &t, W &w)
{
    if (!t.ok) return;
    int const big = 0x7fffffff;
    int o_lo, o_hi, z_lo, z_hi, s_lo = 0, s_hi = 0;
    if (t.s_tol > 0) {
        o_lo = t.s_short - t.s_tol; o_hi = t.s_short + t.s_tol;
        z_lo = t.s_long - t.s_tol;  z_hi = t.s_long + t.s_tol;
        if (t.s_sync > 0) { s_lo = t.s_sync - t.s_tol; s_hi = t.s_sync + t.s_tol; }
    } else if (t.s_sync <= 0) {
        o_lo = 0; o_hi = (t.s_short + t.s_long) / 2 + 1;
        z_lo = o_hi - 1; z_hi = big;
    } else if (t.s_sync < t.s_short) {
        s_lo = 0; s_hi = (t.s_sync + t.s_short) / 2 + 1;
        o_lo = s_hi - 1; o_hi = (t.s_short + t.s_long) / 2 + 1;
        z_lo = o_hi - 1; z_hi = big;
    } else if (t.s_sync < t.s_long) {
        o_lo = 0; o_hi = (t.s_short + t.s_sync) / 2 + 1;
        s_lo = o_hi - 1; s_hi = (t.s_sync + t.s_long) / 2 + 1;
        z_lo = s_hi - 1; z_hi = big;
    } else {
        o_lo = 0; o_hi = (t.s_short + t.s_long) / 2 + 1;
        z_lo = o_hi - 1; z_hi = (t.s_long + t.s_sync) / 2 + 1;
        s_lo = z_hi - 1; s_hi = big;
    }
    for (unsigned n = 0; n < p.n; ++n) {
        int v = p.pulse[n];
        if (v > o_lo && v < o_hi) w.add_bit(1);
        else if (v > z_lo && v < z_hi) w.add_bit(0);
        else if (v > s_lo && v < s_hi) w.add_sync();
        else if (v <= o_lo) { }
        else w.add_row();
        if ((n == p.n - 1 || p.gap[n] > t.s_reset) && w.num_rows > 0)
            w.emit();
        else if (t.s_gap > 0 && p.gap[n] > t.s_gap && w.num_rows > 0 && w.last_row_bits() > 0)
            w.add_row();
    }
}

// src/pulse_slicer.c:451-527; the 1.5 x short comparisons are in double
template <class W>
R4_HD void slice_manchester(PulseView const &p, SlicerParams const &t, W &w)
{
    if (!t.ok) return;
    int since = 0;
    w.add_bit(0);
    double const edge = dmul((double)t.s_short, 1.5);
    int const lo = t.s_short - t.s_tol, hi = t.s_short * 2 + t.s_tol;
    for (unsigned n = 0; n < p.n; ++n) {
        int v = p.pulse[n], g = p.gap[n];
        if (t.s_tol > 0 && (v < lo || v > hi || g < lo || g > hi)) {
            if ((double)v > edge && v <= hi) w.add_bit(1);
            w.add_row();
            w.add_bit(0);
            since = 0;
        } else if ((double)(v + since) > edge) {
            w.add_bit(1);
            since = 0;
        } else {
            since += v;
        }
        if ((n == p.n - 1 || g > t.s_reset) && w.num_rows > 0) {
            w.emit();
            w.add_bit(0);
            since = 0;
        } else if ((double)(g + since) > edge) {
            w.add_bit(0);
            since = 0;
        } else {
            since += g;
        }
    }
}

R4_HD int symbol_at(PulseView const &p, unsigned k) { return (k & 1) ? p.gap[k >> 1] : p.pulse[k >> 1]; } // :529-535

// src/pulse_slicer.c:537-595
template <class W>
R4_HD void slice_dmc(PulseView const &p, SlicerParams const &t, W &w)
{
    if (!t.ok) return;
    unsigned const total = p.n * 2;
    for (unsigned k = 0; k < total; ++k) {
        int s = symbol_at(p, k);
        if (iabs(s - t.s_short) < t.s_tol) {
            w.add_bit(1);
            s = k + 1 < total ? symbol_at(p, ++k) : 0;
            if (iabs(s - t.s_short) > t.s_tol) {
                if (s >= t.s_reset - t.s_tol)
                    k--;
                else if (w.num_rows > 0 && w.last_row_bits() > 0)
                    w.add_row();
            }
        } else if (iabs(s - t.s_long) < t.s_tol) {
            w.add_bit(0);
        } else if (s >= t.s_reset - t.s_tol && w.num_rows > 0) {
            w.emit();
        }
    }
}

// src/pulse_slicer.c:597-657
template <class W>
R4_HD void slice_piwm_raw(PulseView const &p, SlicerParams const &t, W &w)
{
    if (!t.ok) return;
    unsigned const total = p.n * 2;
    for (unsigned k = 0; k < total; ++k) {
        int s = symbol_at(p, k);
        int cnt = (int)dadd((double)fmul((float)s, t.f_short), 0.5);
        if (s > t.s_long) {
            w.add_row();
        } else if (iabs(s - cnt * t.s_short) < t.s_tol) {
            for (; cnt > 0; --cnt) w.add_bit(1 - (int)(k & 1));
        } else if (s < t.s_reset && w.num_rows > 0 && w.last_row_bits() > 0) {
            w.add_row();
        }
        if ((k == total - 1 || s > t.s_reset) && w.num_rows > 0) w.emit();
    }
}

// src/pulse_slicer.c:659-713
template <class W>
R4_HD void slice_piwm_dc(PulseView const &p, SlicerParams const &t, W &w)
{
    if (!t.ok) return;
    unsigned const total = p.n * 2;
    for (unsigned k = 0; k < total; ++k) {
        int s = symbol_at(p, k);
        if (iabs(s - t.s_short) < t.s_tol) w.add_bit(1);
        else if (iabs(s - t.s_long) < t.s_tol) w.add_bit(0);
        else if (s < t.s_reset && w.num_rows > 0 && w.last_row_bits() > 0) w.add_row();
        if ((k == total - 1 || s > t.s_reset) && w.num_rows > 0) w.emit();
    }
}

// src/pulse_slicer.c:715-759
template <class W>
R4_HD void slice_nrzs(PulseView const &p, SlicerParams const &t, W &w)
{
    if (!t.ok || t.s_short == 0) return;
    int const limit = t.s_short;
    for (unsigned n = 0; n < p.n; ++n) {
        if (p.pulse[n] > limit) {
            int reps = p.pulse[n] / limit;
            for (int i = 0; i < reps; ++i) w.add_bit(1);
            w.add_bit(0);
        } else if (p.pulse[n] < limit) {
            w.add_bit(0);
        }
        if (n == p.n - 1 || p.gap[n] >= t.s_reset) w.emit();
    }
}

// src/pulse_slicer.c:775-864
template <class W>
R4_HD void slice_osv1(PulseView const &p, SlicerParams const &t, W &w)
{
    if (!t.ok) return;
    int pre = 0, man = 0;
    int const half_lo = t.s_short / 2, half_hi = t.s_short * 3 / 2, sync_lo = 2 * half_hi;
    unsigned n;
    for (n = 0; n < p.n; ++n) {
        if (p.pulse[n] > half_lo && p.gap[n] > half_lo) {
            pre++;
            if (p.gap[n] > half_hi) break;
        } else {
            return;
        }
    }
    if (pre != 12) return;
    ++n;
    if (n >= (unsigned)kMaxPulses) return; // the reference reads past the array here
    // n may equal num_pulses: the entry after the last pulse is part of the package record
    if (p.pulse[n] < sync_lo || p.gap[n] < sync_lo) return;
    if (p.gap[n] > p.pulse[n]) {
        man ^= 1;
        if (man) w.add_bit(0);
    }
    for (n++; n < p.n; ++n) {
        man ^= 1;
        if (man) w.add_bit(1);
        if (p.pulse[n] > half_hi) {
            man ^= 1;
            if (man) w.add_bit(1);
        }
        if ((n == p.n - 1 || p.gap[n] > t.s_reset) && w.num_rows > 0) {
            w.emit();
            return;
        }
        man ^= 1;
        if (man) w.add_bit(0);
        if (p.gap[n] > half_hi) {
            man ^= 1;
            if (man) w.add_bit(0);
        }
    }
}

// src/pulse_slicer.c:866-918 (its rate check only looks at short/long/reset: `ok` bit 1)
template <class W>
R4_HD void slice_rzi(PulseView const &p, SlicerParams const &t, W &w)
{
    if (!(t.ok & 2) || t.s_long == 0) return;
    int const s_base = t.s_long - t.s_short;
    bool fresh = true;
    for (unsigned n = 0; n < p.n; ++n) {
        int high = p.pulse[n];
        int ones = fresh ? (high + t.s_long / 2) / t.s_long : (high - s_base + t.s_long / 2) / t.s_long;
        fresh = false;
        for (int k = 0; k < ones; ++k) w.add_bit(1);
        if (p.gap[n] > t.s_reset || n == p.n - 1) {
            if (w.first_row_bits() > 0) w.emit();
            w.reset_event();
            fresh = true;
            continue;
        }
        w.add_bit(0);
    }
}

enum { // include/r_device.h:24-40
    kModOokMc = 3, kModOokPcm = 4, kModOokPpm = 5, kModOokPwm = 6, kModOokPiwmRaw = 8, kModOokDmc = 9,
    kModOokOsv1 = 10, kModOokPiwmDc = 11, kModOokNrzs = 12, kModOokRzi = 13,
    kModFskPcm = 16, kModFskPwm = 17, kModFskMc = 18
};

// does run_ook_demods / run_fsk_demods (src/r_api.c:438-550) hand this package type to this device?
R4_HD bool device_takes(int modulation, int package_type)
{
    if (package_type == 1) return modulation >= 3 && modulation <= 13 && modulation != 7;
    return modulation >= 16 && modulation <= 18;
}

template <class W>
R4_HD void slice_dispatch(PulseView const &p, SlicerParams const &t, W &w)
{
    switch (t.modulation) {
    case kModOokPcm: case kModFskPcm: slice_pcm(p, t, w); break;
    case kModOokPpm: slice_ppm(p, t, w); break;
    case kModOokPwm: case kModFskPwm: slice_pwm(p, t, w); break;
    case kModOokMc: case kModFskMc: slice_manchester(p, t, w); break;
    case kModOokPiwmRaw: slice_piwm_raw(p, t, w); break;
    case kModOokPiwmDc: slice_piwm_dc(p, t, w); break;
    case kModOokDmc: slice_dmc(p, t, w); break;
    case kModOokOsv1: slice_osv1(p, t, w); break;
    case kModOokNrzs: slice_nrzs(p, t, w); break;
    case kModOokRzi: slice_rzi(p, t, w); break;
    default: break;
    }
}

} // namespace r433b
