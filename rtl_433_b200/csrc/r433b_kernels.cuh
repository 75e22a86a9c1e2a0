// r433b_kernels.cuh -- sm_100a kernels of the IQ -> package -> event path.
//
//   k_detect<SS> : one WARP per capture stream, walking it tile by tile with all carried state
//                  in registers.  Per tile: 128-bit loads of the lane's contiguous IQ chunk,
//                  envelope / phase-discriminator maps in parallel on all lanes, the two
//                  integer IIR low-passes evaluated EXACTLY in parallel (bracket rounds, below),
//                  AM/FM tile staged in shared memory, then the package detector state machine
//                  over the tile (warp-uniform, with warp-ballot scans for the threshold-constant
//                  states).  IQ is read once; intermediates never touch HBM.
//   k_slice      : one thread per (package, device): slicer run twice (count, then store) with a
//                  warp-aggregated arena allocation in between.
//
// Exact parallel IIR.  y' = (a*y + c[n]) >> 14 is monotone in y and contracts by a/2^14 per
// sample, but the floor makes it non-associative.  Each lane owns C consecutive samples and
// keeps a bracket [lo, hi] for the filter state at the start of its chunk (lane 0: the exact
// carried state; others: the full int16 range).  One "round" pushes both ends through the
// PREVIOUS lane's chunk; monotonicity keeps the true state inside, contraction shrinks the
// bracket by a^C per round, and after every round at least one more lane is exact, so the loop
// terminates in <= 31 rounds (constant input, where lo/hi sit on different fixed points of the
// floor map) and typically in 3.  Lanes then run their chunk once from the exact state.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/r433b.h"
#include "r433b_core.cuh"
#include "r433b_slice.cuh"

#ifndef R4_UNROLL_IIR
#define R4_UNROLL_IIR 2
#endif
#ifndef R4_UNROLL_TILE
#define R4_UNROLL_TILE 2
#endif
#ifndef R4_UNROLL_IDLE
#define R4_UNROLL_IDLE 4
#endif
#ifndef R4_UNROLL_PULSE
#define R4_UNROLL_PULSE 2
#endif
#define R4_PRAGMA(x) _Pragma(#x)
#define R4_UNROLL(n) R4_PRAGMA(unroll n)

namespace r433b {

constexpr int kTrainInts = 4 * kMaxPulses; // per-stream scratch: ook pulse/gap, fsk pulse/gap
constexpr int kDetectWarps = 4;            // warps (streams) per CTA
constexpr int kDetectCtasPerSm = 7;        // 28 warps per SM: 4096 streams are co-resident on 148 SMs

// Everything one stream carries from one launch to the next when a batch is processed in
// several time slices (so that copy-in of slice k+1 overlaps the kernels of slice k).
struct StreamState {
    DetState d;
    int y_am, y_fm, x_prev, xf_prev;
    unsigned fm_tile; // index of the tile for which (y_fm, xf_prev) are the FM low-pass carry-in
    unsigned seq;
    int flushed;
};

struct DetectParams {
    uint8_t const *data;
    unsigned long long const *offsets; // bytes, n_streams + 1
    unsigned long long const *lengths; // optional: bytes of stream i actually used
    unsigned n_streams;
    unsigned stream0, stream_end; // the streams this launch covers
    unsigned long long sample_begin, sample_end; // the slice of every stream this launch covers (multiples of the tile)
    int first_chunk;              // start from reset_sdr_flow() state instead of the saved one
    struct StreamState *state;    // per-stream carried state between launches of one batch
    int use_mag, enable_fm, fpdm;
    int lazy_fm; // compute discriminator + FM low-pass only for tiles the detector reads them in
    unsigned flip; // XOR mask applied to every loaded word: 0x80808080 turns cs8 into cu8
    unsigned rate, block_samples;
    Levels lv;
    int lpf_a1, lpf_b0, fm_a1, fm_b0;
    int wrap_free;
    int *train_scratch;
    r433b_package *pkgs;
    unsigned pkg_cap;
    int *pulse_pool, *gap_pool;
    unsigned pool_cap;
    unsigned *counters; // [0] packages, [1] pool entries, [2] overflow flag
    int16_t *am_out, *fm_out; // optional stage dump, indexed by offsets[s]/SS + n
};

struct WarpCtx {
    int lane;
    int nlanes;
    __device__ __forceinline__ void sync() { __syncwarp(); }
};

// Shared-memory tile of one warp: sample n of the tile lives at word (n / C) * (W*C + 1) + (n % C) * W.
// The odd chunk stride makes both access patterns conflict-free: every lane walking its own
// chunk (IIR passes) and 32 lanes reading 32 consecutive samples (detector scans).
// Word 0 of a sample holds envelope | discriminator << 16 (cu8) and later AM | FM << 16;
// cs16 keeps its 32-bit discriminator output in word 1.
template <int C, int W>
__device__ __forceinline__ int word_index(int n)
{
    return (n / C) * (W * C + 1) + (n % C) * W;
}

// ------------------------------------------------------------------------ k_detect ------

template <int SS>
struct TileCfg {
    static constexpr int C = SS == 2 ? 32 : 16; // samples per lane per tile (64 bytes of IQ)
    static constexpr int T = 32 * C;
    static constexpr int W = SS == 2 ? 1 : 2;   // shared-memory words per sample
    static constexpr int kTileWords = 32 * (W * C + 1);
};

// 16 contiguous bytes (8 cu8 / 4 cs16 samples) of tile tt for this lane: one fully coalesced
// 128-bit load per lane; zero-filled past the nvt valid samples of the tile.
template <int SS>
__device__ __forceinline__ void load_group(uint8_t const *src, unsigned long long tt, int n0, int nvt, unsigned flip,
        uint32_t (&rw)[4])
{
    constexpr int SPL = 16 / SS;
    rw[0] = rw[1] = rw[2] = rw[3] = 0u;
    uint8_t const *g = src + (tt + (unsigned long long)n0) * SS;
    if (n0 + SPL <= nvt) {
        uint4 v = __ldg(reinterpret_cast<uint4 const *>(g));
        rw[0] = v.x ^ flip;
        rw[1] = v.y ^ flip;
        rw[2] = v.z ^ flip;
        rw[3] = v.w ^ flip;
    } else if (n0 < nvt) { // ragged end of the stream
        int nb = (nvt - n0) * SS;
        for (int bidx = 0; bidx < nb; ++bidx) rw[bidx >> 2] |= (uint32_t)(g[bidx] ^ (flip & 0xff)) << (8 * (bidx & 3));
    }
}

// ------------------------------------------------------------------- FM on demand ------
//
// The discriminator and its low-pass are only looked at inside packages, so k_detect computes
// them per tile when the detector first asks.  fm_make() is deliberately NOT inlined: it is the
// cold side of the tile loop and keeps its own register allocation.
//
// fm_tile(tt): discriminator of tile tt (IQ re-read from global memory: L1/L2 hits for the
// current tile) into the high halves / second words of the shared-memory tile, then the FM
// low-pass by bracket rounds.  For the CURRENT tile every lane needs its exact start state and a
// final pass writes FM next to AM.  For an EARLIER tile only the state at its end matters: rounds
// stop as soon as lane 31's end bracket has collapsed.  `unknown` starts lane 0 from the full
// range too (its first sample only provides x[n-1]).
//
// If the filter state is stale (tiles were skipped) it is rebuilt from the previous tile alone:
// from ANY start state the brackets collapse within a few dozen samples of a live signal.  If they
// do not (a constant discriminator output parks the two ends on different fixed points of the
// floor map), walk forward from the last exact state instead.
struct FmCarry {
    int y, xf; // filter state and discriminator output after the last sample of the tile
};

template <int SS, bool NOWRAP>
__device__ __forceinline__ FmCarry fm_make(uint8_t const *src, uint32_t *tile, unsigned long long t0, int nv_tile,
        unsigned long long fm_at, int y_fm, int xf_prev, unsigned flip, long long fa1, long long fb0)
{
    using Cfg = TileCfg<SS>;
    constexpr int C = Cfg::C;
    constexpr int T = Cfg::T;
    constexpr int W = Cfg::W;
    constexpr int SPL = 16 / SS;
    constexpr int NQ = T / (32 * SPL);
    int const lane = threadIdx.x & 31;
    uint32_t *mine = tile + lane * (W * C + 1);
    auto step16 = [](int y, int ca, int cb, int xsum) { return NOWRAP ? iir16_nowrap(y, ca, cb, xsum) : iir16(y, ca, cb, xsum); };

    auto fm_tile = [&](unsigned long long tt, int nvt, bool current, bool unknown, int &cy, int &cf) -> bool {
        int pri = 0, prq = 0; // IQ in front of lane 0's group
        if (tt > 0) {
            uint8_t const *g = src + (tt - 1) * SS;
            if (SS == 2) {
                pri = (int)(g[0] ^ (flip & 0xff)) - 128;
                prq = (int)(g[1] ^ (flip & 0xff)) - 128;
            } else {
                uint32_t w = *reinterpret_cast<uint32_t const *>(g);
                pri = (int)(int16_t)(w & 0xffff);
                prq = (int)(int16_t)(w >> 16);
            }
        }
#pragma unroll 1
        for (int q = 0; q < NQ; ++q) {
            int const n0 = q * 32 * SPL + lane * SPL;
            uint32_t rw[4];
            load_group<SS>(src, tt, n0, nvt, flip, rw);
            int li, lq; // last sample of this lane's group, for the lane to the right
            if (SS == 2) {
                li = (int)((rw[3] >> 16) & 0xff) - 128;
                lq = (int)((rw[3] >> 24) & 0xff) - 128;
            } else {
                li = (int)(int16_t)(rw[3] & 0xffff);
                lq = (int)(int16_t)(rw[3] >> 16);
            }
            int pi_ = __shfl_up_sync(0xffffffffu, li, 1);
            int pq_ = __shfl_up_sync(0xffffffffu, lq, 1);
            if (lane == 0) {
                pi_ = pri;
                pq_ = prq;
            }
            pri = __shfl_sync(0xffffffffu, li, 31);
            prq = __shfl_sync(0xffffffffu, lq, 31);
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                int ci, cq, xf;
                if (SS == 2) {
                    uint32_t w = rw[j >> 1];
                    ci = (int)((w >> ((j & 1) * 16)) & 0xff) - 128;
                    cq = (int)((w >> ((j & 1) * 16 + 8)) & 0xff) - 128;
                    xf = atan16(cq * pi_ - ci * pq_, ci * pi_ + cq * pq_);
                } else {
                    uint32_t w = rw[j];
                    ci = (int)(int16_t)(w & 0xffff);
                    cq = (int)(int16_t)(w >> 16);
                    long long re = (long long)ci * pi_ + (long long)cq * pq_;
                    long long im = (long long)cq * pi_ - (long long)ci * pq_;
                    xf = atan32((int)(unsigned)(unsigned long long)im, (int)(unsigned)(unsigned long long)re);
                }
                pi_ = ci;
                pq_ = cq;
                int at = word_index<C, W>(n0 + j);
                if (W == 1)
                    reinterpret_cast<uint16_t *>(tile)[2 * at + 1] = (uint16_t)xf;
                else
                    tile[at + 1] = (uint32_t)xf;
            }
        }
        __syncwarp();

        auto xf_at = [&](uint32_t const *w) { return W == 1 ? (int)(int16_t)(w[0] >> 16) : (int)w[1]; };
        int nvl = nvt - lane * C;
        nvl = nvl < 0 ? 0 : (nvl > C ? C : nvl);
        int kb = 0;
        int fl; // discriminator output of the sample in front of the chunk
        if (lane == 0) {
            fl = cf;
            if (unknown) {
                kb = 1;
                fl = xf_at(mine);
            }
        } else {
            fl = xf_at(mine - 1 - W);
        }
        int lo_f, hi_f;
        if (lane == 0 && !unknown) {
            lo_f = hi_f = cy;
        } else {
            lo_f = SS == 2 ? -32768 : (int)0x80000000;
            hi_f = SS == 2 ? 32767 : 0x7fffffff;
        }
        bool ok = false;
        int yend = 0;
#pragma unroll 1
        for (int round = 0; round < 32; ++round) {
            if (current) {
                // lanes 0..round are exact by induction even if the filter could wrap
                bool trust = NOWRAP ? (lo_f == hi_f) : (lane <= round);
                if (__all_sync(0xffffffffu, trust)) {
                    ok = true;
                    break;
                }
            }
            int e0 = lo_f, e1 = hi_f;
            int fp = fl;
R4_UNROLL(R4_UNROLL_IIR)
            for (int k = kb; k < nvl; ++k) {
                int v = xf_at(mine + k * W);
                if (SS == 2) {
                    int fsum = v + fp;
                    e0 = step16(e0, (int)fa1, (int)fb0, fsum);
                    e1 = step16(e1, (int)fa1, (int)fb0, fsum);
                } else {
                    long long fsum = (long long)v + fp;
                    e0 = iir32(e0, fa1, fb0, fsum);
                    e1 = iir32(e1, fa1, fb0, fsum);
                }
                fp = v;
            }
            if (!current) {
                int z0 = __shfl_sync(0xffffffffu, e0, 31);
                int z1 = __shfl_sync(0xffffffffu, e1, 31);
                if (z0 == z1) {
                    yend = z0;
                    ok = true;
                    break;
                }
            }
            int n0 = __shfl_up_sync(0xffffffffu, e0, 1);
            int n1 = __shfl_up_sync(0xffffffffu, e1, 1);
            if (lane != 0) {
                lo_f = n0;
                hi_f = n1;
            }
        }
        if (!current) {
            if (ok) {
                cy = yend;
                cf = xf_at(tile + 31 * (W * C + 1) + (C - 1) * W);
            }
            __syncwarp();
            return ok;
        }
        // final pass of the current tile: FM next to AM
        int yf = lo_f;
        int fp = fl;
R4_UNROLL(R4_UNROLL_IIR)
        for (int k = 0; k < nvl; ++k) {
            int v = xf_at(mine + k * W);
            int fo;
            if (SS == 2) {
                yf = step16(yf, (int)fa1, (int)fb0, v + fp);
                fo = yf;
            } else {
                yf = iir32(yf, fa1, fb0, (long long)v + fp);
                fo = yf >> 16;
            }
            fp = v;
            reinterpret_cast<uint16_t *>(mine)[2 * k * W + 1] = (uint16_t)(int16_t)fo;
        }
        int last_lane = (nvt - 1) / C;
        cy = __shfl_sync(0xffffffffu, yf, last_lane);
        cf = __shfl_sync(0xffffffffu, fp, last_lane);
        __syncwarp();
        return true;
    };

    int cy = y_fm, cf = xf_prev;
    unsigned long long tt = fm_at == t0 ? t0 : t0 - T;
    bool unknown = tt != fm_at;
    for (;;) {
        bool const current = tt == t0;
        bool ok = fm_tile(tt, current ? nv_tile : T, current, unknown, cy, cf);
        if (!ok) {
            cy = y_fm;
            cf = xf_prev;
            tt = fm_at;
            unknown = false;
            continue;
        }
        unknown = false;
        if (current) break;
        tt += T;
    }
    FmCarry r;
    r.y = cy;
    r.xf = cf;
    return r;
}

template <int SS, bool NOWRAP>
__global__ void __launch_bounds__(kDetectWarps * 32, kDetectCtasPerSm) k_detect(DetectParams p)
{
    using Cfg = TileCfg<SS>;
    constexpr int C = Cfg::C;
    constexpr int T = Cfg::T;
    constexpr int W = Cfg::W;
    extern __shared__ __align__(16) uint32_t smem[];

    int const warp = threadIdx.x >> 5;
    int const lane = threadIdx.x & 31;
    unsigned const s = p.stream0 + blockIdx.x * kDetectWarps + warp;
    if (s >= p.stream_end) return;

    uint32_t *tile = smem + warp * Cfg::kTileWords;
    __shared__ DetState s_park[kDetectWarps];
    DetState *park = &s_park[warp];
    bool const fm_on = p.enable_fm != 0;

    unsigned long long const byte0 = p.offsets[s];
    unsigned long long const N = (p.lengths ? p.lengths[s] : p.offsets[s + 1] - byte0) / SS;

    Trains tr;
    tr.ook_pulse = p.train_scratch + (size_t)s * kTrainInts;
    tr.ook_gap = tr.ook_pulse + kMaxPulses;
    tr.fsk_pulse = tr.ook_gap + kMaxPulses;
    tr.fsk_gap = tr.fsk_pulse + kMaxPulses;

    WarpCtx cx;
    cx.lane = lane;
    cx.nlanes = 32;

    DetState d;
    unsigned seq = 0;
    int const per_ms = (int)(p.rate / 1000);

    // carried filter / demod state (reset_sdr_flow(): all zero)
    int y_am = 0, y_fm = 0;
    int x_prev = 0;          // raw envelope of the previous sample
    int xf_prev = 0;         // discriminator output of the sample in front of tile `fm_at`
    unsigned fm_tile = 0;    // index of the tile for which (y_fm, xf_prev) are the exact carry-in
    int flushed = 0;
    if (p.first_chunk) {
        det_reset(d);
        d.ook_hw = d.fsk_hw = kMaxPulses; // scratch is not assumed to be zero: first package clears it
    } else {
        StreamState const &ss = p.state[s];
        d = ss.d;
        y_am = ss.y_am;
        y_fm = ss.y_fm;
        x_prev = ss.x_prev;
        xf_prev = ss.xf_prev;
        fm_tile = ss.fm_tile;
        seq = ss.seq;
        flushed = ss.flushed;
    }
    // The discriminator and its low-pass are only looked at inside packages (PULSE, and GAP_START
    // of the first pulse: src/pulse_detect.c:362-365, :367-383).  With a filter that cannot wrap
    // they are therefore computed per tile ON DEMAND; see make_fm below.
    bool const lazy_fm = NOWRAP && fm_on && p.lazy_fm && !p.am_out;

    auto emit = [&](int type, unsigned long long pos, bool flush) {
        PackageHeader h = package_header(d, type);
        unsigned cnt = h.num_pulses + 1 < (unsigned)kMaxPulses ? h.num_pulses + 1 : (unsigned)kMaxPulses;
        unsigned idx = 0, off = 0;
        if (lane == 0) {
            idx = atomicAdd(&p.counters[0], 1u);
            off = atomicAdd(&p.counters[1], cnt);
        }
        idx = __shfl_sync(0xffffffffu, idx, 0);
        off = __shfl_sync(0xffffffffu, off, 0);
        bool fits = idx < p.pkg_cap && (unsigned long long)off + cnt <= p.pool_cap;
        if (!fits) {
            if (lane == 0) atomicOr(&p.counters[2], 1u);
        } else {
            __syncwarp();
            int const *sp = type == 1 ? tr.ook_pulse : tr.fsk_pulse;
            int const *sg = type == 1 ? tr.ook_gap : tr.fsk_gap;
            for (unsigned i = lane; i < cnt; i += 32) {
                p.pulse_pool[off + i] = sp[i];
                p.gap_pool[off + i] = sg[i];
            }
            if (lane == 0) {
                unsigned long long blk = flush ? (N + p.block_samples - 1) / p.block_samples : pos / p.block_samples;
                unsigned long long bstart = blk * p.block_samples;
                unsigned long long blen = flush ? 0 : (N - bstart < p.block_samples ? N - bstart : p.block_samples);
                r433b_package k;
                k.stream = s;
                k.seq = seq;
                k.type = type;
                k.block = (int)blk;
                k.offset = h.offset;
                k.end_pos = pos;
                k.start_ago = flush ? (unsigned)(N - h.start_abs) : (unsigned)(bstart + blen - h.start_abs);
                k.end_ago = flush ? 0u : (unsigned)(blen - (pos - bstart));
                k.num_pulses = h.num_pulses;
                k.pulse_off = off;
                k.pulse_count = cnt;
                k.ook_low_estimate = h.low;
                k.ook_high_estimate = h.high;
                k.fsk_f1_est = h.f1;
                k.fsk_f2_est = h.f2;
                k.first_pair = 0;
                p.pkgs[idx] = k;
            }
        }
        seq++;
    };

    for (unsigned long long t0 = p.sample_begin; t0 < p.sample_end && t0 < N; t0 += T) {
        uint8_t const *src = p.data + p.offsets[s]; // re-derived per tile: not worth two registers across the walk
        unsigned long long const remain = N - t0;
        int const nv_tile = remain < (unsigned long long)T ? (int)remain : T;
        int nv = nv_tile - lane * C; // valid samples in this lane's chunk
        nv = nv < 0 ? 0 : (nv > C ? C : nv);

        // ---- the tile front: sample maps + exact low-pass(es) --------------------------------
        // Two instances of the same code.  front(true): envelope AND discriminator, both filters
        // in the same loops (four independent dependency chains per lane) -- used while a package
        // is open (and always when FM cannot be deferred).  front(false): envelope and AM filter
        // only -- used while the detector is IDLE; if a package starts inside such a tile,
        // fm_make() supplies FM for it afterwards.
        constexpr int SPL = 16 / SS;          // samples per 128-bit load
        constexpr int NQ = T / (32 * SPL);    // load iterations per tile (4)
        uint32_t *mine = tile + lane * (W * C + 1); // lane l owns samples [l*C, l*C + C) of the tile; nv of them exist
        auto step16 = [](int y, int ca, int cb, int xsum) { return NOWRAP ? iir16_nowrap(y, ca, cb, xsum) : iir16(y, ca, cb, xsum); };
        auto front = [&](auto with_fm) {
            constexpr bool FM = decltype(with_fm)::value;
            // phase 1: maps, coalesced.  Iteration q: lane l takes the 16 contiguous bytes (8 cu8 /
            // 4 cs16 samples) at q*512 + l*16 of the tile: one 128-bit load per lane.
            int pri = 0, prq = 0; // IQ in front of lane 0's group
            if (FM && t0 > 0) {
                uint8_t const *g = src + (t0 - 1) * SS;
                if (SS == 2) {
                    pri = (int)(g[0] ^ (p.flip & 0xff)) - 128;
                    prq = (int)(g[1] ^ (p.flip & 0xff)) - 128;
                } else {
                    uint32_t w = *reinterpret_cast<uint32_t const *>(g);
                    pri = (int)(int16_t)(w & 0xffff);
                    prq = (int)(int16_t)(w >> 16);
                }
            }
#pragma unroll 1
            for (int q = 0; q < NQ; ++q) {
                int const n0 = q * 32 * SPL + lane * SPL;
                uint32_t rw[4];
                load_group<SS>(src, t0, n0, nv_tile, p.flip, rw);
                int pi_ = 0, pq_ = 0;
                if (FM) {
                    int li, lq; // last sample of this lane's group, for the lane to the right
                    if (SS == 2) {
                        li = (int)((rw[3] >> 16) & 0xff) - 128;
                        lq = (int)((rw[3] >> 24) & 0xff) - 128;
                    } else {
                        li = (int)(int16_t)(rw[3] & 0xffff);
                        lq = (int)(int16_t)(rw[3] >> 16);
                    }
                    pi_ = __shfl_up_sync(0xffffffffu, li, 1);
                    pq_ = __shfl_up_sync(0xffffffffu, lq, 1);
                    if (lane == 0) {
                        pi_ = pri;
                        pq_ = prq;
                    }
                    pri = __shfl_sync(0xffffffffu, li, 31);
                    prq = __shfl_sync(0xffffffffu, lq, 31);
                }
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    int ci, cq, x, xf = 0;
                    if (SS == 2) {
                        uint32_t w = rw[j >> 1];
                        int ri = (int)((w >> ((j & 1) * 16)) & 0xff);
                        int rq = (int)((w >> ((j & 1) * 16 + 8)) & 0xff);
                        ci = ri - 128;
                        cq = rq - 128;
                        x = p.use_mag ? mag_cu8(ri, rq) : env_cu8(ri, rq);
                    } else {
                        uint32_t w = rw[j];
                        ci = (int)(int16_t)(w & 0xffff);
                        cq = (int)(int16_t)(w >> 16);
                        x = mag_cs16(ci, cq);
                    }
                    if (FM) {
                        if (SS == 2) {
                            xf = atan16(cq * pi_ - ci * pq_, ci * pi_ + cq * pq_);
                        } else {
                            long long re = (long long)ci * pi_ + (long long)cq * pq_;
                            long long im = (long long)cq * pi_ - (long long)ci * pq_;
                            xf = atan32((int)(unsigned)(unsigned long long)im, (int)(unsigned)(unsigned long long)re);
                        }
                        pi_ = ci;
                        pq_ = cq;
                    }
                    int at = word_index<C, W>(n0 + j);
                    if (W == 1) {
                        tile[at] = (uint32_t)x | ((uint32_t)xf << 16);
                    } else {
                        tile[at] = (uint32_t)x;
                        if (FM) tile[at + 1] = (uint32_t)xf;
                    }
                }
            }
            __syncwarp();

            // phase 2: the low-pass(es), exact and lane-parallel
            int xl, fl = 0; // envelope / discriminator of the sample in front of the chunk
            if (lane == 0) {
                // the reference keeps x[-1] as int16 across block calls (src/baseband.c:167)
                xl = (t0 % p.block_samples == 0) ? (int)(int16_t)x_prev : x_prev;
                fl = xf_prev;
            } else {
                uint32_t const *left = mine - 1 - W; // last sample of the lane to the left
                xl = (int)(left[0] & 0xffff);
                if (FM) fl = W == 1 ? (int)(int16_t)(left[0] >> 16) : (int)left[1];
            }
            int const a1 = p.lpf_a1, b0 = p.lpf_b0;
            long long const fa1 = p.fm_a1, fb0 = p.fm_b0;
            int lo_a, hi_a, lo_f, hi_f;
            if (lane == 0) {
                lo_a = hi_a = y_am;
                lo_f = hi_f = y_fm;
            } else {
                lo_a = -32768;
                hi_a = 32767;
                lo_f = SS == 2 ? -32768 : (int)0x80000000;
                hi_f = SS == 2 ? 32767 : 0x7fffffff;
            }
#pragma unroll 1
            for (int round = 0; round < 31; ++round) {
                bool mine_ok = (lo_a == hi_a) && (!FM || lo_f == hi_f);
                // lanes 0..round are exact by induction even if the filter could wrap
                bool trust = NOWRAP ? mine_ok : (lane <= round);
                if (__all_sync(0xffffffffu, trust)) break;
                // both ends of every bracket advance together: independent dependency chains
                int ea_lo = lo_a, ea_hi = hi_a, ef_lo = lo_f, ef_hi = hi_f;
                int xp = xl, fp = fl;
R4_UNROLL(R4_UNROLL_IIR)
                for (int k = 0; k < nv; ++k) {
                    uint32_t w0 = mine[k * W];
                    int x = (int)(w0 & 0xffff);
                    int xsum = x + xp;
                    ea_lo = step16(ea_lo, a1, b0, xsum);
                    ea_hi = step16(ea_hi, a1, b0, xsum);
                    xp = x;
                    if (FM) {
                        if (SS == 2) {
                            int v = (int)(int16_t)(w0 >> 16);
                            int fsum = v + fp;
                            ef_lo = step16(ef_lo, (int)fa1, (int)fb0, fsum);
                            ef_hi = step16(ef_hi, (int)fa1, (int)fb0, fsum);
                            fp = v;
                        } else {
                            int v = (int)mine[k * W + 1];
                            long long fsum = (long long)v + fp;
                            ef_lo = iir32(ef_lo, fa1, fb0, fsum);
                            ef_hi = iir32(ef_hi, fa1, fb0, fsum);
                            fp = v;
                        }
                    }
                }
                int na_lo = __shfl_up_sync(0xffffffffu, ea_lo, 1);
                int na_hi = __shfl_up_sync(0xffffffffu, ea_hi, 1);
                if (lane != 0) {
                    lo_a = na_lo;
                    hi_a = na_hi;
                }
                if (FM) {
                    int nf_lo = __shfl_up_sync(0xffffffffu, ef_lo, 1);
                    int nf_hi = __shfl_up_sync(0xffffffffu, ef_hi, 1);
                    if (lane != 0) {
                        lo_f = nf_lo;
                        hi_f = nf_hi;
                    }
                }
            }

            // phase 3: final pass from the exact state; AM|FM replace x|xf in place
            int ya = lo_a, yf = lo_f;
            int xp = xl, fp = fl;
            unsigned long long const gbase = p.am_out ? p.offsets[s] / SS + t0 + (unsigned long long)lane * C : 0;
R4_UNROLL(R4_UNROLL_IIR)
            for (int k = 0; k < nv; ++k) {
                uint32_t w0 = mine[k * W];
                int x = (int)(w0 & 0xffff);
                ya = step16(ya, a1, b0, x + xp);
                xp = x;
                int fo;
                if (FM) {
                    if (SS == 2) {
                        int v = (int)(int16_t)(w0 >> 16);
                        yf = step16(yf, (int)fa1, (int)fb0, v + fp);
                        fp = v;
                        fo = yf;
                    } else {
                        int v = (int)mine[k * W + 1];
                        yf = iir32(yf, fa1, fb0, (long long)v + fp);
                        fp = v;
                        fo = yf >> 16;
                    }
                } else {
                    fo = fm_on ? 0 : (int)(int16_t)x; // buf.fm aliases the raw envelope when FM is off
                }
                mine[k * W] = (uint32_t)(uint16_t)(int16_t)ya | ((uint32_t)(uint16_t)(int16_t)fo << 16);
                if (p.am_out) {
                    p.am_out[gbase + k] = (int16_t)ya;
                    p.fm_out[gbase + k] = (int16_t)fo;
                }
            }
            // carries for the next tile: the state after the last valid sample
            int last_lane = (nv_tile - 1) / C;
            y_am = __shfl_sync(0xffffffffu, ya, last_lane);
            x_prev = __shfl_sync(0xffffffffu, xp, last_lane);
            if (FM) {
                y_fm = __shfl_sync(0xffffffffu, yf, last_lane);
                xf_prev = __shfl_sync(0xffffffffu, fp, last_lane);
                fm_tile = (unsigned)(t0 / T) + 1;
            }
        };
        // d.st != IDLE at a tile start implies FM was made for the previous tile
        bool const fused = fm_on && (!lazy_fm || (d.st != kIdle && fm_tile == (unsigned)(t0 / T)));
        // The (warp-uniform) detector state is not needed by the front: park it in shared memory
        // so that the filter loops have the registers (otherwise ptxas spills inside them).
        if (lane == 0) *park = d;
        __syncwarp();
        if (fused)
            front(std::true_type{});
        else
            front(std::false_type{});
        __syncwarp();
        d = *park;

        bool fm_ready = !fm_on || fused;

        // ---- package detector over the tile (warp-uniform) -------------------------------
        if (t0 % p.block_samples == 0) det_call_boundary(d, p.lv);

        // Warp-cooperative fast paths.  Each looks at up to 32 consecutive samples (one per
        // lane), proves with a ballot that the state machine stays on one simple trajectory for
        // a prefix of them, and advances the (warp-uniform) state over that prefix at once.
        // They return the number of samples consumed; 0 hands the current sample to det_step().

        // IDLE over a long stretch, lane-parallel.  While |am - low| < 1024 the tracker is
        // low += (am > low) ? +1 : -1, so low keeps the parity of (low0 + samples seen) and two
        // trajectories of equal parity never cross and merge once the data passes between them.
        // That is the IIR trick again: lane l takes chunk l of the tile, starts from a bracket
        // [lo, hi] of the right parity that provably contains the true value (low can never
        // leave [min(low0, min(am)-1), max(low0, max(am))]), pushes both ends through its chunk
        // and hands them to the next lane until every bracket has collapsed.  Chunks in which a
        // trigger is conceivable (or |am - low| could reach 1024) end the stretch; they are left
        // to the 32-sample path below.
        auto idle_tile = [&](int n) -> int {
            if (nv_tile - n < 3 * C) return 0;
            {
                int hs = p.lv.ratio * d.low;
                if (hs < p.lv.min_high) hs = p.lv.min_high;
                if (d.high != hs) return 0;
            }
            int const c0 = n / C;
            int const k0 = lane == c0 ? n - c0 * C : 0;
            int k1 = nv_tile - lane * C;
            k1 = k1 > C ? C : k1;
            bool const in_region = lane >= c0 && k1 > k0;
            uint32_t const *chunk = tile + lane * (W * C + 1);
            int cmin = 32767, cmax = -32768;
            if (in_region) {
R4_UNROLL(R4_UNROLL_TILE)
                for (int k = k0; k < k1; ++k) {
                    int a = (int)(int16_t)(chunk[k * W] & 0xffff);
                    cmin = a < cmin ? a : cmin;
                    cmax = a > cmax ? a : cmax;
                }
            }
            int pmin = cmin, pmax = cmax; // over chunks c0..lane
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t1 = __shfl_up_sync(0xffffffffu, pmin, o);
                int t2 = __shfl_up_sync(0xffffffffu, pmax, o);
                if (lane >= o) {
                    pmin = t1 < pmin ? t1 : pmin;
                    pmax = t2 > pmax ? t2 : pmax;
                }
            }
            int Lmin = d.low < pmin - 1 ? d.low : pmin - 1;
            int Lmax = d.low > pmax ? d.low : pmax;
            int hmin = p.lv.ratio * Lmin;
            if (hmin < p.lv.min_high) hmin = p.lv.min_high;
            Thresholds th = det_thresholds(Lmin, hmin, p.lv);
            bool const armed = d.lead_in + (nv_tile - n) > kLeadIn;
            bool ok = in_region && !(armed && cmax > th.up) && (pmax - Lmin < 1024) && (Lmax - pmin < 1024);
            unsigned bad = ~__ballot_sync(0xffffffffu, ok) & (0xffffffffu << c0);
            int const e = bad ? __ffs(bad) - 1 : 32; // chunks c0 .. e-1 form the stretch
            if (e - c0 < 3) return 0;
            int const RLmin = __shfl_sync(0xffffffffu, Lmin, e - 1);
            int const RLmax = __shfl_sync(0xffffffffu, Lmax, e - 1);
            bool const act = lane >= c0 && lane < e;
            int const par = (d.low + (lane * C + k0 - n)) & 1; // parity of the true value at this lane's start
            // Start bracket.  Over K samples whose values lie in [m, M] the tracker climbs one per
            // sample until it is >= m - 1 and falls one per sample until it is <= M, so from any
            // start in [A, B] it ends in [min(A + K, m - 1), max(B - K, M)].  Use the two chunks to
            // the left (the one chunk and the exact start value for the second lane of the stretch).
            int m1 = __shfl_up_sync(0xffffffffu, cmin, 1), M1 = __shfl_up_sync(0xffffffffu, cmax, 1);
            int m2 = __shfl_up_sync(0xffffffffu, cmin, 2), M2 = __shfl_up_sync(0xffffffffu, cmax, 2);
            int K1 = __shfl_up_sync(0xffffffffu, k1 - k0, 1);
            int lo, hi;
            if (lane == c0 + 1) {
                lo = d.low + K1 < m1 - 1 ? d.low + K1 : m1 - 1;
                hi = d.low - K1 > M1 ? d.low - K1 : M1;
            } else {
                int mm = m1 < m2 ? m1 : m2, MM = M1 > M2 ? M1 : M2;
                int Kk = K1 + (lane == c0 + 2 ? 0 : C); // chunk c0 may be partial: count only lane-1 then
                if (lane == c0 + 2) {
                    mm = m1;
                    MM = M1;
                }
                lo = RLmin + Kk < mm - 1 ? RLmin + Kk : mm - 1;
                hi = RLmax - Kk > MM ? RLmax - Kk : MM;
            }
            lo = lo < RLmin ? RLmin : lo;
            hi = hi > RLmax ? RLmax : hi;
            lo -= (lo - par) & 1;
            hi += (hi - par) & 1;
            if (lane == c0) lo = hi = d.low;
            int result = 0;
            bool done = false;
#pragma unroll 1
            for (int round = 0; round < 8; ++round) {
                int elo = lo, ehi = hi;
                if (act) {
R4_UNROLL(R4_UNROLL_TILE)
                    for (int k = k0; k < k1; ++k) {
                        int a = (int)(int16_t)(chunk[k * W] & 0xffff);
                        elo += a > elo ? 1 : -1;
                        ehi += a > ehi ? 1 : -1;
                    }
                }
                if (__all_sync(0xffffffffu, !act || elo == ehi)) {
                    result = __shfl_sync(0xffffffffu, elo, e - 1);
                    done = true;
                    break;
                }
                int nlo = __shfl_up_sync(0xffffffffu, elo, 1);
                int nhi = __shfl_up_sync(0xffffffffu, ehi, 1);
                if (act && lane != c0) {
                    lo = nlo;
                    hi = nhi;
                }
            }
            if (!done) return 0;
            int const len = (e * C < nv_tile ? e * C : nv_tile) - n;
            d.low = result;
            int hh = p.lv.ratio * d.low;
            d.high = hh < p.lv.min_high ? p.lv.min_high : hh;
            int li = d.lead_in + len;
            d.lead_in = li > kLeadIn + 1 ? kLeadIn + 1 : li;
            return len;
        };

        // IDLE: only the noise-floor tracker moves (src/pulse_detect.c:325-334).  While
        // |am - low| < 1024 it is low += (am > low) ? +1 : -1; with q = low + j that is
        // q += 2 * (am_j + j > q): two dependent instructions per sample.
        auto idle_fast = [&](int n) -> int {
            int cnt = nv_tile - n < 32 ? nv_tile - n : 32;
            int hs = p.lv.ratio * d.low;
            if (hs < p.lv.min_high) hs = p.lv.min_high;
            if (d.high != hs) return 0; // first IDLE sample after a package: not yet re-derived
            int a = lane < cnt ? (int)(int16_t)(tile[word_index<C, W>(n + lane)] & 0xffff) : -32768;
            int lmin = d.low - cnt;
            int hmin = p.lv.ratio * lmin;
            if (hmin < p.lv.min_high) hmin = p.lv.min_high;
            Thresholds th = det_thresholds(lmin, hmin, p.lv); // lowest trigger level reachable in this chunk
            bool armed = d.lead_in + cnt - 1 > kLeadIn;
            bool stop = lane < cnt && ((armed && a > th.up) || (a - lmin >= 1024) || (d.low + cnt - a >= 1024));
            unsigned m = __ballot_sync(0xffffffffu, stop);
            if (m) {
                int first = __ffs(m) - 1;
                cnt = first < cnt ? first : cnt;
            }
            if (cnt == 0) return 0;
            int q = d.low;
            int b = a + lane;
R4_UNROLL(R4_UNROLL_IDLE)
            for (int j = 0; j < cnt; ++j) {
                int bj = __shfl_sync(0xffffffffu, b, j);
                if (bj > q) q += 2;
            }
            d.low = q - cnt;
            int hh = p.lv.ratio * d.low;
            d.high = hh < p.lv.min_high ? p.lv.min_high : hh;
            int li = d.lead_in + cnt;
            d.lead_in = li > kLeadIn + 1 ? kLeadIn + 1 : li;
            return cnt;
        };

        // GAP: thresholds are frozen; the next event is the first sample above `up` or the run
        // length reaching an end-of-package limit (src/pulse_detect.c:422-470).  Look for either in
        // the rest of the tile, 32 samples per ballot.
        auto gap_fast = [&](int n) -> int {
            if (d.eop_flag) return 0;
            int const cnt = nv_tile - n;
            Thresholds th = det_thresholds(d.low, d.high, p.lv);
            long long lim_a = 10ll * d.longest > 10ll * per_ms ? 10ll * d.longest : 10ll * per_ms;
            long long lim_b = 100ll * per_ms;
            long long rstar = (lim_a < lim_b ? lim_a : lim_b) + 1; // first run length that ends the package
            long long je = rstar - d.run - 1;
            if (je < 0) je = 0;
            int const horizon = je < cnt ? (int)je + 1 : cnt; // samples that matter
            int ja = 0x7fffffff;
#pragma unroll 1
            for (int base = 0; base < horizon; base += 32) {
                int a = base + lane < cnt ? (int)(int16_t)(tile[word_index<C, W>(n + base + lane)] & 0xffff) : -32768;
                unsigned m = __ballot_sync(0xffffffffu, a > th.up);
                if (m) {
                    ja = base + __ffs(m) - 1;
                    break;
                }
            }
            if (ja < cnt && ja <= je) { // a new pulse starts first
                d.run += ja + 1;
                put(tr.ook_gap, d.ook_hw, d.ook_n, d.run);
                d.ook_n += 1;
                if (d.ook_n >= (unsigned)kMaxPulses) {
                    d.st = kIdle;
                    emit(1, t0 + n + ja, false);
                    det_call_boundary(d, p.lv);
                    return ja; // that sample is looked at again in IDLE
                }
                d.run = 0;
                d.st = kPulse;
                return ja + 1;
            }
            if (je < cnt) { // end of package by gap length
                d.run += (int)je + 1;
                put(tr.ook_gap, d.ook_hw, d.ook_n, d.run);
                d.ook_n += 1;
                d.st = kIdle;
                emit(1, t0 + n + (int)je, false);
                det_call_boundary(d, p.lv);
                return (int)je;
            }
            d.run += cnt;
            return cnt;
        };

        // PULSE after the first pulse (no FSK sub-detector): the high-level and carrier estimators
        // (src/pulse_detect.c:362-365) are 64-sample moving averages with truncation -- inherently
        // sequential -- but the pulse only ends on a sample below the threshold their value implies.
        // One step never lifts `high` above max(high, 64 * (am / 64) + 63), so the largest am of the
        // chunk bounds every threshold of the chunk from above: samples not below THAT threshold
        // cannot end the pulse.  Advance the two recurrences over exactly those samples; the first
        // sample that might end the pulse is left to det_step(), which tests it exactly.
        auto pulse_fast = [&](int n) -> int {
            if (d.ook_n == 0) return 0;
            int cnt = nv_tile - n < 32 ? nv_tile - n : 32;
            uint32_t wv = lane < cnt ? tile[word_index<C, W>(n + lane)] : 0x00007fffu;
            int a = (int)(int16_t)(wv & 0xffff);
            int f = (int)(int16_t)(wv >> 16);
            int aq = a / 64, fq = f / 64;
            int top = lane < cnt ? aq : -512;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                int t = __shfl_xor_sync(0xffffffffu, top, o);
                top = t > top ? t : top;
            }
            int hmax = 64 * top + 63;
            hmax = d.high > hmax ? d.high : hmax;
            Thresholds th = det_thresholds(d.low, hmax, p.lv);
            unsigned m = __ballot_sync(0xffffffffu, lane < cnt && a < th.down);
            if (m) cnt = __ffs(m) - 1;
            if (cnt == 0) return 0;
            int h = d.high, g = d.ook_f1; // h >= min_high >= 0 here, so h / 64 == h >> 6
            int const minh = p.lv.min_high;
R4_UNROLL(R4_UNROLL_PULSE)
            for (int j = 0; j < cnt; ++j) {
                int aj = __shfl_sync(0xffffffffu, aq, j);
                int fj = __shfl_sync(0xffffffffu, fq, j);
                h += aj - (int)((unsigned)h >> 6);
                h = h < minh ? minh : h;
                g += fj - g / 64;
            }
            d.high = h;
            d.ook_f1 = g;
            d.run += cnt;
            return cnt;
        };

        // PULSE of the FIRST pulse: the same bound, with the FSK sub-detector fed in the loop
        // (src/pulse_detect.c:367-371).  An FSK transmission is one long OOK "pulse", so this is the
        // hot loop of FSK captures; det_step() would re-derive thresholds and re-dispatch per sample.
        auto pulse0_fast = [&](int n) -> int {
            int cnt = nv_tile - n < 32 ? nv_tile - n : 32;
            uint32_t wv = lane < cnt ? tile[word_index<C, W>(n + lane)] : 0x00007fffu;
            int a = (int)(int16_t)(wv & 0xffff);
            int f = (int)(int16_t)(wv >> 16);
            int aq = a / 64;
            int top = lane < cnt ? aq : -512;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                int t = __shfl_xor_sync(0xffffffffu, top, o);
                top = t > top ? t : top;
            }
            int hmax = 64 * top + 63;
            hmax = d.high > hmax ? d.high : hmax;
            Thresholds th = det_thresholds(d.low, hmax, p.lv);
            unsigned m = __ballot_sync(0xffffffffu, lane < cnt && a < th.down);
            if (m) cnt = __ffs(m) - 1;
            if (cnt == 0) return 0;
            int const minh = p.lv.min_high;
#pragma unroll 1
            for (int j = 0; j < cnt; ++j) {
                int aj = __shfl_sync(0xffffffffu, aq, j);
                int fj = __shfl_sync(0xffffffffu, f, j);
                d.high += aj - (int)((unsigned)d.high >> 6);
                d.high = d.high < minh ? minh : d.high;
                d.ook_f1 += fj / 64 - d.ook_f1 / 64;
                if (p.fpdm == 0)
                    fsk_classic(d, tr, fj, cx);
                else
                    fsk_minmax(d, tr, fj, cx);
            }
            d.run += cnt;
            return cnt;
        };

        // GAP_START after the first pulse (no FSK feed): thresholds are frozen and nothing happens
        // until either a sample rises above `up` (spurious gap) or the run reaches 10 samples.
        // Skip the uneventful samples in front of that transition; det_step() takes the transition.
        auto gapstart_fast = [&](int n) -> int {
            if (d.ook_n == 0) return 0;
            int quiet = kMinPulseSamples - 1 - d.run; // samples that can pass without reaching 10
            if (quiet <= 0) return 0;
            int cnt = nv_tile - n < quiet ? nv_tile - n : quiet;
            Thresholds th = det_thresholds(d.low, d.high, p.lv);
            int a = lane < cnt ? (int)(int16_t)(tile[word_index<C, W>(n + lane)] & 0xffff) : -32768;
            unsigned m = __ballot_sync(0xffffffffu, lane < cnt && a > th.up);
            if (m) {
                int ja = __ffs(m) - 1;
                cnt = ja < cnt ? ja : cnt;
            }
            d.run += cnt;
            return cnt;
        };

        // The detector walks the tile until it is done or first needs FM; in that case FM is made
        // out here (not inside the walk: the walk keeps its registers) and the walk resumes.
        int n = 0;
        for (;;) {
        for (; n < nv_tile;) {
            if (!fm_ready && (d.st == kPulse || (d.st == kGapStart && d.ook_n == 0))) break;
            int adv = 0;
            if (d.st == kIdle) {
                adv = idle_tile(n);
                if (!adv) adv = idle_fast(n);
            }
            else if (d.st == kGap)
                adv = gap_fast(n);
            else if (d.st == kPulse)
                adv = d.ook_n ? pulse_fast(n) : pulse0_fast(n);
            else
                adv = gapstart_fast(n);
            if (adv) {
                n += adv;
                continue;
            }
            uint32_t wv = tile[word_index<C, W>(n)];
            int a = (int)(int16_t)(wv & 0xffff);
            int f = (int)(int16_t)(wv >> 16);
            int ev = det_step(d, p.lv, tr, a, f, t0 + n, per_ms, p.fpdm, cx);
            if (ev) {
                emit(ev, t0 + n, false);
                det_call_boundary(d, p.lv);
                continue; // the same sample is examined again, now in IDLE
            }
            ++n;
        }
        if (n >= nv_tile) break;
        {
            if (lane == 0) *park = d;
            __syncwarp();
            FmCarry c = fm_make<SS, NOWRAP>(src, tile, t0, nv_tile, (unsigned long long)fm_tile * T, y_fm, xf_prev, p.flip,
                    p.fm_a1, p.fm_b0);
            y_fm = c.y;
            xf_prev = c.xf;
            fm_tile = (unsigned)(t0 / T) + 1;
            fm_ready = true;
            d = *park;
        }
        } // walk / make FM / resume
        __syncwarp();
    }

    // flush_sdr_flow(): len == 0 call(s) at the end of the file, in the launch that reaches it
    if (N <= p.sample_end && !flushed) {
        for (;;) {
            int ev = det_flush(d, tr, p.fpdm);
            if (!ev) break;
            emit(ev, N, true);
        }
        flushed = 1;
    }
    if (lane == 0 && p.state) {
        StreamState &ss = p.state[s];
        ss.d = d;
        ss.y_am = y_am;
        ss.y_fm = y_fm;
        ss.x_prev = x_prev;
        ss.xf_prev = xf_prev;
        ss.fm_tile = fm_tile;
        ss.seq = seq;
        ss.flushed = flushed;
    }
}

// ----------------------------------------------------------------- cf32 -> cs16 ----------

// src/rtl_433.c:1811-1825: "clamp float to [-1,1] and scale to Q0.15" -- the reference converts a
// cf32 capture to cs16 before anything else looks at it.  Streaming, HBM bound (12 B per float pair
// of traffic per 2 samples); `n4` groups of four floats.  The out-of-range / NaN case follows the
// reference's x86-64 builds: the conversion yields INT_MIN, the clamp makes it -32767.
__global__ void k_cf32_to_cs16(float4 const *in, uint2 *out, size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = __ldcs(in + i);
        float f[4] = {v.x, v.y, v.z, v.w};
        int s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float t = __fmul_rn(f[k], 32767.0f);
            int q = (t >= -2147483648.0f && t < 2147483648.0f) ? __float2int_rz(t) : (int)0x80000000;
            q = q < -32767 ? -32767 : (q > 32767 ? 32767 : q);
            s[k] = q;
        }
        uint2 o;
        o.x = (uint32_t)(uint16_t)(int16_t)s[0] | ((uint32_t)(uint16_t)(int16_t)s[1] << 16);
        o.y = (uint32_t)(uint16_t)(int16_t)s[2] | ((uint32_t)(uint16_t)(int16_t)s[3] << 16);
        out[i] = o;
    }
}

// ------------------------------------------------------------------------- k_slice -------

// What one pipeline group (a contiguous run of streams) produced, filled on the device so the
// next stage never waits for the host.
struct GroupRange {
    unsigned pkg_begin, pkg_end;
    unsigned pool_begin, pool_end;
    unsigned long long arena_begin, arena_end;
    unsigned long long events_end;
    unsigned overflow;
    unsigned next; // k_slice work counter: next (package, device group) item of this range, relative to pkg_begin
};

__global__ void k_mark(GroupRange *r, int which, unsigned const *counters, unsigned long long const *cursor)
{
    if (which == 0) {
        r->pkg_begin = counters[0];
        r->pool_begin = counters[1];
    } else if (which == 1) {
        r->pkg_end = counters[0];
        r->pool_end = counters[1];
        r->overflow = counters[2];
        r->next = 0;
    } else if (which == 2) {
        r->arena_begin = cursor[0];
    } else {
        r->arena_end = cursor[0];
        r->events_end = cursor[1];
        r->overflow |= (unsigned)cursor[2] << 1;
    }
}

struct SliceParams {
    r433b_package *pkgs;
    unsigned n_pkgs;
    GroupRange *range; // packages [pkg_begin, min(pkg_end, n_pkgs)) and the work counter
    int const *pulse_pool, *gap_pool;
    SlicerParams const *dev;  // per device, already scaled to the batch sample rate
    unsigned n_devs;
    unsigned const *ook_list, *fsk_list; // device indices taking OOK / FSK packages, grouped by modulation
    unsigned n_ook, n_fsk;
    r433b_pair *pairs;        // n_pkgs * n_devs, pre-zeroed
    uint8_t *arena;
    unsigned long long arena_cap;
    unsigned long long *cursor; // [0] bytes reserved, [1] events, [2] overflow
    uint32_t *stage;            // stage_words per thread of the (fixed, ranged) grid, or nullptr
    unsigned stage_words;
};

constexpr int kSliceThreads = 128;
constexpr unsigned kStageWords = 1024; // scratch words per k_slice thread (4 KiB): larger outputs take the second pass
constexpr int kSliceCtasPerSm = 12; // 40 registers, 48 warps/SM (measured: 8 -> 13.3 ms, 10 -> 12.4, 12 -> 12.2, 16 -> 13.4)

__global__ void __launch_bounds__(kSliceThreads, kSliceCtasPerSm) k_slice(SliceParams p)
{
    // A fixed grid of independent WARPS.  A work item is (package, group of 32 devices of the list
    // that takes the package type); every warp fetches items from the range's counter until the
    // range is exhausted.  (One CTA per package left most warps waiting at the CTA barrier for the
    // slowest device group: 48 % of all stall samples.)
    unsigned const lane = threadIdx.x & 31;
    unsigned const groups = ((p.n_ook > p.n_fsk ? p.n_ook : p.n_fsk) + 31) / 32;
    if (!groups) return;
    unsigned const pk_begin = p.range->pkg_begin;
    unsigned const pk_end = p.range->pkg_end < p.n_pkgs ? p.range->pkg_end : p.n_pkgs;
    // Staging: the first run of the slicer writes its words into a per-thread scratch region; if
    // every lane's output fitted, the warp reserves arena space and copies the regions over
    // (coalesced, lane by lane) instead of running the slicers a second time.
    uint32_t *stage = p.stage ? p.stage + ((size_t)blockIdx.x * kSliceThreads + threadIdx.x) * p.stage_words : nullptr;
    unsigned const stage_words = stage ? p.stage_words : 0;
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(&p.range->next, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        // device group major: at any moment most warps of the GPU run the same slicer (instruction cache)
        if (pk_end <= pk_begin) break;
        unsigned const n_rel = pk_end - pk_begin;
        unsigned const g = item / n_rel, rel = item - g * n_rel;
        if (g >= groups) break;
        unsigned const pk = pk_begin + rel;
        r433b_package const k = p.pkgs[pk];
        unsigned const n_list = k.type == 1 ? p.n_ook : p.n_fsk;
        unsigned const *list = k.type == 1 ? p.ook_list : p.fsk_list;
        if (g == 0 && lane == 0) p.pkgs[pk].first_pair = pk * p.n_devs;
        if (g * 32 >= n_list) continue;

        PulseView pv;
        pv.pulse = p.pulse_pool + k.pulse_off;
        pv.gap = p.gap_pool + k.pulse_off;
        pv.n = k.num_pulses;

        unsigned const slot = g * 32 + lane;
        unsigned const dev = slot < n_list ? list[slot] : kNoDevice;
        bool const active = dev != kNoDevice;
        unsigned bytes = 0, nev = 0;
        unsigned long long off = 0;
        bool fits = false;
        SlicerParams sp;
        if (active) sp = p.dev[dev];
        // pass 0 counts (and stages), pass 1 stores; one copy of the slicer code serves both
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                // warp-aggregated reservation in the event arena
                unsigned incl = bytes;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((int)lane >= o) incl += v;
                }
                unsigned total = __shfl_sync(0xffffffffu, incl, 31);
                unsigned evs = nev;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) evs += __shfl_xor_sync(0xffffffffu, evs, o);
                unsigned long long wbase = 0;
                if (lane == 0 && total) {
                    wbase = atomicAdd(p.cursor, (unsigned long long)total);
                    atomicAdd(p.cursor + 1, (unsigned long long)evs);
                }
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                off = wbase + incl - bytes;
                fits = off + bytes <= p.arena_cap;
                if (active && bytes && !fits) atomicOr(p.cursor + 2, 1ull);
                bool const staged = bytes <= stage_words * 4;
                if (__all_sync(0xffffffffu, !active || !bytes || staged)) {
                    __syncwarp();
                    unsigned todo = __ballot_sync(0xffffffffu, active && bytes && fits);
                    while (todo) {
                        int const l = __ffs(todo) - 1;
                        todo &= todo - 1;
                        unsigned const wl = __shfl_sync(0xffffffffu, bytes, l) / 4;
                        unsigned long long const ol = __shfl_sync(0xffffffffu, off, l);
                        uint32_t const *from = stage + ((long long)l - (long long)lane) * (long long)stage_words;
                        uint32_t *to = reinterpret_cast<uint32_t *>(p.arena + ol);
                        for (unsigned i = lane; i < wl; i += 32) __stcs(to + i, from[i]);
                    }
                    __syncwarp();
                    break;
                }
            }
            if (active && (pass == 0 || (bytes && fits))) {
                EventWriter w;
                if (pass)
                    w.init(reinterpret_cast<uint32_t *>(p.arena + off), bytes / 4);
                else
                    w.init(stage, stage_words);
                slice_dispatch(pv, sp, w);
                if (pass == 0) {
                    bytes = w.committed * 4;
                    nev = w.events;
                }
            }
        }
        if (active) {
            r433b_pair pr;
            pr.offset = off;
            pr.bytes = bytes;
            pr.events = nev;
            p.pairs[(size_t)pk * p.n_devs + dev] = pr;
        }
    }
}

} // namespace r433b
