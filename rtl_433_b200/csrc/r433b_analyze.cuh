// r433b_analyze.cuh -- the pulse analyzer (`rtl_433 -A`) batched over packages.
//
// src/pulse_analyzer.c:279-560 looks at one package at a time: five width histograms with a 20 % relative
// tolerance (pulses, gaps, pulse+gap periods, gap+pulse periods, all timings), fused, then a guess of the
// modulation from the bin counts, an RfRaw rendering of the package and a trial demodulation with the guessed
// timings.  The histograms are the part that scales with the package (sequential: every width is matched
// against the running bin means) -- k_analyze builds them for all packages of a batch at once, one thread per
// (package, histogram).  The O(1) rest (the guess, the text) is finished on the host from these bins
// (r433b_analyze_host.hpp), and the trial demodulation runs the guessed slicer on the GPU again (k_slice_own).
#pragma once
#include <stdint.h>

#include "../../include/r433b.h"
#include "r433b_core.cuh"
#include "r433b_slice.cuh"

namespace r433b {

constexpr int kHistBins = 16; // MAX_HIST_BINS, src/pulse_analyzer.c:20

using HistBin = r433b_hist_bin;
using Histogram = r433b_histogram;

// |a - b| < tolerance * max(a, b) with the reference's types: the right side is a float product of the float
// tolerance and the int converted to float, the left side the int converted to float (src/pulse_analyzer.c:46)
R4_HD bool hist_close(int a, int b, float tolerance)
{
    int const m = a > b ? a : b;
    int d = a - b;
    if (d < 0) d = -d;
    return (float)d < fmul(tolerance, (float)m);
}

// histogram_sum() of one value, src/pulse_analyzer.c:38-66 (unsorted bins; `sum / count` is an UNSIGNED division)
R4_HD void hist_add(Histogram &h, int v, float tolerance)
{
    unsigned bin = 0;
    for (; bin < h.bins_count; ++bin) {
        HistBin &b = h.bins[bin];
        if (hist_close(v, b.mean, tolerance)) {
            b.count++;
            b.sum += v;
            b.mean = (int)((unsigned)b.sum / b.count);
            if (v < b.min) b.min = v;
            if (v > b.max) b.max = v;
            return;
        }
    }
    if (bin < (unsigned)kHistBins) {
        HistBin &b = h.bins[bin];
        b.count = 1;
        b.sum = b.mean = b.min = b.max = v;
        h.bins_count++;
    }
}

// histogram_delete_bin(), src/pulse_analyzer.c:69-82
R4_HD void hist_delete(Histogram &h, unsigned index)
{
    if (h.bins_count < 1) return;
    for (unsigned n = index; n + 1 < h.bins_count; ++n) h.bins[n] = h.bins[n + 1];
    h.bins_count--;
    HistBin z = {0, 0, 0, 0, 0};
    h.bins[h.bins_count] = z;
}

// histogram_fuse_bins(), src/pulse_analyzer.c:128-153
R4_HD void hist_fuse(Histogram &h, float tolerance)
{
    if (h.bins_count < 2) return;
    for (unsigned n = 0; n + 1 < h.bins_count; ++n) {
        for (unsigned m = n + 1; m < h.bins_count; ++m) {
            HistBin &a = h.bins[n];
            HistBin const &b = h.bins[m];
            if (hist_close(a.mean, b.mean, tolerance)) {
                a.count += b.count;
                a.sum += b.sum;
                a.mean = (int)((unsigned)a.sum / a.count);
                if (b.min < a.min) a.min = b.min;
                if (b.max > a.max) a.max = b.max;
                hist_delete(h, m);
                m--;
            }
        }
    }
}

constexpr float kAnalyzerTolerance = 0.2f; // TOLERANCE, src/pulse_analyzer.c:211

// One of the five histograms of a package (src/pulse_analyzer.c:311-323), `which` in the order they are printed:
// 0 pulses, 1 gaps (last one left out), 2 pulse+gap periods (last left out), 3 gap+pulse periods (not fused),
// 4 all timings (pulses, then gaps).  n >= 1.
R4_HD void analyze_histogram(int which, int const *pulse, int const *gap, unsigned n, Histogram &h)
{
    h.bins_count = 0;
    HistBin z = {0, 0, 0, 0, 0};
    for (int i = 0; i < kHistBins; ++i) h.bins[i] = z;
    float const tol = kAnalyzerTolerance;
    switch (which) {
    case 0:
        for (unsigned i = 0; i < n; ++i) hist_add(h, pulse[i], tol);
        break;
    case 1:
        for (unsigned i = 0; i + 1 < n; ++i) hist_add(h, gap[i], tol);
        break;
    case 2:
        for (unsigned i = 0; i + 1 < n; ++i) hist_add(h, pulse[i] + gap[i], tol);
        break;
    case 3:
        hist_add(h, pulse[0], tol);
        for (unsigned i = 1; i < n; ++i) hist_add(h, pulse[i] + gap[i - 1], tol);
        break;
    default:
        for (unsigned i = 0; i < n; ++i) hist_add(h, pulse[i], tol);
        for (unsigned i = 0; i < n; ++i) hist_add(h, gap[i], tol);
        break;
    }
    if (which != 3) hist_fuse(h, tol);
}

#if defined(__CUDACC__) || defined(R433B_SIMT_EMU)

struct AnalyzeParams {
    r433b_package const *pkgs;
    unsigned n_pkgs;
    int const *pulse_pool, *gap_pool;
    r433b_analysis *out; // n_pkgs entries, device order
};

constexpr int kAnalyzeThreads = 160; // 32 packages x 5 histograms

__global__ void __launch_bounds__(kAnalyzeThreads) k_analyze(AnalyzeParams p)
{
    unsigned const t = blockIdx.x * (unsigned)kAnalyzeThreads + threadIdx.x;
    unsigned const pk = t / 5, which = t % 5;
    if (pk >= p.n_pkgs) return;
    r433b_package const k = p.pkgs[pk];
    int const *pulse = p.pulse_pool + k.pulse_off, *gap = p.gap_pool + k.pulse_off;
    r433b_analysis &a = p.out[pk];
    if (k.num_pulses == 0) { // "No pulses detected."
        a.hist[which].bins_count = 0;
        if (which == 0) a.total_period = 0;
        return;
    }
    Histogram h;
    analyze_histogram((int)which, pulse, gap, k.num_pulses, h);
    a.hist[which] = h;
    if (which == 0) { // pulse_total_period, src/pulse_analyzer.c:289-296: all periods but the last gap
        int total = 0;
        for (unsigned i = 0; i < k.num_pulses; ++i) total += pulse[i] + gap[i];
        a.total_period = total - gap[k.num_pulses - 1];
    }
}

// The trial demodulation of the analyzer (src/pulse_analyzer.c:509-557): every package is sliced by ITS OWN guessed
// device.  One thread per package; pass 0 counts the words, pass 1 (after a host-side prefix sum) stores them.
// PPM / PWM / Manchester guesses first overwrite the package's last gap ("Be sure to terminate package"): the
// thread patches the pool entry for the duration of its run and restores it.
struct OwnSliceParams {
    r433b_package const *pkgs;
    unsigned n_pkgs;
    int const *pulse_pool;
    int *gap_pool;
    SlicerParams const *dev;  // per package; modulation 0 = no guess, nothing to do
    int const *last_gap;      // per package; < 0 = leave the last gap alone
    r433b_pair *pairs;        // per package: pass 0 writes bytes / events, pass 1 reads offset
    uint8_t *arena;
    int pass;
};

__global__ void k_slice_own(OwnSliceParams p)
{
    unsigned const pk = blockIdx.x * blockDim.x + threadIdx.x;
    if (pk >= p.n_pkgs) return;
    SlicerParams const sp = p.dev[pk];
    r433b_pair pr = p.pairs[pk];
    if (p.pass == 0) {
        pr.offset = 0;
        pr.bytes = pr.events = pr.gated_single = pr.gated_multi = 0;
    }
    r433b_package const k = p.pkgs[pk];
    if (sp.modulation != 0 && k.num_pulses > 0 && (p.pass == 0 || pr.bytes)) {
        PulseView pv;
        pv.pulse = p.pulse_pool + k.pulse_off;
        pv.gap = p.gap_pool + k.pulse_off;
        pv.n = k.num_pulses;
        int *last = p.gap_pool + k.pulse_off + k.num_pulses - 1;
        int const saved = *last;
        if (p.last_gap[pk] >= 0) *last = p.last_gap[pk];
        EventWriter w;
        if (p.pass == 0)
            w.init(nullptr, 0);
        else
            w.init(reinterpret_cast<uint32_t *>(p.arena + pr.offset), pr.bytes / 4);
        slice_dispatch(pv, sp, w);
        *last = saved;
        if (p.pass == 0) {
            pr.bytes = w.committed * 4;
            pr.events = w.events;
        }
    }
    if (p.pass == 0) p.pairs[pk] = pr;
}

#endif

} // namespace r433b
