// simt.hpp -- TEST INFRASTRUCTURE: a tiny SIMT emulator so that the product's CUDA kernels
// (rtl_433_b200/csrc/*.cuh) can be executed on a CPU by the `-m "not gpu"` tests.
//
// This is not a CPU implementation of anything: it runs the SAME kernel source, one fibre per CUDA
// thread, 32 fibres per warp, with the warp collectives (__shfl*_sync, __ballot_sync, __all_sync,
// __syncwarp) implemented as rendezvous points.  Lanes run one after the other up to their next
// collective, in ascending or (SIMT_REVERSE=1) descending lane order -- the two adversarial schedules
// for a missing __syncwarp() between a shared-memory write and a cross-lane read.  Warps of a block
// run one after the other (the kernels here never synchronise across warps); `__shared__` becomes
// `static`.  The product never includes this file: it is only on the include path of tests/_build.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define __restrict__

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

namespace simt {

constexpr int kStack = 256 * 1024;

struct Lane {
    void *sp = nullptr;
    char *stack = nullptr;
    bool live = false, waiting = false;
    int kind = 0;            // collective the lane is parked at
    unsigned long long val = 0; // operand / result
    int arg = 0;             // source lane / delta / predicate
    dim3 tid;
};

enum { kShflIdx = 1, kShflUp, kShflDown, kShflXor, kBallot, kSync };

struct State {
    Lane lanes[32];
    int cur = -1;
    void *sched_sp = nullptr;
    dim3 block_idx, block_dim, grid_dim;
    unsigned char *dyn_smem = nullptr;
    void (*entry)(void *) = nullptr;
    void *entry_arg = nullptr;
};

inline State &st()
{
    static State s;
    return s;
}

extern "C" void simt_switch(void **save_sp, void *new_sp);
#ifndef SIMT_NO_ASM
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");
#endif

inline void to_scheduler()
{
    State &s = st();
    Lane &l = s.lanes[s.cur];
    simt_switch(&l.sp, s.sched_sp);
}

extern "C" inline void simt_lane_main()
{
    State &s = st();
    s.entry(s.entry_arg);
    s.lanes[s.cur].live = false;
    to_scheduler();
    abort(); // a finished lane is never resumed
}

inline void prepare(Lane &l)
{
    if (!l.stack) l.stack = (char *)aligned_alloc(64, kStack);
    // stack image popped by simt_switch: r15 r14 r13 r12 rbx rbp, then the return address
    uintptr_t top = ((uintptr_t)l.stack + kStack) & ~(uintptr_t)63;
    void **sp = (void **)(top - 8); // slot of simt_lane_main's (fake) return address: rsp % 16 == 8 at its entry
    *sp = nullptr;
    sp -= 1;
    *sp = (void *)&simt_lane_main;
    for (int i = 0; i < 6; ++i) {
        sp -= 1;
        *sp = nullptr;
    }
    l.sp = sp;
    l.live = true;
    l.waiting = false;
}

inline unsigned long long collective(int kind, unsigned long long v, int arg)
{
    State &s = st();
    Lane &l = s.lanes[s.cur];
    l.kind = kind;
    l.val = v;
    l.arg = arg;
    l.waiting = true;
    to_scheduler();
    return l.val;
}

inline void resolve(State &s)
{
    int kind = 0;
    for (auto &l : s.lanes)
        if (l.live) {
            if (!kind) kind = l.kind;
            if (l.kind != kind) {
                fprintf(stderr, "simt: lanes parked at different collectives (%d vs %d): divergent warp-level call\n", kind, l.kind);
                abort();
            }
        }
    unsigned long long in[32];
    int arg[32];
    bool live[32];
    for (int i = 0; i < 32; ++i) {
        in[i] = s.lanes[i].val;
        arg[i] = s.lanes[i].arg;
        live[i] = s.lanes[i].live;
    }
    unsigned ballot = 0;
    if (kind == kBallot)
        for (int i = 0; i < 32; ++i)
            if (live[i] && arg[i]) ballot |= 1u << i;
    for (int i = 0; i < 32; ++i) {
        if (!live[i]) continue;
        int src = i;
        switch (kind) {
        case kShflIdx: src = arg[i] & 31; break;
        case kShflUp: src = i - arg[i] >= 0 ? i - arg[i] : i; break;
        case kShflDown: src = i + arg[i] < 32 ? i + arg[i] : i; break;
        case kShflXor: src = i ^ arg[i]; break;
        default: break;
        }
        if (kind == kBallot)
            s.lanes[i].val = ballot;
        else if (kind != kSync)
            s.lanes[i].val = live[src] ? in[src] : in[i];
        s.lanes[i].waiting = false;
    }
}

// one warp: 32 fibres executing entry(arg)
inline void run_warp(void (*entry)(void *), void *arg, dim3 const tids[32], int n_lanes)
{
    State &s = st();
    static int reverse = -1;
    if (reverse < 0) reverse = getenv("SIMT_REVERSE") && atoi(getenv("SIMT_REVERSE")) ? 1 : 0;
    s.entry = entry;
    s.entry_arg = arg;
    for (int i = 0; i < 32; ++i) {
        s.lanes[i].live = false;
        if (i < n_lanes) {
            prepare(s.lanes[i]);
            s.lanes[i].tid = tids[i];
        }
    }
    for (;;) {
        bool any = false, progressed = false;
        for (int k = 0; k < 32; ++k) {
            int i = reverse ? 31 - k : k;
            Lane &l = s.lanes[i];
            if (!l.live) continue;
            any = true;
            if (l.waiting) continue;
            s.cur = i;
            simt_switch(&s.sched_sp, l.sp);
            progressed = true;
        }
        if (!any) break;
        bool all_waiting = true;
        for (auto &l : s.lanes)
            if (l.live && !l.waiting) all_waiting = false;
        if (all_waiting) {
            bool some = false;
            for (auto &l : s.lanes) some |= l.live;
            if (some) resolve(s);
        } else if (!progressed) {
            fprintf(stderr, "simt: scheduler made no progress\n");
            abort();
        }
    }
    s.cur = -1;
}

struct ThunkBase {
    virtual void run() = 0;
    virtual ~ThunkBase() {}
    static void call(void *self) { ((ThunkBase *)self)->run(); }
};

template <class F>
struct Thunk : ThunkBase {
    F f;
    explicit Thunk(F g) : f(g) {}
    void run() override { f(); }
};

template <class K, class... Args>
inline void launch(dim3 grid, dim3 block, size_t smem, K kernel, Args... args)
{
    State &s = st();
    std::vector<unsigned char> dyn(smem + 64);
    s.dyn_smem = (unsigned char *)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    s.grid_dim = grid;
    s.block_dim = block;
    auto body = [&]() { kernel(args...); };
    Thunk<decltype(body)> th(body);
    unsigned const nthreads = block.x * block.y * block.z;
    for (unsigned b = 0; b < grid.x; ++b) {
        s.block_idx = dim3(b);
        for (unsigned w0 = 0; w0 < nthreads; w0 += 32) {
            dim3 tids[32];
            int n = 0;
            for (; n < 32 && w0 + n < nthreads; ++n) tids[n] = dim3(w0 + n);
            run_warp(&ThunkBase::call, &th, tids, n);
        }
    }
    s.dyn_smem = nullptr;
}

} // namespace simt

#define threadIdx (simt::st().lanes[simt::st().cur].tid)
#define blockIdx (simt::st().block_idx)
#define blockDim (simt::st().block_dim)
#define gridDim (simt::st().grid_dim)

template <class T>
inline T simt_shfl(int kind, T v, int arg)
{
    static_assert(sizeof(T) <= 8, "shuffle operand");
    unsigned long long u = 0;
    memcpy(&u, &v, sizeof(T));
    u = simt::collective(kind, u, arg);
    T r;
    memcpy(&r, &u, sizeof(T));
    return r;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return simt_shfl(simt::kShflIdx, v, src); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d) { return simt_shfl(simt::kShflUp, v, (int)d); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d) { return simt_shfl(simt::kShflDown, v, (int)d); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return simt_shfl(simt::kShflXor, v, m); }
inline unsigned __ballot_sync(unsigned, int pred) { return (unsigned)simt::collective(simt::kBallot, 0, pred ? 1 : 0); }
inline int __all_sync(unsigned m, int pred)
{
    unsigned live = __ballot_sync(m, 1), b = __ballot_sync(m, pred);
    return b == live;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __reduce_max_sync(unsigned m, int v)
{
    for (int o = 16; o > 0; o >>= 1) {
        int t = __shfl_xor_sync(m, v, o);
        v = t > v ? t : v;
    }
    return v;
}
inline int __reduce_min_sync(unsigned m, int v)
{
    for (int o = 16; o > 0; o >>= 1) {
        int t = __shfl_xor_sync(m, v, o);
        v = t < v ? t : v;
    }
    return v;
}
inline void __syncwarp(unsigned = 0xffffffffu) { simt::collective(simt::kSync, 0, 0); }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
template <class T> inline T __ldg(T const *p) { return *p; }
template <class T> inline T __ldcs(T const *p) { return *p; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    unsigned long long both = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned s = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(both >> (8 * (s & 7))) & 0xff;
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0;
        r |= byte << (8 * i);
    }
    return r;
}
inline int __dp4a(int a, int b, int c)
{
    for (int i = 0; i < 4; ++i) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
    return c;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline int __float2int_rz(float f) { return (int)f; }

inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o | v; return o; }
inline unsigned atomicMax(unsigned *p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
