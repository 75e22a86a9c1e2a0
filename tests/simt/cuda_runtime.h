// cuda_runtime.h -- TEST INFRASTRUCTURE: a stand-in for the CUDA runtime API so that the product's C ABI
// translation unit (rtl_433_b200/csrc/r433b_api.cu) can be compiled by g++ against the SIMT emulator
// (simt.hpp) into tests/_build/libr433b_emu.so.  "Device" memory is host memory, everything is
// synchronous, events measure nothing.  Only tests ever put this directory on an include path.
#pragma once
#include "simt.hpp"
#include <map>
#include <sys/mman.h>
#include <unistd.h>

namespace simt {
inline std::map<void *, void *> &alloc_table()
{
    static std::map<void *, void *> t;
    return t;
}
} // namespace simt

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct simt_stream *cudaStream_t;
typedef struct simt_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaDevAttrMultiProcessorCount = 16 };

inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 2; return cudaSuccess; } // a 2-SM "GPU": small fixed grids
inline char const *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// "Device" allocations sit between two inaccessible guard pages (an electric fence): an out-of-bounds access of
// a kernel faults here instead of passing silently.  The buffer starts on a page boundary right behind the
// front guard (catches reads in front of an allocation); with SIMT_GUARD=back it ends (rounded up to 16 bytes)
// right in front of the rear guard instead.
struct simt_alloc_hdr { void *map; size_t map_len; };
inline cudaError_t cudaMalloc(void **p, size_t n)
{
    static long const page = sysconf(_SC_PAGESIZE);
    static bool const back = getenv("SIMT_GUARD") && !strcmp(getenv("SIMT_GUARD"), "back");
    size_t const body = (n + page - 1) / page * page;
    size_t const len = body + 3 * page; // [header page][guard][body][guard]
    char *m = (char *)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == (char *)MAP_FAILED) { *p = nullptr; return cudaErrorMemoryAllocation; }
    mprotect(m + page, page, PROT_NONE);
    mprotect(m + 2 * page + body, page, PROT_NONE);
    char *u = m + 2 * page;
    if (back) u += (body - n) / 16 * 16;
    memset(m + 2 * page, 0xA5, body); // device memory is not zeroed
    simt_alloc_hdr h{m, len};
    memcpy(m, &h, sizeof(h));
    simt::alloc_table()[(void *)u] = m;
    *p = u;
    return cudaSuccess;
}
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFree(void *p)
{
    if (!p) return cudaSuccess;
    auto &t = simt::alloc_table();
    auto it = t.find(p);
    if (it == t.end()) abort();
    simt_alloc_hdr h;
    memcpy(&h, it->second, sizeof(h));
    t.erase(it);
    munmap(h.map, h.map_len);
    return cudaSuccess;
}
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, void const *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, void const *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, void const *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr)
{
    for (size_t r = 0; r < h; ++r) memcpy((char *)d + r * dp, (char const *)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
inline cudaError_t cudaFuncSetAttribute(void const *, int, int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
