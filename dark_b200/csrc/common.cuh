// common.cuh — types and sm_100a PTX helpers shared by the forward-BWT kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace dark {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 8;  // 64-bit keys, 8-bit digits

__host__ __device__ inline u64 ceil_div(u64 a, u64 b) { return (a + b - 1) / b; }

// Checked build (make -C dark_b200/csrc checked; tests run it through DARK_BWT_LIB): bounds assertions on the scattered
// stores and table indices of the kernels.  compute-sanitizer is closed on the GPU pool this was developed on, so this is
// the memory-safety net besides the parity tests; a violation prints its place and traps (the CUDA context dies loudly).
#ifdef DARK_BWT_CHECKED
#define DARK_ASSERT(cond)                                                                                     \
    do {                                                                                                      \
        if (!(cond)) {                                                                                        \
            printf("DARK_ASSERT failed: %s at %s:%d (block %u thread %u)\n", #cond, __FILE__, __LINE__, blockIdx.x, threadIdx.x); \
            __trap();                                                                                         \
        }                                                                                                     \
    } while (0)
#else
#define DARK_ASSERT(cond) do { } while (0)
#endif

// ---- memory-model helpers -------------------------------------------------------------------
// Tile-status words of the decoupled look-back scans are written by one CTA and polled by
// others; they carry value and flag in ONE word, so relaxed gpu-scope accesses suffice.
__device__ __forceinline__ u32 ld_relaxed(const u32* p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ld_relaxed(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u32* p, u32 v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(u64* p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Flag + separate payload (re-rank scan): release/acquire at gpu scope.
__device__ __forceinline__ u32 ld_acquire(const u32* p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(u32* p, u32 v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Streaming (read-once / write-once) accesses: keep them out of L1, evict-first in L2.
__device__ __forceinline__ u64 ld_stream(const u64* p) {
    u64 v;
    asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 ld_stream(const u32* p) {
    u32 v;
    asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(u64* p, u64 v) { asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v)); }
__device__ __forceinline__ void st_stream(u32* p, u32 v) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v)); }

__device__ __forceinline__ u32 lane_id() {
    u32 l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- TMA bulk copy (cp.async.bulk, 1-D) + mbarrier ------------------------------------------
// Used to stage a tile of keys/values from HBM into shared memory with ONE instruction per
// array (SASS: UBLKCP), completion signalled on an mbarrier.
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// dst (shared), src (global) 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, u32 bytes, u64* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}

}  // namespace dark
