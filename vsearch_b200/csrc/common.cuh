// common.cuh -- shared device/host helpers for the vsearch_b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/vsearch_b200.h"

namespace vs {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char *fmt, ...);
#define VS_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            vs::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return VS_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)
#define VS_REQUIRE(cond, code, ...)                                                        \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            vs::set_error(__VA_ARGS__);                                                    \
            return (code);                                                                 \
        }                                                                                  \
    } while (0)

// ---------------------------------------------------------------- rank keys
// key = ordered(score) << 32 | ~id : larger key ranks first; ties on score -> lower id wins.
// -0.0 is folded into +0.0 (torch compares them equal).  key 0 is never a real candidate.
__device__ __forceinline__ uint64_t make_key(float s, uint32_t id) {
    s = s + 0.0f;
    uint32_t b = __float_as_uint(s);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((uint64_t)b << 32) | (uint64_t)(~id);
}
__device__ __forceinline__ float key_score(uint64_t key) {
    uint32_t b = (uint32_t)(key >> 32);
    b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
    return __uint_as_float(b);
}
__device__ __forceinline__ uint32_t key_id(uint64_t key) { return ~(uint32_t)key; }

__device__ __forceinline__ float round_score(float s, int mode) {
    if (mode == VS_F16) return __half2float(__float2half_rn(s));
    if (mode == VS_BF16) return __bfloat162float(__float2bfloat16_rn(s));
    return s;
}

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {  // read-once index stream: bypass L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk TMA copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---------------------------------------------------------------- group sync (warp or block)
template <bool BLOCK>
__device__ __forceinline__ void group_sync() {
    if constexpr (BLOCK) __syncthreads(); else __syncwarp();
}

// Histogram increment.  Candidate sets are full of ties (binary index, untouched rows of the inverted path): when
// every active lane of the warp holds the same digit -- the case that would serialise 32-way on one shared-memory
// word -- one leader adds the whole count; otherwise plain per-lane atomics.  (__match_any_sync would aggregate the
// mixed case too, but MATCH.ANY costs more than the conflicts it removes: measured, profiles/README.md.)
// Must be called by all 32 lanes.
__device__ __forceinline__ void hist_add_aggregated(uint32_t *hist, uint32_t digit, bool active) {
    const uint32_t amask = __ballot_sync(0xffffffffu, active);
    if (amask == 0) return;
    const int leader = __ffs(amask) - 1;
    const uint32_t d0 = __shfl_sync(0xffffffffu, digit, leader);
    if (__all_sync(0xffffffffu, !active || digit == d0)) {
        if ((int)(threadIdx.x & 31) == leader) atomicAdd(&hist[d0], (uint32_t)__popc(amask));
    } else if (active) {
        atomicAdd(&hist[digit], 1u);
    }
}

// k-th largest of n UNIQUE 64-bit keys held in shared memory (1 <= k <= n).
// 8 passes of 8-bit MSD radix select; `hist` is 256 words of shared memory owned by the group.
// Called by all `nt` threads of the group (a warp when !BLOCK, the whole CTA when BLOCK).
template <bool BLOCK>
__device__ uint64_t radix_kth_largest(const uint64_t *buf, int n, int k, uint32_t *hist, int t, int nt) {
    const int lane = threadIdx.x & 31;
    uint64_t prefix = 0, mask = 0;
    int rem = k;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = t; i < 256; i += nt) hist[i] = 0;
        group_sync<BLOCK>();
        for (int base = 0; base < n; base += nt) {  // warp-uniform trip count: hist_add_aggregated is convergent
            const int i = base + t;
            const uint64_t x = (i < n) ? buf[i] : 0ull;
            hist_add_aggregated(hist, (uint32_t)(x >> shift) & 255u, (i < n) && ((x & mask) == prefix));
        }
        group_sync<BLOCK>();
        // every warp redundantly locates the digit: lane owns bins [8*lane, 8*lane+8)
        uint32_t h[8];
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { h[j] = hist[lane * 8 + j]; s += h[j]; }
        uint32_t incl = s;  // sum over lanes >= lane
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_down_sync(0xffffffffu, incl, d);
            if (lane + d < 32) incl += o;
        }
        uint32_t above = incl - s;
        bool mine = (above < (uint32_t)rem) && ((uint32_t)rem <= incl);
        uint32_t digit = 0, newrem = 0;
        if (mine) {
            uint32_t acc = above;
#pragma unroll
            for (int j = 7; j >= 0; --j) {
                if (acc < (uint32_t)rem && acc + h[j] >= (uint32_t)rem) { digit = lane * 8 + j; newrem = rem - acc; }
                acc += h[j];
            }
        }
        uint32_t owner = __ballot_sync(0xffffffffu, mine);
        int src = __ffs(owner) - 1;  // exactly one lane when k <= n
        digit = __shfl_sync(0xffffffffu, digit, src);
        newrem = __shfl_sync(0xffffffffu, newrem, src);
        prefix |= (uint64_t)digit << shift;
        mask |= (uint64_t)0xff << shift;
        rem = (int)newrem;
        group_sync<BLOCK>();  // hist is rewritten next pass
    }
    return prefix;
}

// In-place compaction by one warp: keep keys >= kth.  Returns the kept count (all lanes).
__device__ __forceinline__ int warp_compact_ge(uint64_t *buf, int n, uint64_t kth) {
    const int lane = threadIdx.x & 31;
    int out = 0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        uint64_t x = (i < n) ? buf[i] : 0;
        bool keep = (i < n) && (x >= kth);
        uint32_t m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) buf[out + __popc(m & lanemask_lt())] = x;  // out + rank <= i : never clobbers unread data
        out += __popc(m);
        __syncwarp();
    }
    return out;
}

}  // namespace vs
