// topk.cuh -- K5: the fused per-CTA top-k machinery shared by the scan (K1/K2) and the
// inverted-list select (K3) kernels.  Candidates are 64-bit rank keys (common.cuh).
//   per-warp staging buffer (kStage keys)  --flush under a smem lock-->  CTA buffer (cap keys)
//   CTA buffer full -> the flushing warp radix-selects it down to k and publishes the new threshold
//   end of pass     -> the whole CTA radix-selects the exact top-k and writes k keys
#pragma once
#include "common.cuh"

namespace vs {

constexpr int kStage = 64;  // per-warp staging entries

struct CtaState {
    uint64_t mbar;
    uint64_t tau;       // current threshold key (0 = accept everything)
    float tau_score;    // score part of tau (-inf while tau == 0): cheap pre-filter
    uint32_t cnt;       // entries in cbuf
    uint32_t lock;
};

// Called by one whole warp holding the lock: shrink cbuf[0..n) to its k largest, publish tau.
__device__ __forceinline__ int warp_prune(uint64_t *cbuf, int n, int k, uint32_t *hist, CtaState *st) {
    const int lane = threadIdx.x & 31;
    uint64_t kth = radix_kth_largest<false>(cbuf, n, k, hist, lane, 32);
    int kept = warp_compact_ge(cbuf, n, kth);
    if (lane == 0) {
        *(volatile float *)&st->tau_score = key_score(kth);
        *(volatile uint64_t *)&st->tau = kth;
    }
    return kept;
}

__device__ __forceinline__ void warp_flush(uint64_t *cbuf, uint64_t *stage, int n_stage, int k, int cap,
                                           uint32_t *hist, CtaState *st) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane == 0) {
        while (atomicCAS(&st->lock, 0u, 1u) != 0u) __nanosleep(32);
    }
    __syncwarp();
    __threadfence_block();
    int c = (int)*(volatile uint32_t *)&st->cnt;
    if (c + n_stage > cap) c = warp_prune(cbuf, c, k, hist, st);
    for (int i = lane; i < n_stage; i += 32) cbuf[c + i] = stage[i];
    __syncwarp();
    __threadfence_block();
    if (lane == 0) {
        *(volatile uint32_t *)&st->cnt = (uint32_t)(c + n_stage);
        __threadfence_block();
        atomicExch(&st->lock, 0u);
    }
    __syncwarp();
}


// Warp-level insertion of this window's qualifying keys (`ins` lanes) into the staging buffer.
__device__ __forceinline__ void stage_insert(bool ins, uint64_t key, uint64_t *stage, int &n_stage, uint64_t *cbuf,
                                             int k, int cap, uint32_t *hist, CtaState *st, uint32_t lt) {
    const uint32_t m = __ballot_sync(0xffffffffu, ins);
    if (m) {
        if (ins) stage[n_stage + __popc(m & lt)] = key;
        n_stage += __popc(m);
        if (n_stage > kStage - 32) {
            warp_flush(cbuf, stage, n_stage, k, cap, hist, st);
            n_stage = 0;
        }
    }
}

// Sampling phase.  The first `cap` keys of a pass are written straight into cbuf (no threshold, no lock;
// unused slots = 0).  Then the whole CTA calls this once (after a __syncthreads()): keep the k largest,
// publish the threshold.  From here on only ~k/cap of the remaining rows pass the threshold, so the locked
// flush / single-warp prune path below becomes the rare case instead of the start-up cost of every pass.
template <int NT, int CAP_MAX>
__device__ __forceinline__ void cta_sample_select(uint64_t *cbuf, int cap, int k, uint32_t *hist, CtaState *st) {
    const int tid = threadIdx.x;
    constexpr int PER = (CAP_MAX + NT - 1) / NT;
    uint64_t mine[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) { const int i = tid + j * NT; mine[j] = (i < cap) ? cbuf[i] : 0ull; }
    // zeros are the only duplicates; they rank last, so the k-th largest is exact whenever >= k real keys exist
    const uint64_t kth = radix_kth_largest<true>(cbuf, cap, k, hist, tid, NT);
    if (tid == 0) st->cnt = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j)
        if (mine[j] != 0ull && mine[j] >= kth) cbuf[atomicAdd(&st->cnt, 1u)] = mine[j];
    if (tid == 0) {
        st->tau = kth;  // 0 when the sample held fewer than k real keys: keep accepting everything
        st->tau_score = kth ? key_score(kth) : -INFINITY;
    }
    __syncthreads();
}

// End of pass, called by the whole CTA after a __syncthreads(): exact top-k of cbuf -> out[0..k) (unsorted,
// zero padded when fewer than k candidates exist).
template <int NT>
__device__ __forceinline__ void cta_write_topk(uint64_t *cbuf, int k, uint32_t *hist, CtaState *st, uint64_t *out) {
    const int tid = threadIdx.x;
    const int n = (int)st->cnt;
    if (n > k) {
        const uint64_t kth = radix_kth_largest<true>(cbuf, n, k, hist, tid, NT);
        __shared__ uint32_t out_cnt;
        if (tid == 0) out_cnt = 0;
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            uint64_t x = cbuf[i];
            if (x >= kth) out[atomicAdd(&out_cnt, 1u)] = x;
        }
    } else {
        for (int i = tid; i < k; i += NT) out[i] = (i < n) ? cbuf[i] : 0ull;
    }
}

}  // namespace vs
