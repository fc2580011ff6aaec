// topk.cuh -- K5: the fused per-CTA top-k machinery shared by the scan (K1/K2) and the inverted-list select
// (K3) kernels.  Candidates are 64-bit rank keys (common.cuh); the CTA keeps them in one shared-memory buffer
// of kCapMax keys:
//
//   phase A  sampling   the first keys of a pass are written straight into per-warp slices of the buffer (no
//                       threshold, no lock); ONE CTA-wide radix select keeps the k best in cbuf[0, k) and
//                       publishes the threshold tau (and its score part, for a cheap float pre-filter);
//   phase B  steady     only ~k/sample of the remaining rows beat tau.  Each warp appends them to its PRIVATE
//                       region cbuf[kSharedKeys + warp*R, ...) with plain stores.  When a private region
//                       fills (rare) that warp alone, under a shared-memory lock, merges it into the shared
//                       top-k set cbuf[0, k) and raises tau; the other warps keep streaming;
//   end of pass         CTA-wide radix select over the shared set + all private regions -> k keys to HBM.
#pragma once
#include "common.cuh"

namespace vs {

constexpr int kCapMax = 8192;       // keys in the CTA candidate buffer (64 KB)
constexpr int kSharedKeys = 2048;   // = VS_MAX_K: the shared top-k set lives in cbuf[0, kSharedKeys)
static_assert(kSharedKeys >= VS_MAX_K, "shared set must hold k keys");

template <int NW>
struct TopkGeom {
    static constexpr int kPrivate = (kCapMax - kSharedKeys) / NW;  // steady-state private keys per warp
    static constexpr int kSample = kCapMax / NW;                   // sampling-phase keys per warp
    static_assert(kPrivate >= 64, "private region too small");
};

struct CtaState {
    uint64_t mbar;
    uint64_t tau;       // current threshold key (0 = accept everything)
    float tau_score;    // score part of tau (-inf while tau == 0): cheap pre-filter
    uint32_t cnt;       // keys in the shared set cbuf[0, cnt)
    uint32_t lock;
    uint32_t scratch;   // CTA-wide counters of the select routines
};

__device__ __forceinline__ void cta_state_reset(CtaState *st) {
    st->cnt = 0;
    st->tau = 0;
    st->tau_score = -INFINITY;
}

// k-th largest over two shared-memory segments (unique keys; zeros allowed as "absent", they rank last).
// PRE: at least k keys in total.  Same radix select as radix_kth_largest, group = warp or CTA.
template <bool BLOCK>
__device__ __forceinline__ uint64_t radix_kth_largest2(const uint64_t *a, int na, int ta, int nta, const uint64_t *b,
                                                       int nb, int tb, int ntb, int k, uint32_t *hist, int t, int nt) {
    const int lane = threadIdx.x & 31;
    uint64_t prefix = 0, mask = 0;
    int rem = k;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = t; i < 256; i += nt) hist[i] = 0;
        group_sync<BLOCK>();
        for (int base = 0; base < na; base += nta) {  // warp-uniform trip counts (na, nta uniform per warp)
            const int i = base + ta;
            const uint64_t x = (i < na) ? a[i] : 0ull;
            hist_add_aggregated(hist, (uint32_t)(x >> shift) & 255u, (i < na) && ((x & mask) == prefix));
        }
        for (int base = 0; base < nb; base += ntb) {
            const int i = base + tb;
            const uint64_t x = (i < nb) ? b[i] : 0ull;
            hist_add_aggregated(hist, (uint32_t)(x >> shift) & 255u, (i < nb) && ((x & mask) == prefix));
        }
        group_sync<BLOCK>();
        uint32_t h[8], s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { h[j] = hist[lane * 8 + j]; s += h[j]; }
        uint32_t incl = s;  // sum over lanes >= lane
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_down_sync(0xffffffffu, incl, d);
            if (lane + d < 32) incl += o;
        }
        const uint32_t above = incl - s;
        const bool mine = (above < (uint32_t)rem) && ((uint32_t)rem <= incl);
        uint32_t digit = 0, newrem = 0;
        if (mine) {
            uint32_t acc = above;
#pragma unroll
            for (int j = 7; j >= 0; --j) {
                if (acc < (uint32_t)rem && acc + h[j] >= (uint32_t)rem) { digit = lane * 8 + j; newrem = rem - acc; }
                acc += h[j];
            }
        }
        const uint32_t owner = __ballot_sync(0xffffffffu, mine);
        const int src = __ffs(owner) - 1;
        digit = __shfl_sync(0xffffffffu, digit, src);
        newrem = __shfl_sync(0xffffffffu, newrem, src);
        prefix |= (uint64_t)digit << shift;
        mask |= (uint64_t)0xff << shift;
        rem = (int)newrem;
        group_sync<BLOCK>();  // hist is rewritten next pass
    }
    return prefix;
}

// ---- phase A -> B: called by the whole CTA after a __syncthreads(); cbuf[0, n) holds the sample (zeros = unused
// slots).  Keeps the k largest in cbuf[0, cnt), publishes tau.
template <int NT>
__device__ __forceinline__ void cta_sample_select(uint64_t *cbuf, int n, int k, uint32_t *hist, CtaState *st) {
    const int tid = threadIdx.x;
    constexpr int PER = (kCapMax + NT - 1) / NT;
    uint64_t mine[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) { const int i = tid + j * NT; mine[j] = (i < n) ? cbuf[i] : 0ull; }
    // zeros rank last, so the k-th largest is exact whenever the sample holds >= k real keys (n >= k by layout)
    const uint64_t kth = radix_kth_largest2<true>(cbuf, n, tid, NT, cbuf, 0, 0, 1, k, hist, tid, NT);
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

// ---- phase B slow path: this warp's private region is full.  Under the lock, fold it into the shared set.
template <int NW>
__device__ __forceinline__ void warp_merge_private(uint64_t *cbuf, int n_priv, int k, uint32_t *hist, CtaState *st) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t *priv = cbuf + kSharedKeys + warp * TopkGeom<NW>::kPrivate;
    __syncwarp();
    if (lane == 0) {
        while (atomicCAS(&st->lock, 0u, 1u) != 0u) __nanosleep(64);
    }
    __syncwarp();
    __threadfence_block();
    const int n_sh = (int)*(volatile uint32_t *)&st->cnt;
    int kept;
    if (n_sh + n_priv <= k) {  // the shared set is not full yet: append
        for (int i = lane; i < n_priv; i += 32) cbuf[n_sh + i] = priv[i];
        kept = n_sh + n_priv;
    } else {
        const uint64_t kth = radix_kth_largest2<false>(cbuf, n_sh, lane, 32, priv, n_priv, lane, 32, k, hist, lane, 32);
        kept = warp_compact_ge(cbuf, n_sh, kth);
        for (int base = 0; base < n_priv; base += 32) {
            const int i = base + lane;
            const uint64_t x = (i < n_priv) ? priv[i] : 0ull;
            const bool keep = (i < n_priv) && (x >= kth);
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) cbuf[kept + __popc(m & lanemask_lt())] = x;
            kept += __popc(m);
        }
        if (lane == 0) {
            *(volatile float *)&st->tau_score = key_score(kth);
            *(volatile uint64_t *)&st->tau = kth;
        }
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) {
        *(volatile uint32_t *)&st->cnt = (uint32_t)kept;
        __threadfence_block();
        atomicExch(&st->lock, 0u);
    }
    __syncwarp();
}

// phase B fast path: append this window's qualifying keys (`ins` lanes) to the warp's private region.
template <int NW>
__device__ __forceinline__ void private_insert(bool ins, uint64_t key, uint64_t *cbuf, int &n_priv, int k,
                                               uint32_t *hist, CtaState *st, uint32_t lt) {
    const uint32_t m = __ballot_sync(0xffffffffu, ins);
    if (m) {
        uint64_t *priv = cbuf + kSharedKeys + (threadIdx.x >> 5) * TopkGeom<NW>::kPrivate;
        if (ins) priv[n_priv + __popc(m & lt)] = key;
        n_priv += __popc(m);
        if (n_priv + 32 > TopkGeom<NW>::kPrivate) {
            warp_merge_private<NW>(cbuf, n_priv, k, hist, st);
            n_priv = 0;
        }
    }
}

// ---- end of pass: called by the whole CTA (it synchronises first).  Exact top-k of the shared set plus every
// warp's private region -> out[0..k) (unsorted; zero padded when fewer than k candidates exist).
template <int NT, int NW>
__device__ __forceinline__ void cta_write_topk(uint64_t *cbuf, int n_priv, int k, uint32_t *hist, CtaState *st,
                                               uint64_t *out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t *priv = cbuf + kSharedKeys + warp * TopkGeom<NW>::kPrivate;
    if (tid == 0) st->scratch = 0;
    __syncthreads();
    const int n_sh = (int)st->cnt;
    if (lane == 0) atomicAdd(&st->scratch, (uint32_t)n_priv);
    __syncthreads();
    const int n_total = n_sh + (int)st->scratch;
    uint64_t kth = 0;
    if (n_total > k)
        kth = radix_kth_largest2<true>(cbuf, n_sh, tid, NT, priv, n_priv, lane, 32, k, hist, tid, NT);
    __syncthreads();
    if (tid == 0) st->scratch = 0;
    __syncthreads();
    for (int i = tid; i < n_sh; i += NT) {
        const uint64_t x = cbuf[i];
        if (x >= kth) out[atomicAdd(&st->scratch, 1u)] = x;
    }
    for (int i = lane; i < n_priv; i += 32) {
        const uint64_t x = priv[i];
        if (x >= kth) out[atomicAdd(&st->scratch, 1u)] = x;
    }
    __syncthreads();
    for (int i = (int)st->scratch + tid; i < k; i += NT) out[i] = 0ull;  // fewer than k candidates in this CTA
}

}  // namespace vs
