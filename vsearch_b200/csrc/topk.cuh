// topk.cuh -- K5: the fused per-CTA top-k machinery shared by the scan (K1/K2) and the inverted-list select
// (K3) kernels.  Candidates are 64-bit rank keys (common.cuh); the CTA keeps them in one shared-memory buffer
// of kCapMax keys:
//
//   phase A  sampling   the first keys of a pass are written straight into per-warp slices of the buffer (no
//                       threshold, no lock); ONE CTA-wide radix select keeps the k best in cbuf[0, k) and
//                       publishes the threshold tau (and its score part, for a cheap float pre-filter);
//   phase B  steady     only ~k/sample of the remaining rows beat tau.  Each warp appends them to its PRIVATE
//                       region cbuf[kSharedKeys + warp*R, ...) with plain stores -- no lock, no atomics.  When
//                       a private region runs low (rare; k ~ 1000 or adversarial order) the warp raises the
//                       CTA's "join" epoch; every warp polls that word once per group of windows (it shares a
//                       64-bit load with the float threshold) and joins ONE CTA-wide re-selection that folds all
//                       private regions into the shared top-k set cbuf[0, k) and raises tau;
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
    // "gate": ONE 64-bit word polled by the streaming warps.  lo = float bits of tau's score (-inf while tau == 0,
    // the cheap pre-filter), hi = join epoch requested.  The halves are written with separate 32-bit stores.
    uint32_t gate_tau_score;
    uint32_t gate_epoch;
    uint32_t cnt;       // keys in the shared set cbuf[0, cnt)
    uint32_t scratch;   // CTA-wide counter of the select routines
    uint32_t done;      // warps that finished streaming this pass
    uint32_t n_app;     // K3: keys in the CTA-wide append region cbuf[kSharedKeys, kCapMax) (may count past the end)
    uint32_t tau_ob;    // K2: best histogram threshold so far, as ordered score bits (atomicMax)
};
static_assert(offsetof(CtaState, gate_tau_score) % 8 == 0, "gate must be 8-byte aligned");

__device__ __forceinline__ void cta_state_reset(CtaState *st) {
    st->cnt = 0;
    st->tau = 0;
    st->gate_tau_score = __float_as_uint(-INFINITY);
    st->gate_epoch = 0;
    st->done = 0;
    st->n_app = 0;
    st->tau_ob = 0;
    st->scratch = 0;
}
__device__ __forceinline__ uint64_t gate_load(const CtaState *st) {
    return *reinterpret_cast<const volatile uint64_t *>(&st->gate_tau_score);
}
__device__ __forceinline__ float gate_tau_score(uint64_t gate) { return __uint_as_float((uint32_t)gate); }
__device__ __forceinline__ uint32_t gate_epoch(uint64_t gate) { return (uint32_t)(gate >> 32); }
__device__ __forceinline__ void publish_tau(CtaState *st, uint64_t kth) {
    *(volatile uint32_t *)&st->gate_tau_score = __float_as_uint(kth ? key_score(kth) : -INFINITY);
    *(volatile uint64_t *)&st->tau = kth;
}

// Threshold of the k largest over two shared-memory segments (unique keys; zeros allowed as "absent", they rank
// last).  PRE: at least k keys in total.  Same MSD radix select as radix_kth_largest, group = warp or CTA, but it
// with EARLY it returns as soon as the answer is decided: the result T is then a THRESHOLD (k-th key with its undecided low bits
// cleared) such that exactly k keys are >= T -- or, with APPROX, at least k and at most max(k, approx_cap) keys
// (phase A only needs a safe filter, not the exact k-th: 2-3 passes instead of 8).
template <bool BLOCK, bool EARLY = false, bool APPROX = false>
__device__ __forceinline__ uint64_t radix_kth_largest2(const uint64_t *a, int na, int ta, int nta, const uint64_t *b,
                                                       int nb, int tb, int ntb, int k, uint32_t *hist, int t, int nt,
                                                       int approx_cap = 0) {
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
        uint32_t digit = 0, newrem = 0, hd = 0;
        if (mine) {
            uint32_t acc = above;
#pragma unroll
            for (int j = 7; j >= 0; --j) {
                if (acc < (uint32_t)rem && acc + h[j] >= (uint32_t)rem) {
                    digit = lane * 8 + j; newrem = rem - acc;
                    if constexpr (EARLY) hd = h[j];
                }
                acc += h[j];
            }
        }
        const uint32_t owner = __ballot_sync(0xffffffffu, mine);
        const int src = __ffs(owner) - 1;
        digit = __shfl_sync(0xffffffffu, digit, src);
        newrem = __shfl_sync(0xffffffffu, newrem, src);
        if constexpr (EARLY) hd = __shfl_sync(0xffffffffu, hd, src);
        prefix |= (uint64_t)digit << shift;
        mask |= (uint64_t)0xff << shift;
        rem = (int)newrem;
        group_sync<BLOCK>();  // hist is rewritten next pass (or by the caller)
        if constexpr (EARLY) {
            const uint32_t ge = (uint32_t)k - newrem + hd;  // keys >= prefix (low bits cleared); same value in every thread
            if (ge == (uint32_t)k) break;                   // the whole bucket is inside the top k: decided
            if (APPROX && ge <= (uint32_t)approx_cap) break;
        }
    }
    return prefix;
}

// ---- phase A -> B: called by the whole CTA after a __syncthreads(); cbuf[0, n) holds the sample (zeros = unused
// slots).  Keeps the k largest in cbuf[0, cnt), publishes tau.
template <int NT, bool APPROX = false>
__device__ __forceinline__ void cta_sample_select(uint64_t *cbuf, int n, int k, uint32_t *hist, CtaState *st,
                                                  int approx_cap = 0) {
    const int tid = threadIdx.x;
    constexpr int PER = (kCapMax + NT - 1) / NT;
    uint64_t mine[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) { const int i = tid + j * NT; mine[j] = (i < n) ? cbuf[i] : 0ull; }
    // zeros rank last, so the k-th largest is exact whenever the sample holds >= k real keys (n >= k by layout)
    const uint64_t kth = radix_kth_largest2<true, APPROX, APPROX>(cbuf, n, tid, NT, cbuf, 0, 0, 1, k, hist, tid, NT, approx_cap);
    if (tid == 0) st->cnt = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j)
        if (mine[j] != 0ull && mine[j] >= kth) cbuf[atomicAdd(&st->cnt, 1u)] = mine[j];
    if (tid == 0) publish_tau(st, kth);  // 0 when the sample held fewer than k real keys: keep accepting everything
    __syncthreads();
}

// ---- score histogram (K2, scan.cu): every row that becomes a candidate is counted in a two-level histogram of its
// ORDERED score bits (the high half of its rank key): 2^13 fine buckets (sign, exponent, 4 mantissa bits) and 128
// coarse ones (each = 64 fine buckets).  The lower bound T of the highest fine bucket whose suffix count reaches k is
// a safe filter at any time -- at least k rows with a score >= T have been seen, so a row scoring below T is not in
// the top k -- and it only rises as rows are counted.  It replaces the sampling phase and its CTA-wide radix select:
// a warp whose private region fills asks the histogram (one warp, ~100 instructions, no barrier), drops its keys
// below T and goes on; at the end of a pass everything >= the final T (between k and ~2k keys) is handed to the merge
// kernel.  The exact 64-bit machinery (cta_join) stays behind it for what a score bucket cannot separate: masses of
// equal scores, adversarially ordered rows.
constexpr int kHistFineBits = 13;
constexpr int kHistFine = 1 << kHistFineBits;   // u32 counters: 32 KB
constexpr int kHistCoarse = 128;

// Count the warp's new candidates (`ins` lanes, ordered score bits `ob`).  Lanes of one fine bucket are aggregated
// first: while the filter is still open every row is counted, and on a sparse query most rows share ONE score (0) --
// per-lane atomics on one shared-memory word serialise.  All 32 lanes.
__device__ __forceinline__ void hist_count(uint32_t *coarse, uint32_t *fine, bool ins, uint32_t ob) {
    const uint32_t fb = ob >> (32 - kHistFineBits);
    const uint32_t grp = __match_any_sync(0xffffffffu, ins ? fb : 0xffffffffu);
    if (ins && (__ffs(grp) - 1) == (int)(threadIdx.x & 31)) {
        const uint32_t n = (uint32_t)__popc(grp);
        atomicAdd(&fine[fb], n);
        atomicAdd(&coarse[fb >> (kHistFineBits - 7)], n);
    }
}

// Called by all 32 lanes of a warp.  Returns the threshold as ordered score bits (bucket lower bound), 0 when fewer
// than k rows have been counted.  Other warps may be counting meanwhile: every increment read here belongs to a
// distinct real row of that bucket, so the answer is safe whatever the interleaving.
static __device__ __noinline__ uint32_t hist_threshold(const uint32_t *coarse, const uint32_t *fine, int k) {
    const int lane = threadIdx.x & 31;
    const volatile uint32_t *vc = coarse + 4 * lane;
    const uint32_t c[4] = {vc[0], vc[1], vc[2], vc[3]};
    const uint32_t s = c[0] + c[1] + c[2] + c[3];
    uint32_t incl = s;  // sum over lanes >= lane
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_down_sync(0xffffffffu, incl, d);
        if (lane + d < 32) incl += o;
    }
    const uint32_t above = incl - s;
    const bool mine = (above < (uint32_t)k) && ((uint32_t)k <= incl);
    uint32_t cb = 0, above_cb = 0;
    if (mine) {
        uint32_t acc = above;
#pragma unroll
        for (int j = 3; j >= 0; --j) {
            if (acc < (uint32_t)k && acc + c[j] >= (uint32_t)k) { cb = 4 * lane + j; above_cb = acc; }
            acc += c[j];
        }
    }
    const uint32_t owner = __ballot_sync(0xffffffffu, mine);
    if (owner == 0) return 0u;
    const int src = __ffs(owner) - 1;
    cb = __shfl_sync(0xffffffffu, cb, src);
    above_cb = __shfl_sync(0xffffffffu, above_cb, src);
    const volatile uint32_t *vf = fine + cb * 64 + 2 * lane;
    const uint32_t f0 = vf[0], f1 = vf[1];
    uint32_t incl2 = f0 + f1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_down_sync(0xffffffffu, incl2, d);
        if (lane + d < 32) incl2 += o;
    }
    const uint32_t above2 = above_cb + incl2 - (f0 + f1);
    const bool mine2 = (above2 < (uint32_t)k) && ((uint32_t)k <= above2 + f0 + f1);
    const uint32_t fb = (above2 + f1 >= (uint32_t)k) ? cb * 64 + 2 * lane + 1 : cb * 64 + 2 * lane;
    const uint32_t owner2 = __ballot_sync(0xffffffffu, mine2);
    if (owner2 == 0) return (cb * 64) << (32 - kHistFineBits);   // fine counters lag the coarse ones: coarse bound
    return __shfl_sync(0xffffffffu, fb, __ffs(owner2) - 1) << (32 - kHistFineBits);
}

// A warp raises the CTA's float pre-filter to the histogram threshold and drops its own keys below it.  All 32 lanes.
template <int NW>
__device__ __forceinline__ void warp_refresh(const uint32_t *coarse, const uint32_t *fine, int k, uint64_t *cbuf, int &n_priv,
                                             CtaState *st) {
    const uint32_t ob = hist_threshold(coarse, fine, k);
    if (ob == 0) return;
    if ((threadIdx.x & 31) == 0) {
        const uint32_t old = atomicMax(&st->tau_ob, ob);
        // a late writer may put back a slightly older (lower) bound: still a valid filter
        if (ob > old) *(volatile uint32_t *)&st->gate_tau_score = __float_as_uint(key_score((uint64_t)ob << 32));
    }
    uint64_t *priv = cbuf + kSharedKeys + (threadIdx.x >> 5) * TopkGeom<NW>::kPrivate;
    n_priv = warp_compact_ge(priv, n_priv, (uint64_t)ob << 32);
}

// ---- phase B: append this window's qualifying keys (`ins` lanes) to the warp's private region (room is
// guaranteed by the caller's join protocol).
template <int NW>
__device__ __forceinline__ void private_insert(bool ins, uint64_t key, uint64_t *cbuf, int &n_priv, uint32_t lt) {
    const uint32_t m = __ballot_sync(0xffffffffu, ins);
    if (m) {
        uint64_t *priv = cbuf + kSharedKeys + (threadIdx.x >> 5) * TopkGeom<NW>::kPrivate;
        if (ins) priv[n_priv + __popc(m & lt)] = key;
        n_priv += __popc(m);
    }
}

// ---- phase B slow path: CTA-wide re-selection, executed by ALL warps of the CTA (each one gets here through
// the epoch poll).  Folds every private region into the shared set, keeps the k best, raises tau, empties the
// private regions.  No lock: the barrier is the synchronisation.
template <int NT, int NW>
__device__ __noinline__ void cta_join(uint64_t *cbuf, const int n_priv, int k, uint32_t *hist, CtaState *st) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t *priv = cbuf + kSharedKeys + warp * TopkGeom<NW>::kPrivate;
    if (tid == 0) st->scratch = 0;
    __syncthreads();
    const int n_sh = (int)st->cnt;
    if (lane == 0) atomicAdd(&st->scratch, (uint32_t)n_priv);
    __syncthreads();
    const int n_total = n_sh + (int)st->scratch;
    uint64_t kth = 0;
    if (n_total > k) kth = radix_kth_largest2<true>(cbuf, n_sh, tid, NT, priv, n_priv, lane, 32, k, hist, tid, NT);
    // survivors -> registers, then rewrite the shared set
    constexpr int SH_PER = (kSharedKeys + NT - 1) / NT;
    constexpr int PR_PER = (TopkGeom<NW>::kPrivate + 31) / 32;
    uint64_t keep_sh[SH_PER], keep_pr[PR_PER];
#pragma unroll
    for (int j = 0; j < SH_PER; ++j) { const int i = tid + j * NT; keep_sh[j] = (i < n_sh) ? cbuf[i] : 0ull; }
#pragma unroll
    for (int j = 0; j < PR_PER; ++j) { const int i = lane + j * 32; keep_pr[j] = (i < n_priv) ? priv[i] : 0ull; }
    __syncthreads();
    if (tid == 0) st->cnt = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SH_PER; ++j)
        if (keep_sh[j] != 0ull && keep_sh[j] >= kth) cbuf[atomicAdd(&st->cnt, 1u)] = keep_sh[j];
#pragma unroll
    for (int j = 0; j < PR_PER; ++j)
        if (keep_pr[j] != 0ull && keep_pr[j] >= kth) cbuf[atomicAdd(&st->cnt, 1u)] = keep_pr[j];
    if (tid == 0 && kth) publish_tau(st, kth);
    __syncthreads();
}

// Poll + join protocol of a streaming warp.  `gate` was loaded at the start of the current group of windows; the warp
// joins when someone asked for a re-selection it has not served yet, or asks itself when its region cannot take
// `need` more keys.
template <int NT, int NW>
__device__ __forceinline__ void join_if_needed(uint64_t gate, uint32_t &epoch, int need, uint64_t *cbuf, int &n_priv,
                                               int k, uint32_t *hist, CtaState *st) {
    const bool full = n_priv + need > TopkGeom<NW>::kPrivate;
    if (gate_epoch(gate) != epoch || gate_epoch(gate_load(st)) != epoch || full) {
        if (full && (threadIdx.x & 31) == 0) *(volatile uint32_t *)&st->gate_epoch = epoch + 1;
        cta_join<NT, NW>(cbuf, n_priv, k, hist, st);
        n_priv = 0;
        ++epoch;
    }
}

// A warp that has finished streaming must keep serving join requests until every warp of the CTA is done.
template <int NT, int NW>
__device__ __forceinline__ void finish_streaming(uint32_t &epoch, uint64_t *cbuf, int &n_priv, int k, uint32_t *hist,
                                                 CtaState *st) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) atomicAdd(&st->done, 1u);
    for (;;) {
        if (gate_epoch(gate_load(st)) != epoch) {
            cta_join<NT, NW>(cbuf, n_priv, k, hist, st);
            n_priv = 0;
            ++epoch;
            continue;
        }
        if (*(volatile uint32_t *)&st->done >= (uint32_t)NW) break;
        __nanosleep(200);
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

// ---- K3 (inverted.cu): one CTA-WIDE append region cbuf[kSharedKeys, kCapMax) filled through st->n_app by any thread
// (the CTA runs in lockstep phases there, so no per-warp regions and no polling are needed).  kAppendCap = its size.

// Fold the append region into the shared set, keep the k best, raise tau, empty the region.  Whole CTA.
template <int NT, int kAppendCap>
__device__ __noinline__ void cta_join_flat(uint64_t *cbuf, uint64_t *app, int k, uint32_t *hist, CtaState *st) {
    const int tid = threadIdx.x;
    __syncthreads();
    const int n_sh = (int)st->cnt;
    const int n_app = min((int)st->n_app, kAppendCap);
    uint64_t kth = 0;
    if (n_sh + n_app > k) kth = radix_kth_largest2<true, true>(cbuf, n_sh, tid, NT, app, n_app, tid, NT, k, hist, tid, NT);
    constexpr int SH_PER = (kSharedKeys + NT - 1) / NT;
    constexpr int AP_PER = (kAppendCap + NT - 1) / NT;
    uint64_t keep_sh[SH_PER], keep_ap[AP_PER];
#pragma unroll
    for (int j = 0; j < SH_PER; ++j) { const int i = tid + j * NT; keep_sh[j] = (i < n_sh) ? cbuf[i] : 0ull; }
#pragma unroll
    for (int j = 0; j < AP_PER; ++j) { const int i = tid + j * NT; keep_ap[j] = (i < n_app) ? app[i] : 0ull; }
    __syncthreads();
    if (tid == 0) { st->cnt = 0; st->n_app = 0; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SH_PER; ++j)
        if (keep_sh[j] != 0ull && keep_sh[j] >= kth) cbuf[atomicAdd(&st->cnt, 1u)] = keep_sh[j];
#pragma unroll
    for (int j = 0; j < AP_PER; ++j)
        if (keep_ap[j] != 0ull && keep_ap[j] >= kth) cbuf[atomicAdd(&st->cnt, 1u)] = keep_ap[j];
    if (tid == 0 && kth) publish_tau(st, kth);
    __syncthreads();
}

// End of pass: exact top-k of the shared set plus the append region -> out[0..k) (unsorted, zero padded).  Whole CTA.
template <int NT, int kAppendCap>
__device__ __forceinline__ void cta_write_topk_flat(uint64_t *cbuf, const uint64_t *app, int k, uint32_t *hist, CtaState *st,
                                                    uint64_t *out) {
    const int tid = threadIdx.x;
    __syncthreads();
    const int n_sh = (int)st->cnt;
    const int n_app = min((int)st->n_app, kAppendCap);
    uint64_t kth = 0;
    if (n_sh + n_app > k) kth = radix_kth_largest2<true, true>(cbuf, n_sh, tid, NT, app, n_app, tid, NT, k, hist, tid, NT);
    __syncthreads();
    if (tid == 0) st->scratch = 0;
    __syncthreads();
    for (int i = tid; i < n_sh; i += NT) {
        const uint64_t x = cbuf[i];
        if (x != 0ull && x >= kth) out[atomicAdd(&st->scratch, 1u)] = x;
    }
    for (int i = tid; i < n_app; i += NT) {
        const uint64_t x = app[i];
        if (x != 0ull && x >= kth) out[atomicAdd(&st->scratch, 1u)] = x;
    }
    __syncthreads();
    for (int i = (int)st->scratch + tid; i < k; i += NT) out[i] = 0ull;  // fewer than k candidates in this CTA
}

}  // namespace vs
