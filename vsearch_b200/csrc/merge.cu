// merge.cu -- K6: merge P candidate lists per query into the final ranked top-k.
// Used twice: (a) per-CTA lists of one scan -> the shard's top-k; (b) the all-gathered per-rank
// lists of the row-sharded index -> the global top-k (no upstream counterpart; upstream searches a
// single device, index.py:179).  Keys are unique (ids are), so the k-th key is exact and the result
// is the canonical (score desc, id asc) order.
#include "index.cuh"

namespace vs {

constexpr int kMergeThreads = 512;

struct MergeParams {
    const uint64_t *in;
    const uint32_t *counts;   // optional [P, B]: only the first counts[p, b] entries of a list are valid
    int64_t P, stride_p, stride_b;
    int k_in, k_out;
    const int *alt_flag;      // optional device flag: != 0 -> the lists hold k_in_alt valid keys instead of k_in (auto mode:
    int k_in_alt;             // the inverted-list kernel wrote k keys per list, the scan kernel scan_kout)
    int64_t id_offset;
    int64_t *ids;      // [B, k_out] or nullptr
    float *scores;     // [B, k_out] or nullptr
    uint64_t *keys;    // [B, k_out] or nullptr (sorted keys, ids already offset)
};

__device__ __forceinline__ uint64_t merge_load(const MergeParams &p, int k_in, int64_t b, int64_t i) {
    int64_t pp = i / k_in, j = i - pp * k_in;
    if (p.counts != nullptr && j >= (int64_t)p.counts[pp * gridDim.x + b]) return 0ull;
    return p.in[pp * p.stride_p + b * p.stride_b + j];
}

// one CTA per query.  dynamic smem: sel[k_pow2] keys.
__global__ void __launch_bounds__(kMergeThreads) merge_topk_kernel(const MergeParams p, int k_pow2) {
    extern __shared__ __align__(16) uint8_t msmem[];
    uint64_t *sel = reinterpret_cast<uint64_t *>(msmem);
    __shared__ uint32_t hist[256];
    __shared__ uint32_t sel_cnt;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t b = blockIdx.x;
    const int k_in = (p.alt_flag != nullptr && *p.alt_flag != 0) ? p.k_in_alt : p.k_in;
    // a single counted list (the dense path's candidate lists: capacity 65,536, a few thousand valid): walk only the
    // valid prefix -- the capacity-sized loop was 11 of the 25 ms of a dense call on a 2.6M-row shard
    const int64_t n = (p.P == 1 && p.counts != nullptr) ? min((int64_t)p.counts[b], (int64_t)k_in) : p.P * k_in;

    // ---- radix-select the k_out-th largest key, streaming the lists from L2/HBM
    uint64_t prefix = 0, mask = 0;
    int rem = p.k_out;
    if (n > p.k_out) {
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += kMergeThreads) hist[i] = 0;
            __syncthreads();
            for (int64_t base = 0; base < n; base += kMergeThreads) {
                const int64_t i = base + tid;
                const uint64_t x = (i < n) ? merge_load(p, k_in, b, i) : 0ull;
                hist_add_aggregated(hist, (uint32_t)(x >> shift) & 255u, (i < n) && ((x & mask) == prefix));
            }
            __syncthreads();
            uint32_t h[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = hist[lane * 8 + j]; s += h[j]; }
            uint32_t incl = s;
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
            int src = __ffs(owner) - 1;
            digit = __shfl_sync(0xffffffffu, digit, src);
            newrem = __shfl_sync(0xffffffffu, newrem, src);
            prefix |= (uint64_t)digit << shift;
            mask |= (uint64_t)0xff << shift;
            rem = (int)newrem;
            __syncthreads();
        }
    }
    const uint64_t kth = prefix;  // 0 when n <= k_out: keep everything

    // ---- gather the winners, pad, bitonic sort descending
    if (tid == 0) sel_cnt = 0;
    for (int i = tid; i < k_pow2; i += kMergeThreads) sel[i] = 0;
    __syncthreads();
    for (int64_t i = tid; i < n; i += kMergeThreads) {
        uint64_t x = merge_load(p, k_in, b, i);
        if (x != 0 && x >= kth) {
            uint32_t pos = atomicAdd(&sel_cnt, 1u);
            if (pos < (uint32_t)k_pow2) sel[pos] = x;
        }
    }
    __syncthreads();
    for (int size = 2; size <= k_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (k_pow2 >> 1); i += kMergeThreads) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = sel[lo], c = sel[hi];
                if ((a < c) == desc) { sel[lo] = c; sel[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < p.k_out; i += kMergeThreads) {
        uint64_t x = sel[i];
        int64_t id = (int64_t)key_id(x) + p.id_offset;
        if (p.ids) p.ids[b * p.k_out + i] = id;
        if (p.scores) p.scores[b * p.k_out + i] = key_score(x);
        if (p.keys) p.keys[b * p.k_out + i] = (x & 0xffffffff00000000ull) | (uint64_t)(~(uint32_t)id);
    }
}

int launch_merge_counted(const uint64_t *d_in, const uint32_t *d_counts, int64_t P, int64_t stride_p, int64_t stride_b,
                         int64_t B, int k_in, int k_out, int64_t id_offset, int64_t *d_ids, float *d_scores,
                         uint64_t *d_keys, cudaStream_t st, const int *d_alt_flag = nullptr, int k_in_alt = 0);

// ---- register-resident, multi-stage variant ---------------------------------------------------------------------
// The streaming kernel above reads all P * k_in keys of a query nine times with ONE CTA: fine when a thousand queries
// keep every SM busy, but a single query against 148 lists of 1,000 keys (config 3, batch 1) spent 0.9 ms there for a
// 0.07 ms scoring kernel.  Here a CTA of 1,024 threads owns at most kMergeRegCap keys, loads them ONCE into registers
// (kMergeR per thread) and radix-selects from there; more keys than that are cut into groups of lists -- grid (B, G),
// every group writes its own top k_out -- and the group results are merged by the next stage.
constexpr int kMergeRegThreads = 1024;
constexpr int kMergeR = 16;
constexpr int kMergeRegCap = kMergeRegThreads * kMergeR;   // 16,384 keys per CTA

struct MergeStage {
    const uint64_t *in;
    int64_t P, stride_p, stride_b;
    int k_in;
    const int *alt_flag;
    int k_in_alt;
    int lists_per_group;
    int k_out;
    uint64_t *part_out;     // [B, G, k_out] unsorted keys (intermediate stage) or nullptr (final stage)
    int64_t id_offset;
    int64_t *ids; float *scores; uint64_t *keys;   // final stage outputs
};

__global__ void __launch_bounds__(kMergeRegThreads, 1) merge_reg_kernel(const MergeStage p, int k_pow2) {
    extern __shared__ __align__(16) uint8_t msmem[];
    uint64_t *sel = reinterpret_cast<uint64_t *>(msmem);
    __shared__ uint32_t hist[256];
    __shared__ uint32_t sel_cnt;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t b = blockIdx.x;
    const int g = blockIdx.y;
    const int k_in = (p.alt_flag != nullptr && *p.alt_flag != 0) ? p.k_in_alt : p.k_in;
    const int64_t l0 = (int64_t)g * p.lists_per_group;
    const int64_t nl = min((int64_t)p.lists_per_group, p.P - l0);
    const int n = (int)(nl * k_in);   // <= kMergeRegCap by construction
    uint64_t key[kMergeR];
#pragma unroll
    for (int j = 0; j < kMergeR; ++j) {
        const int i = tid + j * kMergeRegThreads;
        key[j] = 0ull;
        if (i < n) {
            const int pp = i / k_in, jj = i - pp * k_in;
            key[j] = p.in[(l0 + pp) * p.stride_p + b * p.stride_b + jj];
        }
    }
    uint64_t prefix = 0, mask = 0;
    int rem = p.k_out;
    if (n > p.k_out) {
        for (int shift = 56; shift >= 0; shift -= 8) {
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kMergeR; ++j)
                hist_add_aggregated(hist, (uint32_t)(key[j] >> shift) & 255u, key[j] != 0ull && ((key[j] & mask) == prefix));
            __syncthreads();
            uint32_t h[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = hist[lane * 8 + j]; s += h[j]; }
            uint32_t incl = s;
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
                    if (acc < (uint32_t)rem && acc + h[j] >= (uint32_t)rem) { digit = lane * 8 + j; newrem = rem - acc; hd = h[j]; }
                    acc += h[j];
                }
            }
            const uint32_t owner = __ballot_sync(0xffffffffu, mine);
            if (owner == 0) { prefix = 0; break; }   // fewer than k_out real keys (zeros are padding): keep them all
            const int src = __ffs(owner) - 1;
            digit = __shfl_sync(0xffffffffu, digit, src);
            newrem = __shfl_sync(0xffffffffu, newrem, src);
            hd = __shfl_sync(0xffffffffu, hd, src);
            prefix |= (uint64_t)digit << shift;
            mask |= (uint64_t)0xff << shift;
            rem = (int)newrem;
            __syncthreads();
            if (newrem == hd) break;   // the whole bucket is inside the top k_out: decided (keys >= prefix are exactly k_out)
        }
    }
    const uint64_t kth = prefix;  // keep keys >= kth (0: everything)
    if (tid == 0) sel_cnt = 0;
    if (p.part_out != nullptr) {
        uint64_t *out = p.part_out + (b * gridDim.y + g) * (int64_t)p.k_out;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kMergeR; ++j)
            if (key[j] != 0ull && key[j] >= kth) {
                const uint32_t pos = atomicAdd(&sel_cnt, 1u);
                if (pos < (uint32_t)p.k_out) out[pos] = key[j];
            }
        __syncthreads();
        for (int i = (int)min(sel_cnt, (uint32_t)p.k_out) + tid; i < p.k_out; i += kMergeRegThreads) out[i] = 0ull;
        return;
    }
    for (int i = tid; i < k_pow2; i += kMergeRegThreads) sel[i] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kMergeR; ++j)
        if (key[j] != 0ull && key[j] >= kth) {
            const uint32_t pos = atomicAdd(&sel_cnt, 1u);
            if (pos < (uint32_t)k_pow2) sel[pos] = key[j];
        }
    __syncthreads();
    for (int size = 2; size <= k_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (k_pow2 >> 1); i += kMergeRegThreads) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = sel[lo], c = sel[hi];
                if ((a < c) == desc) { sel[lo] = c; sel[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < p.k_out; i += kMergeRegThreads) {
        uint64_t x = sel[i];
        int64_t id = (int64_t)key_id(x) + p.id_offset;
        if (p.ids) p.ids[b * p.k_out + i] = id;
        if (p.scores) p.scores[b * p.k_out + i] = key_score(x);
        if (p.keys) p.keys[b * p.k_out + i] = (x & 0xffffffff00000000ull) | (uint64_t)(~(uint32_t)id);
    }
}

// Stage plan: groups of lists such that a group holds <= kMergeRegCap keys.  Returns false when the shape does not fit
// the register kernel at all (a single list longer than half the capacity): the streaming kernel serves it.
static bool merge_groups(int64_t P, int k_in, int *lists_per_group, int *n_groups) {
    if ((int64_t)k_in * 2 > kMergeRegCap && P > 1) return false;
    if ((int64_t)k_in > kMergeRegCap) return false;
    int lpg = (int)(kMergeRegCap / k_in);
    if (lpg > P) lpg = (int)P;
    *lists_per_group = lpg;
    *n_groups = (int)((P + lpg - 1) / lpg);
    return true;
}

// bytes of scratch launch_merge_staged needs for B queries (two ping-pong buffers of first-stage size)
size_t merge_scratch_bytes(int64_t P, int k_in, int k_out, int64_t B) {
    int lpg, G;
    if (!merge_groups(P, k_in, &lpg, &G) || G <= 1) return 0;
    return 2 * (size_t)B * G * (size_t)k_out * 8;
}

// Multi-stage merge of P lists of k_in keys per query (k_in_alt when *d_alt_flag != 0: the plan is made for the longer
// of the two).  d_scratch: merge_scratch_bytes() bytes (may be nullptr when that is 0).
int launch_merge_staged(const uint64_t *d_in, int64_t P, int64_t stride_p, int64_t stride_b, int64_t B, int k_in, int k_out,
                        int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, const int *d_alt_flag, int k_in_alt,
                        void *d_scratch, cudaStream_t st) {
    if (B == 0) return VS_OK;
    VS_REQUIRE(k_out >= 1 && k_out <= VS_MAX_K, VS_ERR_INVALID, "k=%d outside [1, %d]", k_out, VS_MAX_K);
    const int k_plan = (d_alt_flag != nullptr && k_in_alt > k_in) ? k_in_alt : k_in;
    int lpg, G;
    if (!merge_groups(P, k_plan, &lpg, &G) || (G > 1 && d_scratch == nullptr))
        return launch_merge_counted(d_in, nullptr, P, stride_p, stride_b, B, k_in, k_out, id_offset, d_ids, d_scores, d_keys, st,
                                    d_alt_flag, k_in_alt);
    int k_pow2 = 2;
    while (k_pow2 < k_out) k_pow2 <<= 1;
    MergeStage s;
    s.in = d_in; s.P = P; s.stride_p = stride_p; s.stride_b = stride_b; s.k_in = k_in; s.alt_flag = d_alt_flag; s.k_in_alt = k_in_alt;
    s.k_out = k_out; s.id_offset = id_offset;
    uint64_t *buf[2] = {(uint64_t *)d_scratch, d_scratch ? (uint64_t *)d_scratch + (size_t)B * G * k_out : nullptr};
    int which = 0;
    while (G > 1) {   // intermediate stages
        s.lists_per_group = lpg; s.part_out = buf[which]; s.ids = nullptr; s.scores = nullptr; s.keys = nullptr;
        merge_reg_kernel<<<dim3((unsigned)B, (unsigned)G), kMergeRegThreads, 0, st>>>(s, k_pow2);
        VS_CUDA(cudaGetLastError());
        // the next stage reads G lists of k_out keys per query
        s.in = buf[which]; s.P = G; s.stride_p = k_out; s.stride_b = (int64_t)G * k_out; s.k_in = k_out; s.alt_flag = nullptr;
        which ^= 1;
        if (!merge_groups(s.P, k_out, &lpg, &G)) return VS_ERR_UNSUPPORTED;   // cannot happen: k_out <= 2048
    }
    s.lists_per_group = lpg; s.part_out = nullptr; s.ids = d_ids; s.scores = d_scores; s.keys = d_keys;
    merge_reg_kernel<<<dim3((unsigned)B, 1), kMergeRegThreads, (size_t)k_pow2 * 8, st>>>(s, k_pow2);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

// lists of k_in valid keys, or of k_in_alt when *d_alt_flag != 0 (same strides)
int launch_merge_alt(const uint64_t *d_in, int64_t P, int64_t stride_p, int64_t stride_b, int64_t B, int k_in, int k_out,
                     int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, const int *d_alt_flag, int k_in_alt,
                     cudaStream_t st) {
    return launch_merge_counted(d_in, nullptr, P, stride_p, stride_b, B, k_in, k_out, id_offset, d_ids, d_scores, d_keys, st,
                                d_alt_flag, k_in_alt);
}

int launch_merge(const uint64_t *d_in, int64_t P, int64_t stride_p, int64_t stride_b, int64_t B, int k_in, int k_out,
                 int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, cudaStream_t st) {
    return launch_merge_counted(d_in, nullptr, P, stride_p, stride_b, B, k_in, k_out, id_offset, d_ids, d_scores, d_keys, st);
}

int launch_merge_counted(const uint64_t *d_in, const uint32_t *d_counts, int64_t P, int64_t stride_p, int64_t stride_b,
                         int64_t B, int k_in, int k_out, int64_t id_offset, int64_t *d_ids, float *d_scores,
                         uint64_t *d_keys, cudaStream_t st, const int *d_alt_flag, int k_in_alt) {
    if (B == 0) return VS_OK;
    VS_REQUIRE(k_out >= 1 && k_out <= VS_MAX_K, VS_ERR_INVALID, "k=%d outside [1, %d]", k_out, VS_MAX_K);
    int k_pow2 = 2;
    while (k_pow2 < k_out) k_pow2 <<= 1;
    MergeParams p;
    p.in = d_in; p.counts = d_counts; p.P = P; p.stride_p = stride_p; p.stride_b = stride_b;
    p.k_in = k_in; p.k_out = k_out; p.id_offset = id_offset;
    p.alt_flag = d_alt_flag; p.k_in_alt = k_in_alt;
    p.ids = d_ids; p.scores = d_scores; p.keys = d_keys;
    merge_topk_kernel<<<(unsigned)B, kMergeThreads, (size_t)k_pow2 * 8, st>>>(p, k_pow2);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vs
