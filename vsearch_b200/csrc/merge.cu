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
    int64_t id_offset;
    int64_t *ids;      // [B, k_out] or nullptr
    float *scores;     // [B, k_out] or nullptr
    uint64_t *keys;    // [B, k_out] or nullptr (sorted keys, ids already offset)
};

__device__ __forceinline__ uint64_t merge_load(const MergeParams &p, int64_t b, int64_t i) {
    int64_t pp = i / p.k_in, j = i - pp * p.k_in;
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
    // a single counted list (the dense path's candidate lists: capacity 65,536, a few thousand valid): walk only the
    // valid prefix -- the capacity-sized loop was 11 of the 25 ms of a dense call on a 2.6M-row shard
    const int64_t n = (p.P == 1 && p.counts != nullptr) ? min((int64_t)p.counts[b], (int64_t)p.k_in) : p.P * p.k_in;

    // ---- radix-select the k_out-th largest key, streaming the lists from L2/HBM
    uint64_t prefix = 0, mask = 0;
    int rem = p.k_out;
    if (n > p.k_out) {
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += kMergeThreads) hist[i] = 0;
            __syncthreads();
            for (int64_t base = 0; base < n; base += kMergeThreads) {
                const int64_t i = base + tid;
                const uint64_t x = (i < n) ? merge_load(p, b, i) : 0ull;
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
        uint64_t x = merge_load(p, b, i);
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
                         uint64_t *d_keys, cudaStream_t st);

int launch_merge(const uint64_t *d_in, int64_t P, int64_t stride_p, int64_t stride_b, int64_t B, int k_in, int k_out,
                 int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, cudaStream_t st) {
    return launch_merge_counted(d_in, nullptr, P, stride_p, stride_b, B, k_in, k_out, id_offset, d_ids, d_scores, d_keys, st);
}

int launch_merge_counted(const uint64_t *d_in, const uint32_t *d_counts, int64_t P, int64_t stride_p, int64_t stride_b,
                         int64_t B, int k_in, int k_out, int64_t id_offset, int64_t *d_ids, float *d_scores,
                         uint64_t *d_keys, cudaStream_t st) {
    if (B == 0) return VS_OK;
    VS_REQUIRE(k_out >= 1 && k_out <= VS_MAX_K, VS_ERR_INVALID, "k=%d outside [1, %d]", k_out, VS_MAX_K);
    int k_pow2 = 2;
    while (k_pow2 < k_out) k_pow2 <<= 1;
    MergeParams p;
    p.in = d_in; p.counts = d_counts; p.P = P; p.stride_p = stride_p; p.stride_b = stride_b;
    p.k_in = k_in; p.k_out = k_out; p.id_offset = id_offset;
    p.ids = d_ids; p.scores = d_scores; p.keys = d_keys;
    merge_topk_kernel<<<(unsigned)B, kMergeThreads, (size_t)k_pow2 * 8, st>>>(p, k_pow2);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vs
