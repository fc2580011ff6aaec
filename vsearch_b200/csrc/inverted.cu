// inverted.cu -- K3: token-major inverted-list scoring for sparse queries, sm_100a.
// Same contract as the scan (replaces upstream index.py:91-92) but touches only the posting lists of the query's
// non-zero tokens.  The lists are BLOCK-PARTITIONED: the rows are cut into blocks of R <= ~38,000 consecutive rows,
// and every block keeps its own token-major posting lists with 16-bit block-local row ids.  A block's fp32 score
// accumulator (R x 4 B) lives in SHARED memory, so scoring one query against one block is
//   zero the accumulator -> add w_t (x value) at every posting of the query's tokens (shared-memory atomics)
//   -> stream the accumulator through the fused top-k (topk.cuh)
// without a byte of accumulator traffic to L2/HBM (the first version kept an [N] fp32 row in global memory:
// RED.ADD.F32 into L2 + a read-and-clear pass over 2 x 4N bytes per query = 200 us per query at 21M rows).
//   build   : WS stream -> blk_rng[V][n_blocks] (uint2: start / end of the token's list inside the block; token-major so
//             that the bounds a CTA needs for its consecutive blocks share a 32-byte sector), blk_base[n_blocks] (uint64),
//             post_row[nnz] (uint16) (+ post_val[nnz]), post_ptr[V+1] = global list lengths (cost model) -- lazy
//   extract : prepared query [vpad] -> compact (token, weight) list + total postings (or the caller's lists as they are)
//   search  : grid (CTAs that own blocks, query lanes); CTA x owns blocks [x*bpc, (x+1)*bpc) and walks the queries
//             y, y + gridDim.y, ... -> one candidate list per (query, CTA) -> merge.cu
// Accumulate: one warp per list piece (a list, or half / a quarter of one when the query has few tokens), 32 consecutive
// postings per load, kInvUnroll loads in flight per lane; lists longer than kLongList postings (popular tokens) are
// stored transposed in 256-posting chunks and worked on by all warps, a chunk per 16-byte load.  Binary index, positive
// query weights: the weights go to fixed point (inv_fixed_point_kernel) and the adds are native integer shared-memory
// atomics instead of fp32 CAS loops; blocks after a CTA's first one then catch the rows whose sums cross the pre-filter
// while adding (the adds return the old sums) and the select visits only those.
// Thresholds: the score histogram of topk.cuh.  The first 3,072 rows of a CTA's first block are counted whole, the k-th
// bucket's lower bound becomes the float pre-filter; everything after counts and appends only what passes it, the bound
// is refreshed inside the first block and between blocks, and only rises.  Hits of the scan go to per-warp queues that
// are drained 32 at a time; the leftovers of a block are dealt to a few warps.  At the end everything >= the final bound
// (k .. ~2k keys) goes to the merge kernel.  The exact 64-bit machinery (CTA-wide radix selects) remains as the fallback
// for masses of equal scores / adversarial order.
// Rows never touched keep score 0 and compete like any other row.
// Algorithmic bytes per query: sum_t len(post_t) * (2 + b_val)  (+ 8 B of list bounds per (block, token)).
#include <cub/device/device_scan.cuh>

#include <stdlib.h>

#include <vector>

#include "index.cuh"
#include "topk.cuh"

namespace vs {

constexpr int kInvThreads = 768;      // 24 warps, one CTA per SM (shared memory bound)
constexpr int kInvWarps = kInvThreads / 32;
constexpr int kInvBuildThreads = 1024;
constexpr int kMaxQueryNnz = 4096;    // queries denser than this are served by the scan kernels
constexpr int kBlockRowsMax = 37120;  // accumulator rows per block: 145 KB of the 227 KB
constexpr int kInvAppend = 4096;      // keys in the CTA-wide append region (>= one replay step of kInvThreads * 4 rows)
constexpr int kInvQueue = 64;         // per-warp queue of accumulator float4s that passed the pre-filter (drained 32 at a time)
static_assert(kInvAppend >= kInvThreads * 4, "a replay step must fit the append region");
constexpr int kTokTile = 512;         // query tokens staged per pass over a block
constexpr uint32_t kLongList = 1024;  // postings; longer (block, token) lists are shared by all warps of the CTA
// Long lists (popular tokens: the rows of consecutive postings are consecutive or nearly so) start on a 16-byte
// boundary and are stored TRANSPOSED inside every full chunk of 256 postings: posting l + 32 j of the chunk sits at
// position 8 l + j.  One 16-byte load then hands lane l the postings l, l + 32, ..., l + 224, and the j-th atomics of a
// warp go to 32 consecutive rows = 32 different banks; read in list order, eight consecutive rows per lane, they would
// hit 4 banks 8 ways.  The tail (< 256 postings) stays in list order.
__host__ __device__ __forceinline__ uint32_t long_list_pos(uint32_t rel, uint32_t len) {
    return rel < (len & ~255u) ? ((rel & ~255u) | ((rel & 31u) << 3) | ((rel >> 5) & 7u)) : rel;
}

struct BlockLists {
    uint2 *blk_rng;        // [V, n_blocks]: (start, end) of the token's list inside the block's posting region
    uint64_t *blk_base;    // [n_blocks + 1]
    uint16_t *post_row;    // [nnz]
    void *post_val;        // [nnz] in store_dtype, or nullptr
    uint64_t *tok_total;   // [V + 1] global postings per token (counts, then exclusive prefix = post_ptr)
    int rows_per_block, n_blocks, V;
};

template <int NT>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp, uint32_t &block_total) {
    constexpr int NWARP = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NWARP; ++i) { uint32_t x = s_warp[i]; if (i < warp) base += x; tot += x; }
    __syncthreads();
    block_total = tot;
    return base + incl - v;
}

// ------------------------------------------------------------------------------------------- build
// One CTA per row block, one thread per row (rows are runs of 16-byte chunks in the WS stream, row_chunk[] = first
// chunk of each row).  Count pass: per-token histogram in shared memory -> exclusive offsets -> blk_rng[*][b];
// fill pass: the same walk with the list starts as cursors.
template <bool FILL>
__global__ void __launch_bounds__(kInvBuildThreads, 1) inv_build_kernel(const WsView idx, const BlockLists bl, int *err) {
    extern __shared__ uint32_t s_cnt[];  // V counters / cursors
    __shared__ uint32_t s_warp[kInvBuildThreads / 32];
    __shared__ unsigned long long s_total;
    const int V = bl.V, tid = threadIdx.x, b = blockIdx.x;
    if (tid == 0) s_total = 0;
    const size_t nb = (size_t)bl.n_blocks;
    uint32_t *s_long = s_cnt + V;   // FILL: one bit per token, set for long lists
    if constexpr (FILL) {
        for (int i = tid; i < (V + 31) / 32; i += kInvBuildThreads) s_long[i] = 0u;
        __syncthreads();
    }
    for (int i = tid; i < V; i += kInvBuildThreads) {
        if constexpr (FILL) {
            const uint2 rg = bl.blk_rng[(size_t)i * nb + b];
            s_cnt[i] = rg.x;
            if (rg.y - rg.x > kLongList) atomicOr(&s_long[i >> 5], 1u << (i & 31));
        } else {
            s_cnt[i] = 0u;
        }
    }
    __syncthreads();
    const int64_t r0 = (int64_t)b * bl.rows_per_block;
    const int64_t r1 = min(idx.n_rows, r0 + bl.rows_per_block);
    const uint64_t base = FILL ? bl.blk_base[b] : 0ull;
    for (int64_t r = r0 + tid; r < r1; r += kInvBuildThreads) {
        const uint32_t c0 = idx.row_chunk[r], c1 = idx.row_chunk[r + 1];
        for (uint32_t lc = c0; lc < c1; ++lc) {
            const uint64_t ch = ws_phys_chunk(lc, idx.cpl_shift);
            const uint4 u = idx.cols[ch];
            const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t c = ((e & 1) ? (wv[e >> 1] >> 16) : (wv[e >> 1] & 0xffffu)) & 0x7fffu;  // bit 15 = row-end flag
                if (c >= (uint32_t)V) continue;  // padding
                const uint32_t slot = atomicAdd(&s_cnt[c], 1u);
                if constexpr (FILL) {
                    uint64_t pos = base + slot;
                    if (s_long[c >> 5] & (1u << (c & 31))) {   // a long list: transposed chunks (long_list_pos)
                        const uint2 rg = __ldg(&bl.blk_rng[(size_t)c * nb + b]);
                        pos = base + rg.x + long_list_pos(slot - rg.x, rg.y - rg.x);
                    }
                    bl.post_row[pos] = (uint16_t)(r - r0);
                    if (idx.kind == 1) {
                        const uint64_t src = ch * 8ull + e;
                        if (idx.store_dtype == VS_F32) ((float *)bl.post_val)[pos] = ((const float *)idx.vals)[src];
                        else ((uint16_t *)bl.post_val)[pos] = ((const uint16_t *)idx.vals)[src];
                    }
                }
            }
        }
    }
    if constexpr (!FILL) {
        __syncthreads();
        const int per = (V + kInvBuildThreads - 1) / kInvBuildThreads;
        const int lo = min(V, tid * per), hi = min(V, lo + per);
        uint64_t sum = 0;   // slots: a long list gets 7 spare ones so that it can start on a 16-byte boundary
        for (int i = lo; i < hi; ++i) sum += s_cnt[i] + (s_cnt[i] > kLongList ? 7u : 0u);
        atomicAdd(&s_total, (unsigned long long)sum);
        uint32_t total;
        uint32_t run = block_exclusive_scan<kInvBuildThreads>((uint32_t)sum, s_warp, total);
        for (int i = lo; i < hi; ++i) {
            const uint32_t c = s_cnt[i];
            const bool lng = c > kLongList;
            if (lng) atomicOr(err, 2);   // the index has long lists (vs_index::inv_has_long)
            const uint32_t start = lng ? (run + 7u) & ~7u : run;
            bl.blk_rng[(size_t)i * nb + b] = make_uint2(start, start + c);
            run += c + (lng ? 7u : 0u);
            if (c) atomicAdd((unsigned long long *)&bl.tok_total[i], (unsigned long long)c);
        }
        __syncthreads();
        if (tid == 0) {
            if (s_total >= (1ull << 32) - 8ull) atomicOr(err, 1);
            bl.blk_base[b] = (s_total + 7ull) & ~7ull;  // sizes now (multiples of 8: every block starts aligned), exclusive prefix after the host-side scan
        }
    }
}

// Blocks are as large as the shared-memory accumulator allows: the per-(query, block) fixed costs (threshold sample,
// final select, list-pointer fetches) are what a small shard pays for, so fewer and longer blocks win; a query then
// occupies ceil(N / R) CTAs and the grid's query dimension fills the remaining SMs.
static void block_geometry(const vs_index *idx, int *rows_per_block, int *n_blocks, int *blocks_per_cta) {
    const int64_t N = idx->n_rows > 0 ? idx->n_rows : 1;
    int64_t r_max = kBlockRowsMax;
    if (const char *e = getenv("VSEARCH_B200_K3_BLOCK_ROWS")) {   // tests: several blocks per CTA on a small index
        const long v = atol(e);
        if (v >= 64 && v < kBlockRowsMax) r_max = v / 4 * 4;
    }
    int64_t R = N < r_max ? N : r_max;
    int64_t nb = (N + R - 1) / R;
    int64_t bpc = (nb + idx->n_ctas - 1) / idx->n_ctas;
    if (nb > idx->n_ctas) {   // a large index: equal blocks, the same number for every CTA
        R = (N + idx->n_ctas * bpc - 1) / (idx->n_ctas * bpc);
    }
    R = (R + 3) / 4 * 4;
    nb = (N + R - 1) / R;
    *rows_per_block = (int)R;
    *n_blocks = (int)nb;
    *blocks_per_cta = (int)bpc;
}

int build_inverted(vs_index *idx, cudaStream_t st) {
    if (idx->inv_built) return VS_OK;
    const int V = (int)idx->n_cols;
    const size_t smem = (size_t)V * 4 + (size_t)((V + 31) / 32) * 4;
    VS_REQUIRE(smem <= 200 * 1024, VS_ERR_UNSUPPORTED, "vocabulary too large for the inverted-list builder");
    block_geometry(idx, &idx->blk_rows, &idx->n_blocks, &idx->blocks_per_cta);
    const int nb = idx->n_blocks;
    int *d_err = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0, tmp2 = 0;
    auto cleanup = [&]() { cudaFree(d_err); cudaFree(d_tmp); };
    VS_CUDA(cudaMalloc(&d_err, 4));
    VS_CUDA(cudaMemsetAsync(d_err, 0, 4, st));
    VS_CUDA(cudaMalloc(&idx->post_ptr, (size_t)(V + 1) * 8));
    VS_CUDA(cudaMemsetAsync(idx->post_ptr, 0, (size_t)(V + 1) * 8, st));
    VS_CUDA(cudaMalloc(&idx->blk_ptr, (size_t)nb * V * sizeof(uint2)));
    VS_CUDA(cudaMalloc(&idx->blk_base, (size_t)(nb + 1) * 8));
    VS_CUDA(cudaMemsetAsync(idx->blk_base, 0, (size_t)(nb + 1) * 8, st));
    // postings + the alignment slack of long lists (< 7 per 1,025 postings) and of the blocks (< 8 each)
    const size_t n_slots = (size_t)idx->nnz + (size_t)idx->nnz / 128 + (size_t)nb * 8 + 64;
    VS_CUDA(cudaMalloc(&idx->post_row, n_slots * 2));
    const size_t vbytes = idx->kind == 1 ? (idx->store_dtype == VS_F32 ? 4 : 2) : 0;
    if (vbytes) VS_CUDA(cudaMalloc(&idx->post_val, n_slots * vbytes));
    BlockLists bl;
    bl.blk_rng = (uint2 *)idx->blk_ptr; bl.blk_base = idx->blk_base; bl.post_row = idx->post_row; bl.post_val = idx->post_val;
    bl.tok_total = idx->post_ptr; bl.rows_per_block = idx->blk_rows; bl.n_blocks = nb; bl.V = V;

    VS_CUDA(cudaFuncSetAttribute(inv_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_CUDA(cudaFuncSetAttribute(inv_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    inv_build_kernel<false><<<nb, kInvBuildThreads, smem, st>>>(ws_view(idx), bl, d_err);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, idx->post_ptr, idx->post_ptr, V + 1, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, idx->blk_base, idx->blk_base, nb + 1, st);
    tmp_bytes = tmp_bytes > tmp2 ? tmp_bytes : tmp2;
    VS_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, idx->post_ptr, idx->post_ptr, V + 1, st);
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, idx->blk_base, idx->blk_base, nb + 1, st);
    inv_build_kernel<true><<<nb, kInvBuildThreads, smem, st>>>(ws_view(idx), bl, d_err);
    int h_err = 0;
    cudaError_t e = cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    VS_CUDA(e);
    VS_REQUIRE((h_err & 1) == 0, VS_ERR_UNSUPPORTED, "a row block holds 2^32 or more postings");
    idx->inv_has_long = (h_err & 2) != 0;
    idx->inv_bytes = (int64_t)(n_slots * (2 + vbytes) + (size_t)nb * V * 8 + (size_t)(V + 1) * 8);
    idx->device_bytes += idx->inv_bytes;
    idx->inv_built = true;
    return VS_OK;
}

// ------------------------------------------------------------------------------------------- extract
// one CTA per query: compact the non-zero slots of the prepared query (ascending token order) and the
// exclusive prefix of their posting-list lengths.
struct QueryLists {
    uint32_t *tok;     // [B, kMaxQueryNnz]
    float *w;          // [B, kMaxQueryNnz]
    uint32_t *pref;    // [B, kMaxQueryNnz + 1]
    uint32_t *cnt;     // [B]   number of non-zero tokens (may exceed kMaxQueryNnz: then lists are truncated, unusable)
    uint64_t *total;   // [B]   total postings of the query
    int *fx;           // [B]   binary index: fixed-point exponent of the query's weights, or kNoFixed (inv_fixed_point_kernel)
};
constexpr int kNoFixed = 0x7fffffff;

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *s_warp, uint32_t &block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { uint32_t x = s_warp[i]; if (i < warp) base += x; tot += x; }
    __syncthreads();
    block_total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(256) inv_extract_kernel(const float *q, int vpad, int V, const uint64_t *post_ptr,
                                                          QueryLists L) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint64_t s_total;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *qb = q + (size_t)b * vpad;
    uint32_t *tok = L.tok + (size_t)b * kMaxQueryNnz;
    float *w = L.w + (size_t)b * kMaxQueryNnz;
    uint32_t *pref = L.pref + (size_t)b * (kMaxQueryNnz + 1);
    const int per = (V + 255) / 256;
    const int lo = tid * per, hi = min(V, lo + per);
    uint32_t n = 0;
    for (int i = lo; i < hi; ++i) n += (qb[i] != 0.0f);
    uint32_t total_nnz;
    uint32_t off = block_exclusive_scan_256(n, s_warp, total_nnz);
    for (int i = lo; i < hi; ++i) {
        float v = qb[i];
        if (v != 0.0f) {
            if (off < (uint32_t)kMaxQueryNnz) { tok[off] = (uint32_t)i; w[off] = v; }
            ++off;
        }
    }
    if (tid == 0) { L.cnt[b] = total_nnz; s_total = 0; }
    __syncthreads();
    unsigned long long my_total = 0;
    const uint32_t m = min(total_nnz, (uint32_t)kMaxQueryNnz);
    uint64_t run = 0;
    for (uint32_t base = 0; base < m; base += 256) {
        uint32_t i = base + tid;
        uint64_t len = 0;
        if (i < m) { uint32_t t = tok[i]; len = post_ptr[t + 1] - post_ptr[t]; my_total += len; }
        uint32_t tile_total;
        uint32_t ex = block_exclusive_scan_256((uint32_t)len, s_warp, tile_total);  // lists < 2^32 (checked at build)
        if (i < m) pref[i] = (uint32_t)(run + ex);
        run += tile_total;
    }
    // exact 64-bit total: when it is < 2^32 every 32-bit partial sum above was exact, otherwise the
    // lists are not used (inverted_usable)
    atomicAdd((unsigned long long *)&s_total, my_total);
    __syncthreads();
    if (tid == 0) { pref[m] = (uint32_t)run; L.total[b] = s_total; }
}

// Sparse queries given as CSR-style (token, weight) lists (vs_search_sparse; what the reference's top-k sparsifier
// produces, upstream utils/sparse.py:8-19): one CTA per query copies the valid entries (0 <= token < V, weight != 0
// after rounding to the index dtype) into the same QueryLists the extract kernel fills -- no dense [B, V] detour.
__global__ void __launch_bounds__(256) inv_lists_from_csr_kernel(const void *q_ptr, int ptr_dtype, const int32_t *q_tok,
                                                                 const float *q_w, int64_t b0, int V, int round_mode,
                                                                 const uint64_t *post_ptr, QueryLists L) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint64_t s_total;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int64_t lo = ptr_dtype == VS_I32 ? (int64_t)((const int32_t *)q_ptr)[b0 + b] : ((const int64_t *)q_ptr)[b0 + b];
    const int64_t hi = ptr_dtype == VS_I32 ? (int64_t)((const int32_t *)q_ptr)[b0 + b + 1] : ((const int64_t *)q_ptr)[b0 + b + 1];
    uint32_t *tok = L.tok + (size_t)b * kMaxQueryNnz;
    float *w = L.w + (size_t)b * kMaxQueryNnz;
    uint32_t *pref = L.pref + (size_t)b * (kMaxQueryNnz + 1);
    if (tid == 0) s_total = 0;
    __syncthreads();
    uint32_t n_out = 0;
    uint64_t run = 0;
    unsigned long long my_total = 0;
    for (int64_t base = lo; base < hi; base += 256) {
        const int64_t j = base + tid;
        int32_t t = -1;
        float v = 0.f;
        if (j < hi) { t = q_tok[j]; v = round_score(q_w[j], round_mode); }
        const bool keep = t >= 0 && t < V && v != 0.0f;
        uint64_t len = 0;
        if (keep) { len = post_ptr[t + 1] - post_ptr[t]; my_total += len; }
        uint32_t n_tile, len_tile;
        const uint32_t off = block_exclusive_scan_256(keep ? 1u : 0u, s_warp, n_tile);
        const uint32_t ex = block_exclusive_scan_256((uint32_t)len, s_warp, len_tile);
        const uint32_t at = n_out + off;
        if (keep && at < (uint32_t)kMaxQueryNnz) { tok[at] = (uint32_t)t; w[at] = v; pref[at] = (uint32_t)(run + ex); }
        n_out += n_tile;
        run += len_tile;
    }
    atomicAdd((unsigned long long *)&s_total, my_total);
    __syncthreads();
    if (tid == 0) {
        L.cnt[b] = n_out;
        pref[min(n_out, (uint32_t)kMaxQueryNnz)] = (uint32_t)run;
        L.total[b] = s_total;
    }
}

// Binary index, all query weights positive: a row's score is a sum of query weights, bounded by their total.  The weights
// go to 32-bit fixed point (scale 2^e, total <= 2^31) and K3's accumulator holds integers -- native shared-memory adds
// instead of CAS loops -- when that is at least as exact as the 1e-5 contract asks: every weight keeps >= 17 bits below
// its leading one, so each term, hence the sum, is within 2^-18 of its real value (the integer adds are exact; fp32
// accumulation itself drifts by ~n * 2^-24).  Otherwise (fx = kNoFixed) the query is accumulated in fp32.  A warp per query.
__global__ void __launch_bounds__(256) inv_fixed_point_kernel(QueryLists L, int Bc) {
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= Bc) return;
    const uint32_t cnt = L.cnt[q];
    const float *w = L.w + (size_t)q * kMaxQueryNnz;
    float tot = 0.f, mn = INFINITY;
    if (cnt <= (uint32_t)kTokTile)
        for (uint32_t i = lane; i < cnt; i += 32) { const float x = w[i]; tot += x; mn = fminf(mn, x); }   // a NaN weight makes tot NaN
    for (int d = 16; d; d >>= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, d); mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d)); }
    int e = kNoFixed;
    if (cnt > 0 && cnt <= (uint32_t)kTokTile && mn > 0.f && tot < 1e30f && tot > 1e-30f) {
        int x;
        (void)frexpf(tot * 1.0001f, &x);      // tot (whatever the summation order) < 2^x
        if (ldexpf(mn, 31 - x) >= 131072.f) e = 31 - x;
    }
    if (lane == 0) L.fx[q] = e;
}

// scan or inverted lists?  One CTA looks at the extracted queries of the chunk and writes the decision where the two
// scoring kernels and the merge read it -- the host never waits for it.  mode: VS_MODE_AUTO applies the cost model
// (seconds per query, constants measured on B200), VS_MODE_INVERTED takes the lists whenever they can serve the chunk
// (<= kMaxQueryNnz non-zeros and < 2^32 postings per query; otherwise the scan answers).
__global__ void __launch_bounds__(256) inv_decide_kernel(const QueryLists L, int Bc, int mode, double t_scan, double postings_per_sec,
                                                         double t_rows_fixed, int *flag, int *last_mode) {
    __shared__ uint32_t s_max_cnt;
    __shared__ unsigned long long s_max_tot, s_sum_tot;
    const int tid = threadIdx.x;
    if (tid == 0) { s_max_cnt = 0; s_max_tot = 0; s_sum_tot = 0; }
    __syncthreads();
    uint32_t mc = 0;
    unsigned long long mt = 0, st = 0;
    for (int i = tid; i < Bc; i += 256) {
        mc = max(mc, L.cnt[i]);
        const unsigned long long t = L.total[i];
        mt = max(mt, t);
        st += t;
    }
    atomicMax(&s_max_cnt, mc);
    atomicMax(&s_max_tot, mt);
    atomicAdd(&s_sum_tot, st);
    __syncthreads();
    if (tid == 0) {
        const bool usable = s_max_cnt <= (uint32_t)kMaxQueryNnz && s_max_tot < (1ull << 32);
        bool use = usable;
        if (mode == VS_MODE_AUTO) {
            const double t_inv = ((double)s_sum_tot / (double)Bc) / postings_per_sec + t_rows_fixed;
            use = usable && t_inv < t_scan;
        }
        *flag = use ? 1 : 0;
        *last_mode = use ? VS_MODE_INVERTED : VS_MODE_SCAN;
    }
}

// ------------------------------------------------------------------------------------------- search
struct InvSearchParams {
    QueryLists L;
    const uint2 *blk_rng;        // [V, n_blocks]
    const uint64_t *blk_base;
    const uint16_t *post_row;
    const void *post_val;
    int val_kind;        // 0 none (binary), 1 f32, 2 f16, 3 bf16
    int V, rows_per_block, n_blocks, blocks_per_cta;
    int64_t n_rows;
    uint64_t *cand;      // [B, gridDim.x, cand_stride]: up to kout keys per list, zero padded
    int k, kout, score_round, cand_stride;
    int n_queries;
    int n_lists;         // candidate lists per query the merge reads (= CTAs of the scan grid); this kernel's grid may be
                         // narrower (only CTAs that own row blocks are launched) and zero-fills the lists nobody writes
    int flags;           // experiment switches (VSEARCH_B200_K3_FLAGS): 1 = L1 prefetch of a tile's list heads, 2 = L2 prefetch of the
                         // next block's lists, 16 = no fixed-point accumulation on binary indices
    const int *use_inv;  // device flag written by inv_decide_kernel: 0 = the scan serves this chunk, this kernel exits
    unsigned long long *prof;   // diagnostic (vs_debug_scan_profile): per (query, CTA) nanoseconds spent per phase, or nullptr
};
// phases: 0 setup, 1 zero, 2 accumulate, 3 first-block histogram, 4 block select, 5 refresh / compaction, 6 final write, 7 total,
// 8-10 the first block's select: first rows / refreshes / rest
struct InvProf {
    unsigned long long t, acc[16];
    bool on;
    __device__ __forceinline__ static unsigned long long now() {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    }
    __device__ __forceinline__ void start(bool enabled) {
        on = enabled;
        if (on) { for (int i = 0; i < 16; ++i) acc[i] = 0; t = now(); acc[7] = t; }
    }
    __device__ __forceinline__ void lap(int phase) {
        if (on) { const unsigned long long n = now(); acc[phase] += n - t; t = n; }
    }
};

__device__ __forceinline__ float posting_value(const void *vals, int kind, uint64_t pos) {
    if (kind == 1) return ((const float *)vals)[pos];
    if (kind == 2) return __half2float(((const __half *)vals)[pos]);
    return __bfloat162float(((const __nv_bfloat16 *)vals)[pos]);
}

constexpr int kInvUnroll = 6;          // posting loads in flight per lane (6 x 32 covers the ~150-posting lists of the 21M-row configs in one go;
                                       // measured: 8 -> 6 = 33.8 -> 32.0 us per query on config 2, fewer dead slots and registers)

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t atom_shared_add(uint32_t *p, uint32_t v) {  // plain ATOMS.ADD (no compiler-made warp aggregation)
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t ordered_bits(float s) {   // high half of make_key(): larger = better, -0 folded into +0
    const uint32_t b = __float_as_uint(s + 0.0f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ void hist_count_plain(uint32_t *coarse, uint32_t *fine, uint32_t ob, uint32_t n = 1u) {
    atomicAdd(&fine[ob >> (32 - kHistFineBits)], n);
    atomicAdd(&coarse[ob >> 25], n);
}


// Add w * value at every posting of one list slice: offsets begin, begin + stride, ... < end of the list at `rows`
// (`vals`: its values, in the index dtype); kInvUnroll loads in flight per lane.  All lanes of a warp work on the SAME
// list, so their rows are distinct (no intra-warp conflicts) and the token's weight / list start are warp-uniform.
template <int VK>
__device__ __forceinline__ void accumulate_slice(float *acc, const uint16_t *__restrict__ rows, const void *__restrict__ vals,
                                                 uint32_t begin, uint32_t end, uint32_t stride, float w) {
    for (uint32_t off = begin; off < end; off += stride * kInvUnroll) {
        uint32_t row[kInvUnroll];
        float v[kInvUnroll];
#pragma unroll
        for (int u = 0; u < kInvUnroll; ++u) {
            const uint32_t o = off + (uint32_t)u * stride;
            row[u] = 0xffffffffu;
            v[u] = w;
            if (o < end) {
                row[u] = rows[o];
                if constexpr (VK == 1) v[u] *= ((const float *)vals)[o];
                else if constexpr (VK == 2) v[u] *= __half2float(((const __half *)vals)[o]);
                else if constexpr (VK == 3) v[u] *= __bfloat162float(((const __nv_bfloat16 *)vals)[o]);
            }
        }
#pragma unroll
        for (int u = 0; u < kInvUnroll; ++u)
            if (row[u] != 0xffffffffu) atomicAdd(&acc[row[u]], v[u]);
    }
}

// The same for a binary index whose query weights went to fixed point (see `fx` in inv_search_kernel): every posting of
// the list adds the token's integer weight, and 32-bit integer adds ARE native on shared memory (ATOMS.ADD, no CAS loop:
// three times the CAS ceiling in scripts/micro/smem_atomics.cu).
constexpr uint32_t kCrossCap = kInvWarps * kInvQueue;   // rows the crossing queue holds (it lives in the hit queues' memory)
constexpr uint32_t kCrossMaxPostings = 24576;           // per (query, block): beyond, scanning the sums is cheaper than returning adds

// CROSS: the adds return the old sums, and a row whose sum climbs over the pre-filter `tau_u` with this posting (weights
// are positive: that happens once per row) is noted in the CTA's crossing queue -- the select then has nothing to scan.
template <bool CROSS>
__device__ __forceinline__ void accumulate_slice_fixed(uint32_t *acc, const uint16_t *__restrict__ rows, uint32_t begin, uint32_t end,
                                                       uint32_t stride, uint32_t wq, uint32_t tau_u, uint16_t *cross_q, uint32_t *cross_n) {
    for (uint32_t off = begin; off < end; off += stride * kInvUnroll) {
        uint32_t row[kInvUnroll];
#pragma unroll
        for (int u = 0; u < kInvUnroll; ++u) {
            const uint32_t o = off + (uint32_t)u * stride;
            row[u] = o < end ? (uint32_t)rows[o] : 0xffffffffu;
        }
        if constexpr (CROSS) {
            uint32_t old[kInvUnroll];
#pragma unroll
            for (int u = 0; u < kInvUnroll; ++u) old[u] = row[u] != 0xffffffffu ? atomicAdd(&acc[row[u]], wq) : 0xffffffffu;
#pragma unroll
            for (int u = 0; u < kInvUnroll; ++u)
                if (old[u] < tau_u && old[u] + wq >= tau_u) {   // (sums stay below 2^32: no wrap; idle slots hold ~0)
                    const uint32_t slot = atomicAdd(cross_n, 1u);
                    if (slot < kCrossCap) cross_q[slot] = (uint16_t)row[u];
                }
        } else {
#pragma unroll
            for (int u = 0; u < kInvUnroll; ++u)
                if (row[u] != 0xffffffffu) atomicAdd(&acc[row[u]], wq);
        }
    }
}

// The short lists of one token tile, fixed point, G pieces of a warp at a time: the postings of all G pieces are asked for
// before the first add of the group, so a warp waits for one load round trip per G pieces instead of one per piece.
#ifndef VS_K3_GROUP
#define VS_K3_GROUP 3
#endif
template <bool CROSS, int G>
__device__ __forceinline__ void accumulate_pieces_fixed(uint32_t *accu, const uint16_t *__restrict__ post_row, const uint64_t base,
                                                        const uint32_t *s_beg, const uint32_t *s_len, const float *s_w, const float fx_scale,
                                                        const int tn, const int psh, const int warp, const int lane,
                                                        const uint32_t tau_u, uint16_t *cross_q, uint32_t *cross_n) {
    const int n_it = tn << psh;
    for (int it0 = warp; it0 < n_it; it0 += G * kInvWarps) {
        uint32_t row[G][kInvUnroll], wq[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int it = it0 + g * kInvWarps;
            const int ti = min(it >> psh, tn - 1);
            const uint32_t pc = (uint32_t)(it & ((1 << psh) - 1)), len = s_len[ti];
            const bool ok = it < n_it && len - 1u < kLongList;
            const uint32_t lo = ok ? (len * pc) >> psh : 0u, hi = ok ? (len * (pc + 1u)) >> psh : 0u;
            const uint16_t *rows = post_row + base + s_beg[ti];
            wq[g] = __float2uint_rn(s_w[ti] * fx_scale);
#pragma unroll
            for (int u = 0; u < kInvUnroll; ++u) {
                const uint32_t o = lo + (uint32_t)lane + 32u * (uint32_t)u;
                row[g][u] = o < hi ? (uint32_t)rows[o] : 0xffffffffu;
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if constexpr (CROSS) {
                uint32_t old[kInvUnroll];
#pragma unroll
                for (int u = 0; u < kInvUnroll; ++u) old[u] = row[g][u] != 0xffffffffu ? atomicAdd(&accu[row[g][u]], wq[g]) : 0xffffffffu;
#pragma unroll
                for (int u = 0; u < kInvUnroll; ++u)
                    if (old[u] < tau_u && old[u] + wq[g] >= tau_u) {
                        const uint32_t slot = atomicAdd(cross_n, 1u);
                        if (slot < kCrossCap) cross_q[slot] = (uint16_t)row[g][u];
                    }
            } else {
#pragma unroll
                for (int u = 0; u < kInvUnroll; ++u)
                    if (row[g][u] != 0xffffffffu) atomicAdd(&accu[row[g][u]], wq[g]);
            }
            // what a piece has beyond 32 * kInvUnroll postings goes the plain way
            const int it = it0 + g * kInvWarps;
            if (it < n_it) {
                const int ti = it >> psh;
                const uint32_t pc = (uint32_t)(it & ((1 << psh) - 1)), len = s_len[ti];
                const uint32_t lo = (len * pc) >> psh, hi = (len * (pc + 1u)) >> psh;
                if (len - 1u < kLongList && hi - lo > 32u * kInvUnroll)
                    accumulate_slice_fixed<CROSS>(accu, post_row + base + s_beg[ti], lo + 32u * kInvUnroll + (uint32_t)lane, hi, 32u, wq[g], tau_u, cross_q, cross_n);
            }
        }
    }
}

// A LONG list (more postings than kLongList: a popular token), worked on by all warps of the CTA.  Its full 256-posting
// chunks are stored transposed (long_list_pos): a warp takes a chunk with ONE 16-byte load per lane (the scalar loop has 8
// two-byte loads in flight per lane), lane l gets postings l, l + 32, ..., and the warp's j-th atomics go to consecutive
// rows.  MODE 0: fp32 sums (CAS loop; values of kind VK), 1: fixed point, 2: fixed point + crossing detection.
template <int VK, int MODE>
__device__ __noinline__ void accumulate_long(float *acc, const uint16_t *__restrict__ rows, const void *__restrict__ vals,
                                                const uint32_t len, const float w, const uint32_t wq, const int tid, const uint32_t tau_u,
                                                uint16_t *cross_q, uint32_t *cross_n) {
    uint32_t *accu = reinterpret_cast<uint32_t *>(acc);
    const int lane = tid & 31, warp = tid >> 5;
    auto add = [&](const uint32_t r, const float v) {
        if constexpr (MODE == 0) {
            atomicAdd(&acc[r], v);
        } else if constexpr (MODE == 1) {
            atomicAdd(&accu[r], wq);
        } else {
            const uint32_t old = atomicAdd(&accu[r], wq);
            if (old < tau_u && old + wq >= tau_u) {
                const uint32_t slot = atomicAdd(cross_n, 1u);
                if (slot < kCrossCap) cross_q[slot] = (uint16_t)r;
            }
        }
    };
    auto value = [&](const uint32_t pos) -> float {
        if constexpr (MODE != 0 || VK == 0) return w;
        else if constexpr (VK == 1) return w * ((const float *)vals)[pos];
        else if constexpr (VK == 2) return w * __half2float(((const __half *)vals)[pos]);
        else return w * __bfloat162float(((const __nv_bfloat16 *)vals)[pos]);
    };
    const uint32_t n_chunks = len >> 8, full = n_chunks << 8;
    for (uint32_t i = full + (uint32_t)tid; i < len; i += kInvThreads) add(rows[i], value(i));   // tail, in list order
    constexpr int U = (MODE == 0 && VK == 1) ? 1 : 2;   // chunks in flight per warp (register budget: 80 per thread)
    for (uint32_t c0 = (uint32_t)warp; c0 < n_chunks; c0 += kInvWarps * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t c = c0 + (uint32_t)u * kInvWarps;
            if (c < n_chunks) v[u] = __ldg(reinterpret_cast<const uint4 *>(rows + ((size_t)c << 8)) + lane);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t c = c0 + (uint32_t)u * kInvWarps;
            if (c < n_chunks) {
                const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                const uint32_t p0 = (c << 8) + ((uint32_t)lane << 3);   // this lane's 8 positions
                if constexpr (MODE == 0 && VK == 1) {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>((const float *)vals + p0));
                    const float4 b2 = __ldg(reinterpret_cast<const float4 *>((const float *)vals + p0) + 1);
                    const float f[8] = {a.x, a.y, a.z, a.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) add((e & 1) ? (w4[e >> 1] >> 16) : (w4[e >> 1] & 0xffffu), w * f[e]);
                } else if constexpr (MODE == 0 && (VK == 2 || VK == 3)) {
                    const uint4 hv = __ldg(reinterpret_cast<const uint4 *>((const uint16_t *)vals + p0));
                    const uint32_t h4[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint16_t bits = (uint16_t)((e & 1) ? (h4[e >> 1] >> 16) : (h4[e >> 1] & 0xffffu));
                        float fv;
                        if constexpr (VK == 2) fv = __half2float(__ushort_as_half(bits));
                        else fv = __uint_as_float((uint32_t)bits << 16);
                        add((e & 1) ? (w4[e >> 1] >> 16) : (w4[e >> 1] & 0xffffu), w * fv);
                    }
                } else if constexpr (MODE == 2) {   // all eight adds go out before the first old sum is looked at
                    uint32_t old[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) old[e] = atomicAdd(&accu[(e & 1) ? (w4[e >> 1] >> 16) : (w4[e >> 1] & 0xffffu)], wq);
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (old[e] < tau_u && old[e] + wq >= tau_u) {
                            const uint32_t slot = atomicAdd(cross_n, 1u);
                            if (slot < kCrossCap) cross_q[slot] = (uint16_t)((e & 1) ? (w4[e >> 1] >> 16) : (w4[e >> 1] & 0xffffu));
                        }
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) add((e & 1) ? (w4[e >> 1] >> 16) : (w4[e >> 1] & 0xffffu), w);
                }
            }
        }
    }
}

// Drop the keys of the append region that fell below the histogram bound `ob` (ordered score bits): called by all
// threads of the CTA with the same (n_app, ob), nobody appending meanwhile.  Ends with a barrier.
static __device__ __noinline__ void compact_appended(uint64_t *app, CtaState *st, const uint32_t n_app, const uint32_t ob) {
    constexpr int NT = kInvThreads, PER = (kInvAppend + NT - 1) / NT;
    const int tid = threadIdx.x;
    uint64_t keep[PER];
    const uint64_t tmin = (uint64_t)ob << 32;
#pragma unroll
    for (int j = 0; j < PER; ++j) { const int i = tid + j * NT; keep[j] = (i < (int)n_app) ? app[i] : 0ull; }
    __syncthreads();
    if (tid == 0) st->n_app = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j)
        if (keep[j] != 0ull && keep[j] >= tmin) app[atom_shared_add(&st->n_app, 1u)] = keep[j];
    __syncthreads();
}

// LONG: the index has lists longer than kLongList (popular tokens) -- their vectorised all-warps path is compiled in.  It
// costs the rest of the kernel registers and ~8 % of its speed, so an index without such lists runs the variant without.
// PROF: the phase timers of vs_debug_scan_profile (compiled for the binary and the fp32-valued unrounded variants only).
template <int VK, bool ROUND, bool LONG, bool PROF = false>
__global__ void __launch_bounds__(kInvThreads, 1) inv_search_kernel(const InvSearchParams p) {
    constexpr int NT = kInvThreads, NW = kInvWarps;
    constexpr int vbytes = VK == 0 ? 0 : (VK == 1 ? 4 : 2);
    extern __shared__ __align__(128) uint8_t ssmem[];
    // the score histogram and the exact fallback's shared top-k set share their memory: a query that has to fall back
    // to exact CTA-wide selections (masses of equal scores, adversarial row order) stops using the histogram
    uint32_t *fine = reinterpret_cast<uint32_t *>(ssmem);                       // kHistFine counters (32 KB) ...
    uint64_t *shset = reinterpret_cast<uint64_t *>(ssmem);                      // ... or kSharedKeys keys (16 KB)
    uint32_t *coarse = fine + kHistFine;                                        // kHistCoarse
    uint32_t *hist = coarse + kHistCoarse;                                      // 256: radix-select scratch of the fallback
    uint64_t *app = reinterpret_cast<uint64_t *>(hist + 256);                   // kInvAppend keys: CTA-wide append region
    uint32_t *s_tok = reinterpret_cast<uint32_t *>(app + kInvAppend);           // kTokTile each
    float *s_w = reinterpret_cast<float *>(s_tok + kTokTile);
    uint32_t *s_beg = reinterpret_cast<uint32_t *>(s_w + kTokTile);
    uint32_t *s_len = s_beg + kTokTile;
    uint16_t *s_queue = reinterpret_cast<uint16_t *>(s_len + kTokTile);         // kInvWarps x kInvQueue
    float *acc = reinterpret_cast<float *>(s_queue + kInvWarps * kInvQueue);    // rows_per_block
    __shared__ CtaState st;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (*p.use_inv == 0) return;
    // persistent over the queries: CTA (x, y) scores queries y, y + gridDim.y, ... against its row blocks (a fresh CTA
    // with 227 KB of shared memory per (query, block group) left the SMs idle half the time on a 71-block shard)
    for (int q = blockIdx.y; q < p.n_queries; q += gridDim.y) {
    const int cnt = (int)min(p.L.cnt[q], (uint32_t)kMaxQueryNnz);
    const uint32_t *tok = p.L.tok + (size_t)q * kMaxQueryNnz;
    const float *w = p.L.w + (size_t)q * kMaxQueryNnz;
    const int R = p.rows_per_block;
    const int blk0 = blockIdx.x * p.blocks_per_cta, blk1 = min(p.n_blocks, blk0 + p.blocks_per_cta);
    const bool cached = cnt <= kTokTile;   // the whole token list stays in shared memory across blocks
    const size_t nb = (size_t)p.n_blocks;
    __shared__ uint32_t s_ob, s_qn[kInvWarps], s_cross_n;
    InvProf prof;
    prof.start(PROF && p.prof != nullptr && tid == 0);
    if (tid == 0) cta_state_reset(&st);
    for (int i = tid; i < kHistFine + kHistCoarse; i += NT) fine[i] = 0u;   // coarse follows fine
    if (cached && tid < cnt) { s_tok[tid] = tok[tid]; s_w[tid] = w[tid]; }
    __syncthreads();
    bool fixed = false;
    float fx_scale = 1.f, fx_inv = 1.f;
    if constexpr (VK == 0) {   // fixed-point weights (inv_fixed_point_kernel)
        const int e = p.L.fx[q];
        if (e != kNoFixed && !(p.flags & 16)) { fixed = true; fx_scale = ldexpf(1.f, e); fx_inv = ldexpf(1.f, -e); }
    }
    uint32_t *accu = reinterpret_cast<uint32_t *>(acc);
    const uint4 *acc4u = reinterpret_cast<const uint4 *>(acc);
    // the largest integer pre-filter that still lets through every row whose (rounded) float score can reach t
    auto fx_floor = [&](const float t) -> uint32_t {
        if (!(t > 0.f)) return 0u;
        float lo = t;
        if constexpr (ROUND) lo = t * (1.f - 0.0078125f) - 6e-8f;
        const float x = lo * fx_scale * (1.f - 1e-6f) - 1.f;
        return x > 0.f ? (uint32_t)x : 0u;
    };
    // list bounds of token `tid` in the NEXT block to score: loaded one block ahead, and the lists themselves are
    // pulled into L2 while the previous block is being selected
    uint2 nrng = make_uint2(0u, 0u);
    uint64_t nbase = 0;
    if (cached && blk0 < blk1 && tid < cnt) nrng = p.blk_rng[(size_t)s_tok[tid] * nb + blk0];
    float4 *acc4 = reinterpret_cast<float4 *>(acc);
    float tau_s = -INFINITY;               // float pre-filter: lower bound of the histogram's k-th bucket
    bool booted = false;                   // the CTA's first block has set the pre-filter
    bool exact = false;                    // exact fallback engaged: histogram memory now holds the shared top-k set

    auto rnd = [&](float x) -> float { if constexpr (ROUND) return round_score(x, p.score_round); else return x; };
    // The slow path of the select, for up to 32 accumulator float4s at once (a lane with `have` takes float4 `i`): rows at
    // or above the pre-filter and above the exact threshold (if any) are counted in the histogram (rows >= count_from
    // only) and appended -- one ATOMS.ADD per warp for the slots.  Called by all 32 lanes.
    auto take_rows = [&](const bool have, const int i, const int rows_b, const int64_t row0, const uint64_t tau, const int count_from) {
        uint64_t key[4] = {0ull, 0ull, 0ull, 0ull};
        uint32_t n = 0;
        if (have) {
            const int r = i * 4;
            float4 v = acc4[i];
            if (fixed) {
                const uint4 u = acc4u[i];
                v = make_float4(__uint2float_rn(u.x) * fx_inv, __uint2float_rn(u.y) * fx_inv, __uint2float_rn(u.z) * fx_inv, __uint2float_rn(u.w) * fx_inv);
            }
            const float sc[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float se = rnd(sc[e]);
                if (r + e < rows_b && se >= tau_s) {
                    const uint64_t ke = make_key(se, (uint32_t)(row0 + r + e));
                    if (ke > tau) {
                        key[e] = ke; ++n;
                        if (!exact && r + e >= count_from) hist_count_plain(coarse, fine, (uint32_t)(ke >> 32));
                    }
                }
            }
        }
        uint32_t incl = n;   // slots: exclusive prefix over the lanes, one atomic for the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) return;
        uint32_t base_slot = 0;
        if (lane == 0) base_slot = atom_shared_add(&st.n_app, total);
        uint32_t slot = __shfl_sync(0xffffffffu, base_slot, 0) + incl - n;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (key[e] != 0ull) { if (slot < (uint32_t)kInvAppend) app[slot] = key[e]; ++slot; }
    };
    // float4 indices [i_lo, i_hi) of the block's accumulator through the pre-filter: the loop the select spends its time in.
    // Hits (one float4 in ~60 at best, so some lane of most warp iterations has one) are not handled where they occur --
    // a divergent slow path per hit -- but queued per warp and taken 32 at a time with all lanes busy.  What is left in
    // the queues (`qn` entries) stays there across calls; flush_cta() / flush_own() take it at the end of the block.
    uint16_t *const qw = s_queue + warp * kInvQueue;
    int qn = 0;
    auto select_range = [&](const int i_lo, const int i_hi, const int rows_b, const int64_t row0, const uint64_t tau, const int count_from) {
        const uint32_t tau_u = fx_floor(tau_s);
        for (int i0 = i_lo; i0 < i_hi; i0 += NT) {
            const int i = i0 + tid;
            bool hit = false;
            if (i < i_hi) {
                if (fixed) {
                    const uint4 u = acc4u[i];
                    hit = max(max(u.x, u.y), max(u.z, u.w)) >= tau_u;
                } else {
                    const float4 v = acc4[i];
                    hit = rnd(fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w))) >= tau_s;   // rounding is monotonic
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                if (hit) qw[qn + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
                qn += __popc(m);
                if (qn >= 32) {
                    __syncwarp();
                    take_rows(true, qw[lane], rows_b, row0, tau, count_from);
                    const uint16_t rest = lane < qn - 32 ? qw[32 + lane] : (uint16_t)0;
                    __syncwarp();
                    if (lane < qn - 32) qw[lane] = rest;
                    qn -= 32;
                }
            }
        }
    };
    auto flush_own = [&](const int rows_b, const int64_t row0, const uint64_t tau, const int count_from) {
        __syncwarp();
        if (qn) take_rows(lane < qn, qw[lane < qn ? lane : 0], rows_b, row0, tau, count_from);
        __syncwarp();
        qn = 0;
    };
    // The leftovers of all warps' queues (a handful per warp at the end of a block) taken together: 24 warps running the
    // slow path for five entries each cost as much as for 32 each, so the entries are dealt to the first
    // ceil(total / 32) warps instead.  Called by all threads; contains one barrier, the caller adds the one after.
    auto flush_cta = [&](const int rows_b, const int64_t row0, const uint64_t tau, const int count_from) {
        if (lane == 0) s_qn[warp] = (uint32_t)qn;
        __syncthreads();
        const uint32_t c = lane < NW ? s_qn[lane] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        qn = 0;
        if ((uint32_t)warp * 32u >= total) return;
        const uint32_t g = min((uint32_t)warp * 32u + (uint32_t)lane, total - 1u);
        int lo = 0, hi = NW - 1;   // the warp whose queue holds entry g: the first one with incl > g
#pragma unroll
        for (int s5 = 0; s5 < 5; ++s5) {
            const int mid = (lo + hi) >> 1;
            if (__shfl_sync(0xffffffffu, incl, mid) > g) hi = mid; else lo = mid + 1;
        }
        const uint32_t first = __shfl_sync(0xffffffffu, incl, lo) - __shfl_sync(0xffffffffu, c, lo);
        take_rows((uint32_t)warp * 32u + (uint32_t)lane < total, s_queue[lo * kInvQueue + (int)(g - first)], rows_b, row0, tau, count_from);
    };
    // A row of the crossing queue (fixed-point blocks): its final sum -> key -> histogram and append region.  All 32 lanes.
    auto take_crossed = [&](const bool have, const uint32_t r, const int rows_b, const int64_t row0, const uint64_t tau) {
        uint64_t key = 0ull;
        if (have && (int)r < rows_b) {
            const float se = rnd(__uint2float_rn(accu[r]) * fx_inv);
            if (se >= tau_s) {
                const uint64_t ke = make_key(se, (uint32_t)(row0 + r));
                if (ke > tau) { key = ke; hist_count_plain(coarse, fine, (uint32_t)(ke >> 32)); }
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, key != 0ull);
        if (m == 0u) return;
        uint32_t base_slot = 0;
        if (lane == 0) base_slot = atom_shared_add(&st.n_app, (uint32_t)__popc(m));
        const uint32_t slot = __shfl_sync(0xffffffffu, base_slot, 0) + (uint32_t)__popc(m & ((1u << lane) - 1u));
        if (key != 0ull && slot < (uint32_t)kInvAppend) app[slot] = key;
    };
    // raise the pre-filter to the histogram's k-th bucket: one warp asks the histogram, everybody reads the answer.
    // Called by all threads; a barrier before the call makes the counts of the rows processed so far visible.
    uint32_t n_app_seen = 0;   // the append count at the last refresh (the same value in every thread)
    auto refresh = [&]() -> uint32_t {
        if (warp == 0) {
            const uint32_t ob = hist_threshold(coarse, fine, p.k);
            if (lane == 0) s_ob = ob;
        }
        n_app_seen = *(volatile uint32_t *)&st.n_app;   // nobody appends between the caller's barrier and this one
        __syncthreads();
        const uint32_t ob = s_ob;
        if (ob) tau_s = fmaxf(tau_s, key_score((uint64_t)ob << 32));
        return ob;
    };

    for (int b = blk0; b < blk1; ++b) {
        const int64_t row0 = (int64_t)b * R;
        const int rows_b = (int)min((int64_t)R, p.n_rows - row0);
        const uint64_t base = p.blk_base[b];
        prof.lap(0);
        // Fixed-point blocks after the CTA's first one: the pre-filter is known before the postings are added, so the rows
        // that end up above it are caught while they cross it (accumulate_slice_fixed<true>) and the block's select
        // only visits those instead of scanning 37 K sums for the ~150 that matter.
        const uint32_t cross_tau = fx_floor(tau_s);
        // ... when the block has few enough postings for this query: the adds that return the old sum cost ~70 ps more per
        // posting than the fire-and-forget ones, the scan they save ~2.3 us per block (heavy-tailed lists: 600 K postings)
        bool cross = fixed && booted && !exact && cached && cross_tau > 0u;
        if (tid == 0) s_cross_n = 0u;   // (ordered before the first add by the barrier below)
        for (int i = tid; i < (R >> 2); i += NT) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        // ---- accumulate: tiles of <= kTokTile query tokens
        for (int t0 = 0; t0 < cnt; t0 += kTokTile) {
            const int tn = min(kTokTile, cnt - t0);
            uint2 rng = nrng;
            if (!cached) {
                __syncthreads();  // previous tile fully consumed
                if (tid < tn) {
                    const uint32_t t = tok[t0 + tid];
                    s_tok[tid] = t; s_w[tid] = w[t0 + tid];
                    rng = p.blk_rng[(size_t)t * nb + b];
                }
            } else if (b + 1 < blk1 && tid < cnt) {  // in flight under this block's accumulate
                nrng = p.blk_rng[(size_t)s_tok[tid] * nb + b + 1];
                nbase = p.blk_base[b + 1];
            }
            if (tid < tn) { s_beg[tid] = rng.x; s_len[tid] = rng.y - rng.x; }
            __syncthreads();  // also orders the zeroing above before the first atomic
            if constexpr (LONG) {   // (an index without long lists has at most 1,024 postings per token and block)
                if (cross) {   // (cached: this is the only tile) every warp adds up the block's postings for itself
                    uint32_t n_post = 0;
                    for (int i = lane; i < tn; i += 32) n_post += s_len[i];
                    n_post = __reduce_add_sync(0xffffffffu, n_post);
                    cross = n_post <= kCrossMaxPostings;
                }
            }
            prof.lap(1);
            if (p.flags & 1) {   // the heads of this tile's lists -> L1 (a warp walks its lists one after the other)
                for (int i = tid; i < tn * 8; i += NT) {
                    const int ti = i >> 3;
                    const uint32_t ln = (uint32_t)(i & 7), len = min(s_len[ti], kLongList);
                    const uint64_t at = base + s_beg[ti];
                    const uintptr_t r0 = reinterpret_cast<uintptr_t>(p.post_row + at);
                    if (((r0 & ~(uintptr_t)127) + ln * 128u) < r0 + len * 2u) prefetch_l1(reinterpret_cast<const void *>((r0 & ~(uintptr_t)127) + ln * 128u));
                    if (vbytes) {
                        const uintptr_t v0 = reinterpret_cast<uintptr_t>(p.post_val) + at * vbytes;
                        for (uint32_t l2 = ln; ((v0 & ~(uintptr_t)127) + l2 * 128u) < v0 + (uint64_t)len * vbytes; l2 += 8u)
                            prefetch_l1(reinterpret_cast<const void *>((v0 & ~(uintptr_t)127) + l2 * 128u));
                    }
                }
            }
            // short lists: one warp per list piece, round robin; few tokens -> lists are cut in 2 or 4 pieces so that
            // every warp gets about the same number of postings
            const int psh = tn >= 2 * NW ? 0 : (tn >= NW ? 1 : 2);   // 1, 2 or 4 pieces per list
            // (fixed point, index without long lists: three pieces of a warp per load round trip; with the heavy-tailed lists
            // of the LONG variant the grouping measured 7 % slower -- most of a piece is its tail there)
            if (!LONG && cross) accumulate_pieces_fixed<true, VS_K3_GROUP>(accu, p.post_row, base, s_beg, s_len, s_w, fx_scale, tn, psh, warp, lane, cross_tau, s_queue, &s_cross_n);
            else if (!LONG && fixed) accumulate_pieces_fixed<false, VS_K3_GROUP>(accu, p.post_row, base, s_beg, s_len, s_w, fx_scale, tn, psh, warp, lane, 0u, nullptr, nullptr);
            else
            for (int it = warp; it < (tn << psh); it += NW) {
                const int ti = it >> psh;
                const uint32_t pc = (uint32_t)(it & ((1 << psh) - 1));
                const uint32_t len = s_len[ti];
                if (len - 1u < kLongList) {
                    const uint32_t lo = (len * pc) >> psh, hi = (len * (pc + 1u)) >> psh;
                    const uint64_t at = base + s_beg[ti];
                    if (LONG && cross) accumulate_slice_fixed<true>(accu, p.post_row + at, lo + lane, hi, 32u, __float2uint_rn(s_w[ti] * fx_scale), cross_tau, s_queue, &s_cross_n);
                    else if (LONG && fixed) accumulate_slice_fixed<false>(accu, p.post_row + at, lo + lane, hi, 32u, __float2uint_rn(s_w[ti] * fx_scale), 0u, nullptr, nullptr);
                    else accumulate_slice<VK>(acc, p.post_row + at, (const uint8_t *)p.post_val + at * vbytes, lo + lane, hi, 32u, s_w[ti]);
                }
            }
            // long lists (heavy-tailed token popularity): every warp takes a share
            if constexpr (LONG)
            for (int tb = 0; tb < tn; tb += 32) {
                const int ti_l = tb + lane;
                uint32_t longm = __ballot_sync(0xffffffffu, ti_l < tn && s_len[ti_l] > kLongList);
                for (; longm; longm &= longm - 1) {
                    const int ti = tb + __ffs(longm) - 1;
                    const uint64_t at = base + s_beg[ti];
                    const uint16_t *lr = p.post_row + at;
                    const void *lv = (const uint8_t *)p.post_val + at * vbytes;
                    if (cross) accumulate_long<VK, 2>(acc, lr, lv, s_len[ti], s_w[ti], __float2uint_rn(s_w[ti] * fx_scale), tid, cross_tau, s_queue, &s_cross_n);
                    else if (fixed) accumulate_long<VK, 1>(acc, lr, lv, s_len[ti], s_w[ti], __float2uint_rn(s_w[ti] * fx_scale), tid, 0u, nullptr, nullptr);
                    else accumulate_long<VK, 0>(acc, lr, lv, s_len[ti], s_w[ti], 0u, tid, 0u, nullptr, nullptr);
                }
            }
        }
        __syncthreads();
        prof.lap(2);
        if ((p.flags & 2) && cached && b + 1 < blk1 && tid < cnt) {  // next block's lists -> L2 while this block is selected
            const uint64_t npos = nbase + nrng.x;
            const uint32_t nlen = min(nrng.y - nrng.x, 2048u);
            const uint8_t *r8 = reinterpret_cast<const uint8_t *>(p.post_row + npos);
            for (uint32_t o = 0; o < nlen * 2u; o += 128u) prefetch_l2(r8 + o);
            if (vbytes) {
                const uint8_t *v8 = reinterpret_cast<const uint8_t *>(p.post_val) + npos * vbytes;
                for (uint32_t o = 0; o < nlen * (uint32_t)vbytes; o += 128u) prefetch_l2(v8 + o);
            }
        }
        // ---- first block of the CTA: its first NT*4 rows are counted whole (rows at score 0 -- untouched rows, most of
        // a sparse query's -- through one warp reduction instead of per-lane atomics) and give the first pre-filter
        int count_from = 0;
        if (!booted) {
            const int r = tid * 4;
            uint32_t zeros = 0;
            if (r < rows_b) {
                float4 v = acc4[r >> 2];
                if (fixed) {
                    const uint4 u = acc4u[r >> 2];
                    v = make_float4(__uint2float_rn(u.x) * fx_inv, __uint2float_rn(u.y) * fx_inv, __uint2float_rn(u.z) * fx_inv, __uint2float_rn(u.w) * fx_inv);
                }
                const float sc[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (r + e < rows_b) {
                        const float se = round_score(sc[e], p.score_round);
                        if (se == 0.0f) ++zeros; else hist_count_plain(coarse, fine, ordered_bits(se));
                    }
            }
            zeros = __reduce_add_sync(0xffffffffu, zeros);
            if (lane == 0 && zeros) hist_count_plain(coarse, fine, ordered_bits(0.0f), zeros);
            __syncthreads();
            refresh();
            count_from = NT * 4;
            prof.lap(3);
        }
        // ---- select.  Optimistic: the whole block in one go, survivors appended through a CTA-wide counter; when the
        // append region overflows (masses of equal scores, adversarial order) the query switches to the exact machinery:
        // the block is replayed in steps of NT*4 rows with a CTA-wide re-selection whenever the next step might not fit.
        const uint32_t app0 = *(volatile uint32_t *)&st.n_app;
        const int n4 = (rows_b + 3) >> 2;   // float4s of the block's accumulator (the tail beyond rows_b is masked in take_rows)
        if (!exact) {
            const uint64_t tau = *(volatile uint64_t *)&st.tau;
            if (cross && *(volatile uint32_t *)&s_cross_n <= kCrossCap) {
                const uint32_t n_cross = *(volatile uint32_t *)&s_cross_n;   // (stable: read after the accumulate barrier)
                for (uint32_t e0 = (uint32_t)warp * 32u; e0 < n_cross; e0 += NT) take_crossed(e0 + lane < n_cross, s_queue[min(e0 + lane, n_cross - 1u)], rows_b, row0, tau);
            } else {
                // One pass over the block's sums.  The CTA's first block has no pre-filter worth the name yet: it is cut into
                // ranges (NT float4s, then up to 3 NT, 7 NT, ...) and, for large k, the pre-filter is raised from the histogram
                // before every range, dropping what fell below the new bound whenever the region is half full, to stay out of
                // the exact fallback.  Small k (<= 128): the bound of the first 3,072 rows already leaves few enough survivors
                // (~1,400 of a 37 K-row block) for the append region, and every refresh costs two barriers and a serial
                // histogram walk -- measured: no refresh inside the block = 31.4 -> 30.7 us per query, shard 6.74 -> 6.43.
                const bool more = p.k > 128;
                int i_done = 0, i_next = booted ? n4 : NT;
#pragma unroll 1
                while (i_done < n4) {
                    const int i_hi = min(n4, i_next);
                    if (!booted && i_done > 0 && more) {
                        __syncthreads();
                        const uint32_t ob = refresh();
                        // (a count beyond the region is an overflow: left alone, the check after the block sees it)
                        if (n_app_seen > (uint32_t)(kInvAppend / 2) && n_app_seen <= (uint32_t)kInvAppend && ob)
                            compact_appended(app, &st, n_app_seen, ob);
                        prof.lap(9);
                    }
                    select_range(i_done, i_hi, rows_b, row0, tau, count_from);
                    if (!booted) prof.lap(10);
                    i_done = i_hi;
                    i_next = (!booted && (more || i_next < 3 * NT)) ? 2 * i_next + NT : n4;
                }
            }
            flush_cta(rows_b, row0, tau, count_from);
            __syncthreads();
            if (*(volatile uint32_t *)&st.n_app > (uint32_t)kInvAppend) {   // overflow: drop this block's appends, go exact
                __syncthreads();
                if (tid == 0) { st.n_app = app0 <= (uint32_t)kInvAppend ? app0 : 0u; st.cnt = 0; }
                exact = true;   // from here on the histogram memory holds the shared top-k set
                __syncthreads();
            }
        }
        if (exact) {   // replay / later blocks of an exact-mode query: steps of NT*4 rows, re-selections whenever a step might not fit
            uint64_t tau = *(volatile uint64_t *)&st.tau;
            for (int i0 = 0; i0 < n4; i0 += NT) {
                // the thread that appended last reads the final count, so the OR is exact
                if (__syncthreads_or(*(volatile uint32_t *)&st.n_app + NT * 4 > (uint32_t)kInvAppend)) {
                    cta_join_flat<NT, kInvAppend>(shset, app, p.k, hist, &st);
                    tau_s = fmaxf(tau_s, gate_tau_score(gate_load(&st)));
                    tau = *(volatile uint64_t *)&st.tau;
                }
                select_range(i0, min(n4, i0 + NT), rows_b, row0, tau, count_from);
                flush_own(rows_b, row0, tau, count_from);
            }
            __syncthreads();
        }
        booted = true;
        prof.lap(4);
        // ---- raise the pre-filter for the next block; when the region is half full, drop what fell below it
        if (!exact && b + 1 < blk1) {
            const uint32_t ob = refresh();
            if (n_app_seen > (uint32_t)(kInvAppend / 2)) {
                compact_appended(app, &st, n_app_seen, ob);
                if (*(volatile uint32_t *)&st.n_app > (uint32_t)(kInvAppend / 2)) {   // a bucket full of equal scores: go exact
                    if (tid == 0) st.cnt = 0;
                    exact = true;
                    __syncthreads();
                    cta_join_flat<NT, kInvAppend>(shset, app, p.k, hist, &st);
                    tau_s = fmaxf(tau_s, gate_tau_score(gate_load(&st)));
                }
            }
        } else if (exact && b + 1 < blk1 && *(volatile uint32_t *)&st.n_app > (uint32_t)(kInvAppend / 4)) {
            cta_join_flat<NT, kInvAppend>(shset, app, p.k, hist, &st);
            tau_s = fmaxf(tau_s, gate_tau_score(gate_load(&st)));
        }
        __syncthreads();
        prof.lap(5);
    }
    // ---- end of query: everything at or above the final histogram bound goes to the merge kernel (k .. kout keys);
    // more than kout (a bucket full of equal scores), or a query in exact mode -> the exact CTA-wide select
    {
        uint64_t *out = p.cand + ((size_t)q * p.n_lists + blockIdx.x) * (size_t)p.cand_stride;
        for (int l = blockIdx.x + gridDim.x; l < p.n_lists; l += gridDim.x) {   // lists of CTAs that were not launched
            uint64_t *z = p.cand + ((size_t)q * p.n_lists + l) * (size_t)p.cand_stride;
            for (int i = tid; i < p.kout; i += NT) z[i] = 0ull;
        }
        if (tid == 0) st.scratch = 0;
        __syncthreads();
        bool use_exact = exact;
        if (!exact) {
            if (warp == 0) {
                const uint32_t ob = hist_threshold(coarse, fine, p.k);
                if (lane == 0) s_ob = ob;
            }
            __syncthreads();
            const uint64_t tmin = (uint64_t)s_ob << 32;
            const int n_app = (int)min(*(volatile uint32_t *)&st.n_app, (uint32_t)kInvAppend);
            for (int i = tid; i < n_app; i += NT) {
                const uint64_t x = app[i];
                if (x != 0ull && x >= tmin) { const uint32_t o = atomicAdd(&st.scratch, 1u); if (o < (uint32_t)p.kout) out[o] = x; }
            }
            __syncthreads();
            const uint32_t total = *(volatile uint32_t *)&st.scratch;
            if (total > (uint32_t)p.kout) {   // uniform: too many ties for the list -> exact select of what was kept
                if (tid == 0) st.cnt = 0;
                use_exact = true;
                __syncthreads();
            } else {
                for (int i = (int)total + tid; i < p.kout; i += NT) out[i] = 0ull;
            }
        }
        if (use_exact) {
            cta_write_topk_flat<NT, kInvAppend>(shset, app, p.k, hist, &st, out);
            for (int i = p.k + tid; i < p.kout; i += NT) out[i] = 0ull;
        }
    }
    if (PROF && prof.on) {
        prof.lap(6);
        prof.acc[7] = prof.t - prof.acc[7];
        for (int i = 0; i < 16; ++i) p.prof[((size_t)q * p.n_lists + blockIdx.x) * 16 + i] = prof.acc[i];
    }
    __syncthreads();   // the next query resets the CTA state and the histogram

    }   // queries
}

// ------------------------------------------------------------------------------------------- host side
size_t inverted_workspace_bytes(const vs_index *idx, int64_t Bc, int group) {
    (void)idx; (void)group;
    size_t lists = (size_t)Bc * ((size_t)kMaxQueryNnz * 8 + ((size_t)kMaxQueryNnz + 1) * 4 + 4 + 4 + 8);
    return (lists + 255) / 256 * 256 + 1024;
}

static QueryLists carve_lists(uint8_t *base, int64_t Bc, uint8_t **end) {
    QueryLists L;
    size_t o = 0;
    L.total = (uint64_t *)(base + o); o += (size_t)Bc * 8;
    L.tok = (uint32_t *)(base + o); o += (size_t)Bc * kMaxQueryNnz * 4;
    L.w = (float *)(base + o); o += (size_t)Bc * kMaxQueryNnz * 4;
    L.pref = (uint32_t *)(base + o); o += (size_t)Bc * (kMaxQueryNnz + 1) * 4;
    L.cnt = (uint32_t *)(base + o); o += (size_t)Bc * 4;
    L.fx = (int *)(base + o); o += (size_t)Bc * 4;
    *end = base + (o + 255) / 256 * 256;
    return L;
}

int scan_cap_for_k(int k);

// Sparse form of Bc queries -> QueryLists in d_ws, from the prepared dense rows (extract) or from caller-supplied
// (token, weight) lists; then the scan-vs-inverted decision, all on the stream (no readback).  d_flag: device int the
// scoring kernels and the merge read; idx->d_last_mode keeps the decision for vs_index_last_mode.
int inverted_prepare(vs_index *idx, const float *d_qprep, int vpad, const void *d_qptr, int ptr_dtype, const int32_t *d_qtok,
                     const float *d_qw, int64_t b0, int64_t Bc, int mode, int round_mode, double t_scan, double postings_per_sec,
                     double t_rows_fixed, void *d_ws, int *d_flag, cudaStream_t st) {
    int rc = build_inverted(idx, st);   // lazy, first use only (SYNC once)
    if (rc) return rc;
    uint8_t *end;
    QueryLists L = carve_lists((uint8_t *)d_ws, Bc, &end);
    if (d_qptr != nullptr)
        inv_lists_from_csr_kernel<<<(unsigned)Bc, 256, 0, st>>>(d_qptr, ptr_dtype, d_qtok, d_qw, b0, (int)idx->n_cols, round_mode,
                                                                idx->post_ptr, L);
    else
        inv_extract_kernel<<<(unsigned)Bc, 256, 0, st>>>(d_qprep, vpad, (int)idx->n_cols, idx->post_ptr, L);
    VS_CUDA(cudaGetLastError());
    if (idx->kind != 1) inv_fixed_point_kernel<<<(unsigned)((Bc + 7) / 8), 256, 0, st>>>(L, (int)Bc);   // binary index
    inv_decide_kernel<<<1, 256, 0, st>>>(L, (int)Bc, mode, t_scan, postings_per_sec, t_rows_fixed, d_flag, idx->d_last_mode);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

// Score Bc extracted queries (lists already in d_ws) -> cand [Bc, n_ctas, k]; a no-op when *d_flag == 0.
int launch_inverted(vs_index *idx, int64_t Bc, int k, int cand_stride, int score_round, void *d_ws, uint64_t *d_cand,
                    const int *d_flag, cudaStream_t st) {
    uint8_t *end;
    QueryLists L = carve_lists((uint8_t *)d_ws, Bc, &end);
    static_assert(kHistFine * 4 >= kSharedKeys * 8, "the exact fallback's top-k set lives in the histogram's memory");
    const size_t smem = (size_t)(kHistFine + kHistCoarse + 256) * 4 + (size_t)kInvAppend * 8 + (size_t)kTokTile * 4 * 4 +
                        (size_t)kInvWarps * kInvQueue * 2 + (size_t)idx->blk_rows * 4;
    VS_REQUIRE(smem <= 227 * 1024, VS_ERR_UNSUPPORTED, "inverted-list search needs %zu bytes of shared memory", smem);
    InvSearchParams p;
    p.L = L; p.blk_rng = (const uint2 *)idx->blk_ptr; p.blk_base = idx->blk_base; p.post_row = idx->post_row; p.post_val = idx->post_val;
    p.val_kind = idx->kind == 1 ? (idx->store_dtype == VS_F32 ? 1 : (idx->store_dtype == VS_F16 ? 2 : 3)) : 0;
    p.V = (int)idx->n_cols; p.rows_per_block = idx->blk_rows; p.n_blocks = idx->n_blocks; p.blocks_per_cta = idx->blocks_per_cta;
    p.n_rows = idx->n_rows; p.cand = d_cand; p.k = k; p.kout = cand_stride; p.score_round = score_round; p.cand_stride = cand_stride;
    static const int k3_flags = getenv("VSEARCH_B200_K3_FLAGS") ? atoi(getenv("VSEARCH_B200_K3_FLAGS")) : 3;
    p.flags = k3_flags;
    p.use_inv = d_flag;
    p.prof = idx->scan_prof;
    p.n_lists = idx->n_ctas;
    p.n_queries = (int)Bc;
    // a CTA with 227 KB of shared memory costs microseconds to launch even when it owns no row block (a shard of 71
    // blocks on 148 SMs spent half its time there): launch only the CTAs that have blocks
    const unsigned grid_x = (unsigned)((idx->n_blocks + idx->blocks_per_cta - 1) / idx->blocks_per_cta);
    auto launch = [&](auto kern) -> int {
        VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // grid: every CTA that owns row blocks, times as many query lanes as fit the SMs in ONE wave
        static const int k3_waves = getenv("VSEARCH_B200_K3_GRIDY") ? atoi(getenv("VSEARCH_B200_K3_GRIDY")) : 0;
        unsigned grid_y = k3_waves > 0 ? (unsigned)k3_waves : (unsigned)(idx->n_ctas / grid_x);
        if (grid_y < 1) grid_y = 1;
        if (grid_y > (unsigned)Bc) grid_y = (unsigned)Bc;
        kern<<<dim3(grid_x, grid_y), kInvThreads, smem, st>>>(p);
        VS_CUDA(cudaGetLastError());
        return VS_OK;
    };
    const bool rnd = score_round != VS_F32;
    if (p.prof != nullptr && p.val_kind == 0 && !rnd)
        return idx->inv_has_long ? launch(inv_search_kernel<0, false, true, true>) : launch(inv_search_kernel<0, false, false, true>);
    if (p.prof != nullptr && p.val_kind == 1 && !rnd && !idx->inv_has_long) return launch(inv_search_kernel<1, false, false, true>);
    if (idx->inv_has_long)
        switch (p.val_kind) {
            case 0: return rnd ? launch(inv_search_kernel<0, true, true>) : launch(inv_search_kernel<0, false, true>);
            case 1: return rnd ? launch(inv_search_kernel<1, true, true>) : launch(inv_search_kernel<1, false, true>);
            case 2: return rnd ? launch(inv_search_kernel<2, true, true>) : launch(inv_search_kernel<2, false, true>);
            default: return rnd ? launch(inv_search_kernel<3, true, true>) : launch(inv_search_kernel<3, false, true>);
        }
    switch (p.val_kind) {
        case 0: return rnd ? launch(inv_search_kernel<0, true, false>) : launch(inv_search_kernel<0, false, false>);
        case 1: return rnd ? launch(inv_search_kernel<1, true, false>) : launch(inv_search_kernel<1, false, false>);
        case 2: return rnd ? launch(inv_search_kernel<2, true, false>) : launch(inv_search_kernel<2, false, false>);
        default: return rnd ? launch(inv_search_kernel<3, true, false>) : launch(inv_search_kernel<3, false, false>);
    }
}

}  // namespace vs
