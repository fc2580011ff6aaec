// inverted.cu -- K3: token-major inverted-list scoring for sparse queries, sm_100a.
// Same contract as the scan (replaces upstream index.py:91-92) but touches only the posting lists of
// the query's non-zero tokens:
//   build   : WS stream -> post_ptr[V+1], post_doc[nnz] (+ post_val) grouped by token (one-off, lazy)
//   extract : prepared query [vpad] -> compact (token, weight) list + prefix of posting-list lengths
//   accum   : flattened postings of the query -> RED.ADD.F32 into a zeroed fp32 accumulator row [N]
//             (L2-resident: 21 M x 4 B = 84 MB)
//   select  : stream the accumulator once, zero it behind, fused top-k (topk.cuh) -> per-CTA key lists
//             -> merge.cu.  Rows never touched keep score 0 and compete like any other row.
// Algorithmic bytes per query: sum_t len(post_t) * (4 + b_val) + 2 * N * 4.
#include <cub/device/device_scan.cuh>

#include <stdlib.h>

#include <vector>

#include "index.cuh"
#include "topk.cuh"

namespace vs {

constexpr int kInvThreads = 1024;
constexpr int kBuildThreads = kScanWarps * 32;   // the builders walk the stream with the scan kernel's partition
constexpr int kMaxQueryNnz = 4096;   // queries denser than this are served by the scan kernels

// ------------------------------------------------------------------------------------------- build
// Both build kernels walk the WS stream exactly like the scan kernel: warp = part, window by window.
template <bool FILL>
__global__ void __launch_bounds__(kBuildThreads, 1)
inv_build_kernel(const WsView idx, uint32_t *cta_hist /* [n_ctas, V] counts (count pass) / offsets (fill pass) */) {
    extern __shared__ uint32_t s_cnt[];  // V counters
    const int V = (int)idx.n_cols;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *mine = cta_hist + (size_t)blockIdx.x * V;
    for (int i = tid; i < V; i += kBuildThreads) s_cnt[i] = FILL ? mine[i] : 0u;
    __syncthreads();
    const int part = blockIdx.x * kScanWarps + warp;
    const uint32_t w_begin = idx.part_win_begin[part];
    const int nwin = (int)(idx.part_win_begin[part + 1] - w_begin);
    uint32_t row = idx.part_row_begin[part];
    const uint32_t lt = lanemask_lt();
    for (int w = 0; w < nwin; ++w) {
        const uint64_t chunk = ((uint64_t)w_begin + w) * 32ull + lane;
        const uint32_t T = idx.tails[w_begin + w];
        const uint32_t rid = row + __popc(T & lt);
        row += __popc(T);
        const uint4 u = idx.cols[chunk];
        const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t c = ((e & 1) ? (wv[e >> 1] >> 16) : (wv[e >> 1] & 0xffffu)) & 0x7fffu;  // bit 15 = tail flag
            if (c >= (uint32_t)V) continue;  // padding
            const uint32_t slot = atomicAdd(&s_cnt[c], 1u);
            if constexpr (FILL) {
                const uint64_t pos = idx.post_ptr[c] + slot;
                idx.post_doc[pos] = rid;
                if (idx.kind == 1) {
                    const uint64_t src = chunk * 8ull + e;
                    if (idx.store_dtype == VS_F32) ((float *)idx.post_val)[pos] = ((const float *)idx.vals)[src];
                    else ((uint16_t *)idx.post_val)[pos] = ((const uint16_t *)idx.vals)[src];
                }
            }
        }
    }
    if constexpr (!FILL) {
        __syncthreads();
        for (int i = tid; i < V; i += kBuildThreads) mine[i] = s_cnt[i];
    }
}

// per token: counts per CTA -> exclusive offsets per CTA (in place) and the token total
__global__ void inv_offsets_kernel(uint32_t *cta_hist, int n_ctas, int V, uint64_t *totals, int *err) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > V) return;
    if (t == V) { totals[t] = 0; return; }
    uint64_t run = 0;
    for (int c = 0; c < n_ctas; ++c) {
        uint32_t h = cta_hist[(size_t)c * V + t];
        cta_hist[(size_t)c * V + t] = (uint32_t)run;
        run += h;
    }
    if (run >= (1ull << 32)) atomicExch(err, 1);
    totals[t] = run;
}

int build_inverted(vs_index *idx, cudaStream_t st) {
    if (idx->inv_built) return VS_OK;
    const int V = (int)idx->n_cols;
    const size_t smem = (size_t)V * 4;
    VS_REQUIRE(smem <= 227 * 1024, VS_ERR_UNSUPPORTED, "vocabulary too large for the inverted-list builder");
    uint32_t *d_hist = nullptr;
    int *d_err = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    auto cleanup = [&]() { cudaFree(d_hist); cudaFree(d_err); cudaFree(d_tmp); };
    VS_CUDA(cudaMalloc(&d_hist, (size_t)idx->n_ctas * V * 4));
    VS_CUDA(cudaMalloc(&d_err, 4));
    VS_CUDA(cudaMemsetAsync(d_err, 0, 4, st));
    VS_CUDA(cudaMalloc(&idx->post_ptr, (size_t)(V + 1) * 8));
    VS_CUDA(cudaMalloc(&idx->post_doc, idx->nnz ? (size_t)idx->nnz * 4 : 4));
    size_t vbytes = idx->kind == 1 ? (idx->store_dtype == VS_F32 ? 4 : 2) : 0;
    if (vbytes) VS_CUDA(cudaMalloc(&idx->post_val, idx->nnz ? (size_t)idx->nnz * vbytes : 4));

    VS_CUDA(cudaFuncSetAttribute(inv_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_CUDA(cudaFuncSetAttribute(inv_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    inv_build_kernel<false><<<idx->n_ctas, kBuildThreads, smem, st>>>(ws_view(idx), d_hist);
    inv_offsets_kernel<<<(V + 1 + 255) / 256, 256, 0, st>>>(d_hist, idx->n_ctas, V, idx->post_ptr, d_err);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, idx->post_ptr, idx->post_ptr, V + 1, st);
    VS_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, idx->post_ptr, idx->post_ptr, V + 1, st);
    inv_build_kernel<true><<<idx->n_ctas, kBuildThreads, smem, st>>>(ws_view(idx), d_hist);
    int h_err = 0;
    cudaError_t e = cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    VS_CUDA(e);
    VS_REQUIRE(h_err == 0, VS_ERR_UNSUPPORTED, "a posting list exceeds 2^32 entries");
    idx->inv_bytes = (int64_t)((size_t)idx->nnz * (4 + vbytes) + (size_t)(V + 1) * 8);
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
};

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

// ------------------------------------------------------------------------------------------- accumulate
struct AccumParams {
    QueryLists L;
    const uint64_t *post_ptr;
    const uint32_t *post_doc;
    const void *post_val;
    int val_kind;       // 0 none (binary), 1 f32, 2 f16, 3 bf16
    float *acc;         // [G, n_pad]
    int64_t n_pad;
    int b0;             // first query of the group
};

__global__ void __launch_bounds__(256) inv_accum_kernel(const AccumParams p) {
    extern __shared__ __align__(16) uint8_t asmem[];
    const int g = blockIdx.y, b = p.b0 + g, tid = threadIdx.x;
    const uint32_t cnt = min(p.L.cnt[b], (uint32_t)kMaxQueryNnz);
    uint32_t *s_pref = reinterpret_cast<uint32_t *>(asmem);                    // cnt + 1
    uint64_t *s_base = reinterpret_cast<uint64_t *>(asmem + (((size_t)cnt + 1) * 4 + 15) / 16 * 16);  // cnt
    float *s_w = reinterpret_cast<float *>(s_base + cnt);                     // cnt
    const uint32_t *pref = p.L.pref + (size_t)b * (kMaxQueryNnz + 1);
    const uint32_t *tok = p.L.tok + (size_t)b * kMaxQueryNnz;
    const float *w = p.L.w + (size_t)b * kMaxQueryNnz;
    for (uint32_t i = tid; i <= cnt; i += 256) s_pref[i] = pref[i];
    for (uint32_t i = tid; i < cnt; i += 256) { s_base[i] = p.post_ptr[tok[i]]; s_w[i] = w[i]; }
    __syncthreads();
    const uint32_t total = s_pref[cnt];
    float *acc = p.acc + (size_t)g * p.n_pad;
    const uint32_t stride = gridDim.x * 256u;
    // 4 independent postings per thread per trip: the id loads overlap, then the reductions are issued back to back
    for (uint64_t idx64 = (uint64_t)blockIdx.x * 256u + tid; idx64 < total; idx64 += 4ull * stride) {
        uint32_t doc[4];
        float v[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t i64 = idx64 + (uint64_t)u * stride;
            ok[u] = i64 < total;
            const uint32_t idx = ok[u] ? (uint32_t)i64 : 0u;
            uint32_t lo = 0, hi = cnt;  // largest lo with s_pref[lo] <= idx
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if (s_pref[mid] <= idx) lo = mid; else hi = mid;
            }
            const uint64_t pos = s_base[lo] + (idx - s_pref[lo]);
            doc[u] = ok[u] ? p.post_doc[pos] : 0u;
            v[u] = s_w[lo];
            if (ok[u]) {
                if (p.val_kind == 1) v[u] *= ((const float *)p.post_val)[pos];
                else if (p.val_kind == 2) v[u] *= __half2float(((const __half *)p.post_val)[pos]);
                else if (p.val_kind == 3) v[u] *= __bfloat162float(((const __nv_bfloat16 *)p.post_val)[pos]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (ok[u]) atomicAdd(acc + doc[u], v[u]);  // result unused -> RED.E.ADD.F32
    }
}

// ------------------------------------------------------------------------------------------- select
struct SelectParams {
    float *acc;          // [G, n_pad]; zeroed behind the read
    uint64_t *cand;      // [B, n_ctas, k]
    int64_t n_rows, n_pad;
    int b0, k, cap, score_round;
    int rows_per_cta;    // multiple of 4
    int zero_behind;     // 1 (always, except timing experiments): clear the accumulator row behind the read
};

__global__ void __launch_bounds__(kInvThreads, 1) inv_select_kernel(const SelectParams p) {
    extern __shared__ __align__(128) uint8_t ssmem[];
    uint64_t *cbuf = reinterpret_cast<uint64_t *>(ssmem);
    uint32_t *hist = reinterpret_cast<uint32_t *>(cbuf + kCapMax);
    __shared__ CtaState st;
    constexpr int NW = kInvThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31;
    const int g = blockIdx.y, b = p.b0 + g;
    const uint32_t lt = lanemask_lt();
    if (tid == 0) cta_state_reset(&st);
    __syncthreads();
    float4 *acc4 = reinterpret_cast<float4 *>(p.acc + (size_t)g * p.n_pad);
    const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_cta;
    const int64_t r_end = min(p.n_rows, r_begin + p.rows_per_cta);
    // ---- phase A: the first kCapMax rows go straight into cbuf, then one CTA-wide select sets the threshold
    for (int i = tid; i < (kCapMax >> 2); i += kInvThreads) {
        const int64_t r = r_begin + (int64_t)i * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < r_end) {
            v = acc4[r >> 2];
            acc4[r >> 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float s[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
            cbuf[i * 4 + e] = (r + e < r_end) ? make_key(round_score(s[e], p.score_round), (uint32_t)(r + e)) : 0ull;
    }
    __syncthreads();
    cta_sample_select<kInvThreads>(cbuf, kCapMax, p.k, hist, &st);
    int n_priv = 0;
    uint32_t epoch = 0;
    // ---- phase B: warp-uniform trip count (every lane of a warp iterates the same number of times).
    // kSelU float4 loads per thread are issued before any of the zeroing stores: interleaving a load and a store
    // to the same line per trip runs 3-5x slower (scripts/micro/red_stream.cu, measured on B200).
    constexpr int kSelU = 4;
    static_assert(TopkGeom<NW>::kPrivate >= 128 + 32, "private region must take one 128-row step");
    for (int64_t r = r_begin + kCapMax + (int64_t)tid * 4; r - (int64_t)lane * 4 < r_end;
         r += (int64_t)kInvThreads * 4 * kSelU) {
        float4 v[kSelU];
        bool in[kSelU];
#pragma unroll
        for (int u = 0; u < kSelU; ++u) {
            const int64_t ru = r + (int64_t)u * kInvThreads * 4;
            in[u] = ru < r_end;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in[u]) v[u] = acc4[ru >> 2];
        }
#pragma unroll
        for (int u = 0; u < kSelU; ++u) {
            const int64_t ru = r + (int64_t)u * kInvThreads * 4;
            if (in[u] && p.zero_behind) acc4[ru >> 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kSelU; ++u) {
            const int64_t ru = r + (int64_t)u * kInvThreads * 4;
            const float s[4] = {round_score(v[u].x, p.score_round), round_score(v[u].y, p.score_round),
                                round_score(v[u].z, p.score_round), round_score(v[u].w, p.score_round)};
            const uint64_t gate = gate_load(&st);
            const float tau_s = gate_tau_score(gate);
            const float mx = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
            if (__any_sync(0xffffffffu, in[u] && mx >= tau_s)) {  // else nothing in these 128 rows can qualify
                const uint64_t tau = *(volatile uint64_t *)&st.tau;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t rid = ru + e;
                    const uint64_t key = make_key(s[e], (uint32_t)rid);
                    private_insert<NW>(in[u] && rid < r_end && key > tau, key, cbuf, n_priv, lt);
                }
            }
            join_if_needed<kInvThreads, NW>(gate, epoch, 128, cbuf, n_priv, p.k, hist, &st);
        }
    }
    finish_streaming<kInvThreads, NW>(epoch, cbuf, n_priv, p.k, hist, &st);
    cta_write_topk<kInvThreads, NW>(cbuf, n_priv, p.k, hist, &st,
                                    p.cand + ((size_t)b * gridDim.x + blockIdx.x) * (size_t)p.k);
}

// ------------------------------------------------------------------------------------------- host side
size_t inverted_workspace_bytes(const vs_index *idx, int64_t Bc, int group) {
    size_t n_pad = ((size_t)idx->n_rows + 3) / 4 * 4;
    size_t lists = (size_t)Bc * ((size_t)kMaxQueryNnz * 8 + ((size_t)kMaxQueryNnz + 1) * 4 + 4 + 8);
    return (lists + 255) / 256 * 256 + ((size_t)group * n_pad * 4 + 255) / 256 * 256 + 1024;
}

static QueryLists carve_lists(uint8_t *base, int64_t Bc, uint8_t **end) {
    QueryLists L;
    size_t o = 0;
    L.total = (uint64_t *)(base + o); o += (size_t)Bc * 8;
    L.tok = (uint32_t *)(base + o); o += (size_t)Bc * kMaxQueryNnz * 4;
    L.w = (float *)(base + o); o += (size_t)Bc * kMaxQueryNnz * 4;
    L.pref = (uint32_t *)(base + o); o += (size_t)Bc * (kMaxQueryNnz + 1) * 4;
    L.cnt = (uint32_t *)(base + o); o += (size_t)Bc * 4;
    *end = base + (o + 255) / 256 * 256;
    return L;
}

int scan_cap_for_k(int k);

// Extract the sparse form of Bc prepared queries and report (SYNC: a 12*Bc-byte readback) the largest
// non-zero count and the mean postings per query, so the caller can choose scan vs inverted.
int inverted_extract(vs_index *idx, const float *d_qprep, int vpad, int64_t Bc, void *d_ws, uint32_t *max_nnz,
                     double *mean_postings, uint64_t *max_postings, cudaStream_t st) {
    int rc = build_inverted(idx, st);
    if (rc) return rc;
    uint8_t *end;
    QueryLists L = carve_lists((uint8_t *)d_ws, Bc, &end);
    inv_extract_kernel<<<(unsigned)Bc, 256, 0, st>>>(d_qprep, vpad, (int)idx->n_cols, idx->post_ptr, L);
    VS_CUDA(cudaGetLastError());
    std::vector<uint32_t> h_cnt((size_t)Bc);
    std::vector<uint64_t> h_tot((size_t)Bc);
    VS_CUDA(cudaMemcpyAsync(h_cnt.data(), L.cnt, (size_t)Bc * 4, cudaMemcpyDeviceToHost, st));
    VS_CUDA(cudaMemcpyAsync(h_tot.data(), L.total, (size_t)Bc * 8, cudaMemcpyDeviceToHost, st));
    VS_CUDA(cudaStreamSynchronize(st));
    uint32_t mx = 0;
    uint64_t mp = 0;
    double sum = 0;
    for (int64_t i = 0; i < Bc; ++i) {
        mx = h_cnt[i] > mx ? h_cnt[i] : mx;
        mp = h_tot[i] > mp ? h_tot[i] : mp;
        sum += (double)h_tot[i];
    }
    *max_nnz = mx;
    *max_postings = mp;
    *mean_postings = Bc ? sum / (double)Bc : 0.0;
    return VS_OK;
}

bool inverted_usable(uint32_t max_nnz, uint64_t max_postings) {
    return max_nnz <= (uint32_t)kMaxQueryNnz && max_postings < (1ull << 32);
}

// Score Bc extracted queries (lists already in d_ws) -> cand [Bc, n_ctas, k].
int launch_inverted(vs_index *idx, int64_t Bc, int k, int score_round, int group, uint32_t max_nnz, void *d_ws, uint64_t *d_cand,
                    cudaEvent_t ev0, cudaEvent_t ev1, cudaStream_t st) {
    uint8_t *acc_base;
    QueryLists L = carve_lists((uint8_t *)d_ws, Bc, &acc_base);
    const int64_t n_pad = (idx->n_rows + 3) / 4 * 4;
    float *acc = (float *)acc_base;
    VS_CUDA(cudaMemsetAsync(acc, 0, (size_t)group * n_pad * 4, st));
    const int cap = scan_cap_for_k(k);
    const size_t sel_smem = (size_t)kCapMax * 8 + 256 * 4;
    VS_CUDA(cudaFuncSetAttribute(inv_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    const size_t acc_smem = ((size_t)max_nnz + 1) * 4 + 16 + (size_t)max_nnz * 12 + 16;
    VS_CUDA(cudaFuncSetAttribute(inv_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_smem));
    int rows_per_cta = (int)((idx->n_rows + idx->n_ctas - 1) / idx->n_ctas);
    rows_per_cta = (rows_per_cta + 3) / 4 * 4;
    if (ev0) VS_CUDA(cudaEventRecord(ev0, st));
    for (int64_t b0 = 0; b0 < Bc; b0 += group) {
        const int G = (int)((Bc - b0) < group ? (Bc - b0) : group);
        AccumParams ap;
        ap.L = L; ap.post_ptr = idx->post_ptr; ap.post_doc = idx->post_doc; ap.post_val = idx->post_val;
        ap.val_kind = idx->kind == 1 ? (idx->store_dtype == VS_F32 ? 1 : (idx->store_dtype == VS_F16 ? 2 : 3)) : 0;
        ap.acc = acc; ap.n_pad = n_pad; ap.b0 = (int)b0;
        inv_accum_kernel<<<dim3(idx->n_ctas * 4, G), 256, acc_smem, st>>>(ap);
        SelectParams sp;
        sp.acc = acc; sp.cand = d_cand; sp.n_rows = idx->n_rows; sp.n_pad = n_pad; sp.b0 = (int)b0; sp.k = k;
        sp.cap = cap; sp.score_round = score_round; sp.rows_per_cta = rows_per_cta;
        sp.zero_behind = getenv("VSEARCH_B200_DEBUG_NOZERO") ? 0 : 1;
        inv_select_kernel<<<dim3(idx->n_ctas, G), kInvThreads, sel_smem, st>>>(sp);
    }
    if (ev1) VS_CUDA(cudaEventRecord(ev1, st));
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vs
