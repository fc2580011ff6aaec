// build_index.cu -- CSR triple -> warp-stream (WS) device format, and back.
// Replaces SparseIndex._scipy_csr_to_torch_csr + .to(device) (upstream index.py:144-161,179).
// One-off work at index load; not on the search path (CUB's scan is used for the prefix sum).
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

#include "index.cuh"

namespace vs {

template <typename T>
__device__ __forceinline__ int64_t load_idx(const void *p, int64_t i) { return (int64_t)((const T *)p)[i]; }
__device__ __forceinline__ int64_t load_index(const void *p, int dtype, int64_t i) {
    return dtype == VS_I32 ? load_idx<int32_t>(p, i) : load_idx<int64_t>(p, i);
}
__device__ __forceinline__ float load_value(const void *p, int dtype, int64_t i) {
    if (dtype == VS_F32) return ((const float *)p)[i];
    if (dtype == VS_F16) return __half2float(((const __half *)p)[i]);
    return __bfloat162float(((const __nv_bfloat16 *)p)[i]);
}

// chunks per row (>= 1 so that every row, even an empty one, owns a tail bit)
__global__ void row_chunks_kernel(const void *crow, int crow_dtype, int64_t n_rows, uint64_t *rc, int *err) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    if (r == n_rows) { rc[r] = 0; return; }
    int64_t a = load_index(crow, crow_dtype, r), e = load_index(crow, crow_dtype, r + 1);
    if (e < a) { atomicExch(err, 1); e = a; }
    uint64_t len = (uint64_t)(e - a);
    rc[r] = len == 0 ? 1 : (len + 7) / 8;
}

// part p starts at the first row whose chunk offset is >= p/n_parts of the total
__global__ void part_rows_kernel(const uint64_t *cptr, int64_t n_rows, int n_parts, uint32_t *part_row_begin) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_parts) return;
    if (p == n_parts) { part_row_begin[p] = (uint32_t)n_rows; return; }
    uint64_t total = cptr[n_rows];
    uint64_t target = (total / (uint64_t)n_parts) * (uint64_t)p + (total % (uint64_t)n_parts) * (uint64_t)p / (uint64_t)n_parts;
    int64_t lo = 0, hi = n_rows;  // first r in [0, N] with cptr[r] >= target
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (cptr[mid] >= target) hi = mid; else lo = mid + 1;
    }
    part_row_begin[p] = (uint32_t)lo;
}

__global__ void part_windows_kernel(const uint64_t *cptr, const uint32_t *part_row_begin, int n_parts,
                                    uint32_t *part_win_begin, uint64_t *n_windows_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint64_t w = 0;
    for (int p = 0; p < n_parts; ++p) {
        part_win_begin[p] = (uint32_t)w;
        uint64_t chunks = cptr[part_row_begin[p + 1]] - cptr[part_row_begin[p]];
        w += (chunks + 63) / 64 * 2;   // whole 64-chunk steps: the scan kernel takes two chunks per lane per step
    }
    part_win_begin[n_parts] = (uint32_t)w;
    *n_windows_out = w;
}

__global__ void fill_u32_kernel(uint32_t *p, uint64_t n, uint32_t v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// one warp per row: scatter the row's entries into its chunks, set the tail bit, record row_chunk
template <typename VT>
__global__ void fill_rows_kernel(const void *crow, int crow_dtype, const void *col, int col_dtype, const void *val,
                                 int val_dtype, int64_t n_rows, int64_t n_cols, const uint64_t *cptr,
                                 const uint32_t *part_row_begin, const uint32_t *part_win_begin, int n_parts,
                                 uint16_t *cols16, VT *vals, uint32_t *tails, uint32_t *row_chunk, int *err) {
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    // last part p with part_row_begin[p] <= r  (empty parts repeat the same begin; the last one owns r)
    int lo = 0, hi = n_parts;  // invariant: part_row_begin[lo] <= r
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if ((int64_t)part_row_begin[mid] <= r) lo = mid; else hi = mid - 1;
    }
    int p = lo;
    uint64_t dst = (uint64_t)part_win_begin[p] * 32ull + (cptr[r] - cptr[part_row_begin[p]]);
    uint64_t nchunks = cptr[r + 1] - cptr[r];
    int64_t a = load_index(crow, crow_dtype, r), e = load_index(crow, crow_dtype, r + 1);
    if (e < a) e = a;
    for (int64_t j = a + lane; j < e; j += 32) {
        int64_t c = load_index(col, col_dtype, j);
        if (c < 0 || c >= n_cols) { atomicExch(err, 2); continue; }  // leaves the sentinel in place
        uint64_t o = dst * 8ull + (uint64_t)(j - a);
        cols16[o] = (uint16_t)c;
        if constexpr (sizeof(VT) == 4) {
            vals[o] = load_value(val, val_dtype, j);
        } else if constexpr (sizeof(VT) == 2) {
            float v = load_value(val, val_dtype, j);
            if constexpr (std::is_same<VT, __half>::value) vals[o] = __float2half_rn(v);
            else vals[o] = __float2bfloat16_rn(v);
        }
    }
    if (lane == 0) {
        uint64_t t = dst + nchunks - 1;
        atomicOr(&tails[t >> 5], 1u << (uint32_t)(t & 31));
        row_chunk[r] = (uint32_t)dst;
        if (r == n_rows - 1) row_chunk[n_rows] = (uint32_t)(dst + nchunks);
    }
}

// ---- bank-aware entry placement ---------------------------------------------------------------------------
// A scan step covers 64 chunks, lane l holding chunks 2l and 2l+1: gather #j of the even (odd) chunks reads slot j of
// the 32 even (odd) chunks of the step at once; the shared-memory cost of that instruction is the largest number of
// lanes hitting one bank (bank = column mod 32).  A dot product does not
// care about the order of a row's entries, so each row's entries are re-dealt over its (chunk, slot) positions:
//   pass 1  slot by slot, take the row's most plentiful bank that this window has not used yet in that slot;
//   pass 2  what is left must collide: put it in the latest free slot, on the bank with the fewest lanes there.
// Random order costs 3.5 wavefronts per gather, this greedy ~2.0 (lower bound ~1.8 for 256-entry windows).
// One thread per stream part, rows in order; one-off work at index build.
constexpr int kPlaceMaxRow = 512;   // longer rows keep their original order

template <typename VT>
__global__ void __launch_bounds__(32) place_entries_kernel(uint16_t *cols16, VT *vals, const uint32_t *tails,
                                                           const uint32_t *part_win_begin, int n_parts, int n_cols,
                                                           int pair_mode) {
    const int part = blockIdx.x * blockDim.x + threadIdx.x;
    if (part >= n_parts) return;
    const uint64_t c_begin = (uint64_t)part_win_begin[part] * 32ull, c_end = (uint64_t)part_win_begin[part + 1] * 32ull;
    const uint16_t sent = (uint16_t)n_cols;
    const int sent_bank = n_cols & 31;
    uint16_t ecol[kPlaceMaxRow];
    uint16_t order[kPlaceMaxRow];    // entry indices grouped by bank
    VT eval[kPlaceMaxRow];
    uint8_t mult2[2][8][32];         // lanes per (chunk parity, slot, bank) in the current 64-chunk step
    uint32_t used2[2][8];
    uint16_t cnt[32], head[32];
    for (int q = 0; q < 2; ++q) for (int j = 0; j < 8; ++j) { used2[q][j] = 0; for (int b = 0; b < 32; ++b) mult2[q][j][b] = 0; }

    uint64_t c = c_begin;
    while (c < c_end) {
        // row = chunks [c, r_end]: r_end is the first chunk at or after c whose tail bit is set
        uint64_t r_end = c;
        while (r_end < c_end && !((tails[r_end >> 5] >> (r_end & 31)) & 1u)) ++r_end;
        if (r_end >= c_end) break;  // trailing padding chunks of the part (no row)
        const int nch = (int)(r_end - c + 1);
        int n = 0;
        bool fits = nch * 8 <= kPlaceMaxRow;
        if (fits) {
            for (int i = 0; i < nch * 8; ++i) {
                const uint16_t col = cols16[c * 8 + i];
                if (col != sent) {
                    ecol[n] = col;
                    if constexpr (sizeof(VT) > 1) eval[n] = vals[c * 8 + i];
                    ++n;
                }
            }
            for (int b = 0; b < 32; ++b) cnt[b] = 0;
            for (int i = 0; i < n; ++i) ++cnt[ecol[i] & 31];
            uint16_t run = 0;
            for (int b = 0; b < 32; ++b) { head[b] = run; run += cnt[b]; }
            {
                uint16_t fill[32];
                for (int b = 0; b < 32; ++b) fill[b] = head[b];
                for (int i = 0; i < n; ++i) order[fill[ecol[i] & 31]++] = (uint16_t)i;
            }
        }
        uint32_t avail = 0;
        if (fits) for (int b = 0; b < 32; ++b) if (cnt[b]) avail |= 1u << b;
        int remaining = n;
        for (int ch = 0; ch < nch; ++ch) {
            const uint64_t cc = c + ch;
            // conflict set of a gather: the 32 even (odd) chunks of a 64-chunk step in pair mode, else 32 consecutive chunks
            if ((cc & (pair_mode ? 63 : 31)) == 0)
                for (int q = 0; q < 2; ++q) for (int j = 0; j < 8; ++j) { used2[q][j] = 0; for (int b = 0; b < 32; ++b) mult2[q][j][b] = 0; }
            if (!fits) continue;
            uint32_t *used = used2[pair_mode ? (cc & 1) : 0];
            uint8_t (*mult)[32] = mult2[pair_mode ? (cc & 1) : 0];
            int slot_entry[8];
            for (int j = 0; j < 8; ++j) slot_entry[j] = -1;
            int n_here = remaining < 8 ? remaining : 8;
            int placed = 0;
            for (int j = 0; j < 8 && placed < n_here; ++j) {  // pass 1: conflict-free picks
                uint32_t cand = avail & ~used[j];
                int best = -1, bc = 0;
                while (cand) {
                    const int b = __ffs(cand) - 1;
                    cand &= cand - 1;
                    if (cnt[b] > bc) { bc = cnt[b]; best = b; }
                }
                if (best >= 0) {
                    slot_entry[j] = order[head[best] + --cnt[best]];
                    if (cnt[best] == 0) avail &= ~(1u << best);
                    used[j] |= 1u << best;
                    ++mult[j][best];
                    ++placed;
                }
            }
            for (int j = 7; j >= 0 && placed < n_here; --j) {  // pass 2: unavoidable collisions, latest slots first
                if (slot_entry[j] >= 0) continue;
                uint32_t cand = avail;
                int best = -1, bm = 255, bc = -1;
                while (cand) {
                    const int b = __ffs(cand) - 1;
                    cand &= cand - 1;
                    if (mult[j][b] < bm || (mult[j][b] == bm && cnt[b] > bc)) { bm = mult[j][b]; bc = cnt[b]; best = b; }
                }
                slot_entry[j] = order[head[best] + --cnt[best]];
                if (cnt[best] == 0) avail &= ~(1u << best);
                used[j] |= 1u << best;
                ++mult[j][best];
                ++placed;
            }
            remaining -= n_here;
            for (int j = 0; j < 8; ++j) {
                const int e = slot_entry[j];
                if (e >= 0) {
                    cols16[cc * 8 + j] = ecol[e];
                    if constexpr (sizeof(VT) > 1) vals[cc * 8 + j] = eval[e];
                } else {
                    cols16[cc * 8 + j] = sent;
                    if constexpr (sizeof(VT) > 1) vals[cc * 8 + j] = VT(0);
                    used[j] |= 1u << sent_bank;  // padding lanes all read the same (zero) slot: a broadcast
                }
            }
        }
        c = r_end + 1;
    }
}

// The scan kernel reads a chunk's "last chunk of its row" flag from bit 15 of the chunk's first entry
// (columns are < 32768), so it needs no row pointers and no separate mask stream.
__global__ void mark_tails_kernel(uint16_t *cols16, const uint32_t *tails, uint64_t n_chunks) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; c < n_chunks; c += stride)
        if ((tails[c >> 5] >> (c & 31)) & 1u) cols16[c * 8] |= 0x8000u;
}

struct NoVal { char c; };  // sizeof == 1: "no values" marker type

int build_ws_index(vs_index *idx, const void *d_crow, int crow_dtype, const void *d_col, int col_dtype,
                   const void *d_val, int val_dtype, cudaStream_t st) {
    const int64_t N = idx->n_rows;
    cudaDeviceProp prop;
    VS_CUDA(cudaGetDeviceProperties(&prop, idx->device));
    idx->n_ctas = prop.multiProcessorCount;
    idx->warps_per_cta = kScanWarps;
    idx->n_parts = idx->n_ctas * idx->warps_per_cta;
    const int P = idx->n_parts;

    uint64_t *d_cptr = nullptr, *d_nwin = nullptr;
    int *d_err = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    auto cleanup = [&]() { cudaFree(d_cptr); cudaFree(d_nwin); cudaFree(d_err); cudaFree(d_tmp); };

    VS_CUDA(cudaMalloc(&d_cptr, sizeof(uint64_t) * (size_t)(N + 1)));
    VS_CUDA(cudaMalloc(&d_nwin, sizeof(uint64_t)));
    VS_CUDA(cudaMalloc(&d_err, sizeof(int)));
    VS_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    {
        int64_t n = N + 1;
        row_chunks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_crow, crow_dtype, N, d_cptr, d_err);
    }
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cptr, d_cptr, (int64_t)(N + 1), st);
    VS_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cptr, d_cptr, (int64_t)(N + 1), st);

    VS_CUDA(cudaMalloc(&idx->part_row_begin, sizeof(uint32_t) * (size_t)(P + 1)));
    VS_CUDA(cudaMalloc(&idx->part_win_begin, sizeof(uint32_t) * (size_t)(P + 1)));
    part_rows_kernel<<<(P + 1 + 255) / 256, 256, 0, st>>>(d_cptr, N, P, idx->part_row_begin);
    part_windows_kernel<<<1, 32, 0, st>>>(d_cptr, idx->part_row_begin, P, idx->part_win_begin, d_nwin);
    uint64_t n_windows = 0;
    VS_CUDA(cudaMemcpyAsync(&n_windows, d_nwin, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VS_CUDA(cudaStreamSynchronize(st));
    if (n_windows >= (1ull << 32)) { cleanup(); VS_REQUIRE(false, VS_ERR_UNSUPPORTED, "index too large for one shard: %llu windows", (unsigned long long)n_windows); }
    idx->n_windows = n_windows;

    const uint64_t n_chunks = (n_windows + kStreamSlack) * 32ull;  // incl. prefetch slack (sentinel-filled)
    size_t val_elem = 0;
    if (idx->kind == 1) val_elem = (idx->store_dtype == VS_F32) ? 4 : 2;
    VS_CUDA(cudaMalloc(&idx->cols, n_chunks ? n_chunks * 16 : 16));
    VS_CUDA(cudaMalloc(&idx->tails, (n_windows + kStreamSlack) * 4));
    VS_CUDA(cudaMalloc(&idx->row_chunk, sizeof(uint32_t) * (size_t)(N + 1)));
    if (val_elem) {
        VS_CUDA(cudaMalloc(&idx->vals, n_chunks ? n_chunks * 8 * val_elem : 16));
        VS_CUDA(cudaMemsetAsync(idx->vals, 0, n_chunks * 8 * val_elem, st));
    }
    VS_CUDA(cudaMemsetAsync(idx->tails, 0, (n_windows + kStreamSlack) * 4, st));
    VS_CUDA(cudaMemsetAsync(idx->row_chunk, 0, sizeof(uint32_t) * (size_t)(N + 1), st));
    const uint32_t sent = (uint32_t)idx->n_cols | ((uint32_t)idx->n_cols << 16);
    if (n_chunks) fill_u32_kernel<<<2048, 256, 0, st>>>((uint32_t *)idx->cols, n_chunks * 4, sent);

    if (N > 0) {
        unsigned blocks = (unsigned)((N * 32 + 255) / 256);
#define VS_FILL(VT)                                                                                              \
    fill_rows_kernel<VT><<<blocks, 256, 0, st>>>(d_crow, crow_dtype, d_col, col_dtype, d_val, val_dtype, N,      \
                                                 idx->n_cols, d_cptr, idx->part_row_begin, idx->part_win_begin,  \
                                                 P, (uint16_t *)idx->cols, (VT *)idx->vals, idx->tails,          \
                                                 idx->row_chunk, d_err)
        if (idx->kind == 2) VS_FILL(NoVal);
        else if (idx->store_dtype == VS_F32) VS_FILL(float);
        else if (idx->store_dtype == VS_F16) VS_FILL(__half);
        else VS_FILL(__nv_bfloat16);
#undef VS_FILL
    }
    if (N > 0 && idx->bank_aware) {
        const unsigned pblocks = (unsigned)((P + 31) / 32);
        const int pair_mode = !(idx->kind == 1 && idx->store_dtype == VS_F32);  // scan.cu: fp32 values keep one chunk per lane
#define VS_PLACE(VT)                                                                                            \
    place_entries_kernel<VT><<<pblocks, 32, 0, st>>>((uint16_t *)idx->cols, (VT *)idx->vals, idx->tails,        \
                                                     idx->part_win_begin, P, (int)idx->n_cols, pair_mode)
        if (idx->kind == 2) VS_PLACE(NoVal);
        else if (idx->store_dtype == VS_F32) VS_PLACE(float);
        else if (idx->store_dtype == VS_F16) VS_PLACE(__half);
        else VS_PLACE(__nv_bfloat16);
#undef VS_PLACE
    }
    if (n_windows) mark_tails_kernel<<<2048, 256, 0, st>>>((uint16_t *)idx->cols, idx->tails, n_windows * 32ull);
    int h_err = 0;
    VS_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    VS_CUDA(cudaStreamSynchronize(st));
    VS_CUDA(cudaGetLastError());
    cleanup();
    VS_REQUIRE(h_err != 1, VS_ERR_INVALID, "crow_indices are not non-decreasing");
    VS_REQUIRE(h_err != 2, VS_ERR_INVALID, "col_indices outside [0, n_cols)");

    idx->stream_bytes = (int64_t)(n_windows * 32ull * (16 + 8 * val_elem) + n_windows * 4);
    idx->device_bytes = idx->stream_bytes + (int64_t)(sizeof(uint32_t) * (size_t)(N + 1 + 2 * (P + 1)));
    return VS_OK;
}

// ---- export: WS -> CSR (int64 crow/col, fp32 val), for SparseIndex.save (upstream index.py:181-202)
__global__ void export_len_kernel(const WsView idx, const uint32_t *row_chunk, uint64_t *len) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > idx.n_rows) return;
    if (r == idx.n_rows) { len[r] = 0; return; }
    const uint16_t *c16 = (const uint16_t *)idx.cols;
    uint64_t a = (uint64_t)row_chunk[r] * 8ull;
    // the row's chunk count comes from its tail bit: walk chunks until the tail
    uint64_t ch = row_chunk[r];
    while (!((idx.tails[ch >> 5] >> (ch & 31)) & 1u)) ++ch;
    uint64_t e = (ch + 1) * 8ull;
    uint64_t n = 0;
    for (uint64_t j = a; j < e; ++j) n += ((c16[j] & 0x7fffu) != (uint16_t)idx.n_cols);
    len[r] = n;
}

__global__ void export_fill_kernel(const WsView idx, const uint32_t *row_chunk, const uint64_t *crow,
                                   int64_t *out_crow, int64_t *out_col, float *out_val) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > idx.n_rows) return;
    out_crow[r] = (int64_t)crow[r];
    if (r == idx.n_rows) return;
    const uint16_t *c16 = (const uint16_t *)idx.cols;
    uint64_t a = (uint64_t)row_chunk[r] * 8ull;
    uint64_t o = crow[r], n = crow[r + 1] - crow[r];
    for (uint64_t j = 0, w = 0; w < n; ++j) {
        uint16_t c = c16[a + j] & 0x7fffu;   // bit 15 of a chunk's first entry is the tail flag
        if (c == (uint16_t)idx.n_cols) continue;
        out_col[o + w] = c;
        float v = 1.0f;
        if (idx.kind == 1) {
            if (idx.store_dtype == VS_F32) v = ((const float *)idx.vals)[a + j];
            else if (idx.store_dtype == VS_F16) v = __half2float(((const __half *)idx.vals)[a + j]);
            else v = __bfloat162float(((const __nv_bfloat16 *)idx.vals)[a + j]);
        }
        out_val[o + w] = v;
        ++w;
    }
}

int export_ws_csr(const vs_index *idx, int64_t *d_crow, int64_t *d_col, float *d_val, cudaStream_t st) {
    const int64_t N = idx->n_rows;
    uint64_t *d_len = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    VS_CUDA(cudaMalloc(&d_len, sizeof(uint64_t) * (size_t)(N + 1)));
    unsigned blocks = (unsigned)((N + 1 + 255) / 256);
    export_len_kernel<<<blocks, 256, 0, st>>>(ws_view(idx), idx->row_chunk, d_len);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, d_len, (int64_t)(N + 1), st);
    cudaError_t e = cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16);
    if (e != cudaSuccess) { cudaFree(d_len); VS_CUDA(e); }
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_len, d_len, (int64_t)(N + 1), st);
    export_fill_kernel<<<blocks, 256, 0, st>>>(ws_view(idx), idx->row_chunk, d_len, d_crow, d_col, d_val);
    e = cudaStreamSynchronize(st);
    cudaFree(d_len);
    cudaFree(d_tmp);
    VS_CUDA(e);
    VS_CUDA(cudaGetLastError());
    // the stream keeps each row's entries in bank-aware order: hand back ascending columns, like the input
    if (idx->nnz > 0) {
        int64_t *k2 = nullptr;
        float *v2 = nullptr;
        void *tmp = nullptr;
        size_t tb = 0;
        auto cleanup = [&]() { cudaFree(k2); cudaFree(v2); cudaFree(tmp); };
        VS_CUDA(cudaMalloc(&k2, (size_t)idx->nnz * 8));
        e = cudaMalloc(&v2, (size_t)idx->nnz * 4);
        if (e != cudaSuccess) { cleanup(); VS_CUDA(e); }
        cub::DeviceSegmentedSort::SortPairs(nullptr, tb, d_col, k2, d_val, v2, idx->nnz, N, d_crow, d_crow + 1, st);
        e = cudaMalloc(&tmp, tb ? tb : 16);
        if (e != cudaSuccess) { cleanup(); VS_CUDA(e); }
        cub::DeviceSegmentedSort::SortPairs(tmp, tb, d_col, k2, d_val, v2, idx->nnz, N, d_crow, d_crow + 1, st);
        cudaMemcpyAsync(d_col, k2, (size_t)idx->nnz * 8, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(d_val, v2, (size_t)idx->nnz * 4, cudaMemcpyDeviceToDevice, st);
        e = cudaStreamSynchronize(st);
        cleanup();
        VS_CUDA(e);
        VS_CUDA(cudaGetLastError());
    }
    return VS_OK;
}

}  // namespace vs
