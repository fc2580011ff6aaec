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
    if (dtype == VS_U16) return load_idx<uint16_t>(p, i);   // columns straight from a loaded shard file (npz.cu)
    return dtype == VS_I32 ? load_idx<int32_t>(p, i) : load_idx<int64_t>(p, i);
}
__device__ __forceinline__ float load_value(const void *p, int dtype, int64_t i) {
    if (dtype == VS_F32) return ((const float *)p)[i];
    if (dtype == VS_F16) return __half2float(((const __half *)p)[i]);
    return __bfloat162float(((const __nv_bfloat16 *)p)[i]);
}

// chunks per row (>= 1 so that every row, even an empty one, owns a tail bit).  col_shift > 0 (upstream's
// `mat[:, shift:]`, index.py:174): entries in columns < col_shift are dropped; the surviving entries are counted into
// *kept_nnz (one atomic per CTA).
__global__ void __launch_bounds__(256) row_chunks_kernel(const void *crow, int crow_dtype, const void *col, int col_dtype,
                                                         int64_t col_shift, int64_t n_rows, int64_t nnz, uint64_t *rc,
                                                         unsigned long long *kept_nnz, int *err) {
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t len = 0;
    if (r < n_rows) {
        int64_t a = load_index(crow, crow_dtype, r), e = load_index(crow, crow_dtype, r + 1);
        if (a < 0 || e > nnz || (r == 0 && a != 0)) { atomicExch(err, 3); a = e = 0; }   // row pointers must stay inside col / val
        if (e < a) { atomicExch(err, 1); e = a; }
        len = (uint64_t)(e - a);
        if (col_shift > 0) {
            len = 0;
            for (int64_t j = a; j < e; ++j) len += load_index(col, col_dtype, j) >= col_shift;
        }
        rc[r] = len == 0 ? 1 : (len + 7) / 8;
    } else if (r == n_rows) {
        rc[r] = 0;
    }
    if (kept_nnz != nullptr) {
        if (len) atomicAdd(&s_sum, (unsigned long long)len);
        __syncthreads();
        if (threadIdx.x == 0 && s_sum) atomicAdd(kept_nnz, s_sum);
    }
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

__global__ void part_windows_kernel(const uint64_t *cptr, const uint32_t *part_row_begin, int n_parts, int cpl_shift,
                                    uint32_t *part_win_begin, uint64_t *n_windows_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint64_t w = 0;
    const uint64_t sc = 32ull << cpl_shift;   // chunks per scan step
    for (int p = 0; p < n_parts; ++p) {
        part_win_begin[p] = (uint32_t)w;
        uint64_t chunks = cptr[part_row_begin[p + 1]] - cptr[part_row_begin[p]];
        w += ((chunks + sc - 1) / sc) << cpl_shift;   // whole steps (32 << cpl_shift chunks each)
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
                                 int val_dtype, int64_t n_rows, int64_t n_cols, int64_t nnz, const uint64_t *cptr,
                                 const uint32_t *part_row_begin, const uint32_t *part_win_begin, int n_parts, int cpl_shift,
                                 int64_t col_shift, uint16_t *cols16, VT *vals, uint32_t *tails, uint32_t *row_chunk, int *err) {
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
    if (a < 0 || e > nnz || (r == 0 && a != 0)) a = e = 0;   // same clamps as row_chunks_kernel (which raised the error)
    if (e < a) e = a;
    uint64_t kept = 0;   // entries of this row written so far (columns < col_shift are dropped, the rest move up)
    for (int64_t j0 = a; j0 < e; j0 += 32) {
        const int64_t j = j0 + lane;
        const int64_t c = j < e ? load_index(col, col_dtype, j) - col_shift : -1;
        const bool in = j < e && c >= 0;
        if (in && c >= n_cols) atomicExch(err, 2);   // leaves a sentinel in place
        if (j < e && c < 0 && c + col_shift < 0) atomicExch(err, 2);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        const uint64_t ent = kept + (uint64_t)__popc(m & ((1u << lane) - 1u));
        kept += (uint64_t)__popc(m);
        if (!in || c >= n_cols) continue;
        const uint64_t o = ws_phys_chunk(dst + (ent >> 3), cpl_shift) * 8ull + (ent & 7ull);
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

struct NoVal { char c; };  // sizeof == 1: "no values" marker type

// ---- bank-aware entry placement ---------------------------------------------------------------------------
// A scan step covers 32*CPL chunks (CPL = chunks per lane: 8 for the binary kernel, 2 for 16-bit values, 1 for fp32
// values).  Gather #j of sub-chunk q reads slot j of the 32 chunks {CPL*l + q} of the step at once; the shared-memory
// cost of that instruction is the largest number of lanes hitting one bank (bank = column mod 32).  A dot product
// does not care about the order of a row's entries, so inside every step each row's entries are re-dealt over the
// row's (chunk, slot) positions of that step: for every gather group (slot, sub-chunk) a maximum bipartite matching
// lanes -> banks (edge = the lane's row still has an entry in that bank; banks tried in order of the row's remaining
// supply; Kuhn's augmenting paths) gives every lane a DISTINCT bank whenever one exists; unmatched lanes collide on
// the emptiest bank their row can still supply.
// The bound is the per-step bank imbalance of the columns: a step of 8*CPL gathers cannot cost fewer wavefronts than
// its most loaded bank holds entries.  Random columns: 3.5 wavefronts per gather in input order, 1.83 after the
// matching with CPL = 2 (64-chunk steps), lower with CPL = 8 (256-chunk steps: the imbalance averages out over 64
// gathers; measured numbers in profiles/README.md).
// One warp per step (steps are independent: entries never leave their step), state in shared memory; the search
// itself is sequential but every inner loop over banks is one ballot / redux.  One-off work at index build.
// Positions inside the scratch are LOGICAL (chunk c = CPL*lane + q); the global arrays are addressed physically
// (chunk c of the step sits at q*32 + lane, index.cuh).
template <typename VT, int CPL>
struct alignas(16) PlaceScratch {
    static constexpr int SC = 32 * CPL;   // chunks per step
    static constexpr int NP = SC * 8;     // entry positions per step
    static constexpr int NV = sizeof(VT) > 1 ? NP : 8;
    uint16_t ecol[NP];                    // the step's entries as loaded (sentinel = padding)
    uint16_t ocol[NP];                    // ... and as re-dealt
    VT eval[NV];
    VT oval[NV];
    static constexpr int MS = SC < 64 ? SC : 64;   // segments a step may hold and still be re-dealt (see below)
    int16_t nxt[NP];                      // per (segment, bank) linked lists of unused entries
    int16_t head[MS][32];
    uint16_t sup[MS][32];                 // remaining entries per (segment, bank); segment = piece of a row in the step
    int16_t rem[MS];                      // remaining entries per segment
    uint8_t chunk_seg[SC];
    uint32_t tw[CPL];                     // row-end flags of the step's chunks (logical order)
    uint32_t tpre[CPL + 1];               // exclusive prefix of their popcounts
};

template <typename VT, int CPL, int PW>
__global__ void __launch_bounds__(PW * 32) place_step_kernel(uint16_t *cols16, VT *vals, const uint32_t *tails,
                                                             uint64_t n_steps, int n_cols) {
    using S = PlaceScratch<VT, CPL>;
    constexpr int SC = S::SC, NP = S::NP;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool kVals = sizeof(VT) > 1;
    extern __shared__ __align__(16) unsigned char place_smem[];
    S &s = reinterpret_cast<S *>(place_smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const uint16_t sent = (uint16_t)n_cols;
    const uint64_t n_warps = (uint64_t)gridDim.x * PW;

    for (uint64_t step = (uint64_t)blockIdx.x * PW + (threadIdx.x >> 5); step < n_steps; step += n_warps) {
        const uint64_t c0 = step * SC;
        if (lane < CPL) s.tw[lane] = tails[step * CPL + lane];
        __syncwarp();
        if (lane == 0) {
            uint32_t run = 0;
            for (int w = 0; w < CPL; ++w) { s.tpre[w] = run; run += __popc(s.tw[w]); }
            s.tpre[CPL] = run;
        }
        __syncwarp();
        // segments = pieces of rows inside the step: one more than the row ends among the first SC-1 chunks
        const int n_seg = (int)s.tpre[CPL] - (int)((s.tw[CPL - 1] >> 31) & 1u) + 1;
        // a step of more than MS row pieces (rows of < 4 chunks on average) keeps its input order: such rows have few
        // entries to re-deal, and the per-segment tables stay small enough for 8 warps per SM
        if (n_seg > S::MS) { __syncwarp(); continue; }
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int c = CPL * lane + q;
            const uint64_t pc = c0 + (uint64_t)q * 32 + lane;   // physical chunk
            *reinterpret_cast<uint4 *>(&s.ecol[c * 8]) = *reinterpret_cast<const uint4 *>(cols16 + pc * 8);
            if constexpr (kVals) {
                constexpr int kVec = (int)sizeof(VT) * 8 / 16;
#pragma unroll
                for (int t = 0; t < kVec; ++t)
                    reinterpret_cast<uint4 *>(&s.eval[c * 8])[t] = reinterpret_cast<const uint4 *>(vals + pc * 8)[t];
            }
            s.chunk_seg[c] = (uint8_t)(s.tpre[c >> 5] + __popc(s.tw[c >> 5] & ((1u << (c & 31)) - 1u)));
        }
        for (int r = 0; r < n_seg; ++r) { s.head[r][lane] = -1; s.sup[r][lane] = 0; }
        __syncwarp();
        // lane = bank: thread every entry of that bank onto its segment's list (no two lanes touch the same list)
        for (int pos = 0; pos < NP; ++pos) {
            const uint16_t col = s.ecol[pos];
            if (col != sent && (col & 31) == lane) {
                const int r = s.chunk_seg[pos >> 3];
                s.nxt[pos] = s.head[r][lane];
                s.head[r][lane] = (int16_t)pos;
                ++s.sup[r][lane];
            }
        }
        __syncwarp();
        for (int r = lane; r < n_seg; r += 32) {
            int t = 0;
            for (int b = 0; b < 32; ++b) t += s.sup[r][b];
            s.rem[r] = (int16_t)t;
        }
        __syncwarp();

        // ---- gather groups: slot j of sub-chunk q -> one position per lane
        for (int j = 0; j < 8; ++j)
            for (int q = 0; q < CPL; ++q) {
                const int c = CPL * lane + q;
                const int r = s.chunk_seg[c];
                const unsigned same = __match_any_sync(FULL, r);
                const int rank = __popc(same & lt), remr = s.rem[r];
                const bool part = rank < remr;  // the segment still has an entry for this lane
                const int ls = part ? r : -1;
                __syncwarp();
                if (rank == 0) s.rem[r] = (int16_t)(remr - min(__popc(same), remr));
                int bl = -1;   // lane b: the lane matched to bank b
                int lb = -1;   // the bank matched to this lane
                int st_lane = 0, st_bank = 0;  // lane t: DFS stack entry t
                for (unsigned pm = __ballot_sync(FULL, part); pm; pm &= pm - 1) {  // Kuhn's augmenting paths
                    const int l0 = __ffs(pm) - 1;
                    unsigned visited = 0;
                    int sp = 0;
                    if (lane == 0) st_lane = l0;
                    while (sp >= 0) {
                        const int l = __shfl_sync(FULL, st_lane, sp);
                        const int rr = __shfl_sync(FULL, ls, l);
                        const unsigned sv = s.sup[rr][lane];
                        const unsigned key = (sv > 0 && !((visited >> lane) & 1u)) ? ((sv << 5) | (31u - lane)) + 1u : 0u;
                        const unsigned mx = __reduce_max_sync(FULL, key);  // unvisited bank with the largest supply
                        if (mx == 0) { --sp; continue; }
                        const int b = 31 - (int)((mx - 1u) & 31u);
                        visited |= 1u << b;
                        if (lane == sp) st_bank = b;
                        const int cur = __shfl_sync(FULL, bl, b);
                        if (cur < 0) {  // free bank: flip the path
                            for (int t = 0; t <= sp; ++t) {
                                const int tl = __shfl_sync(FULL, st_lane, t), tb = __shfl_sync(FULL, st_bank, t);
                                if (lane == tb) bl = tl;
                                if (lane == tl) lb = tb;
                            }
                            break;
                        }
                        ++sp;
                        if (lane == sp) st_lane = cur;
                    }
                }
                int e = -1;
                if (lb >= 0) { e = s.head[ls][lb]; s.head[ls][lb] = s.nxt[e]; --s.sup[ls][lb]; }
                __syncwarp();
                // lanes without a distinct bank collide on the emptiest bank their segment can still supply
                unsigned mult = bl >= 0 ? 1u : 0u;
                for (unsigned um = __ballot_sync(FULL, part && lb < 0); um; um &= um - 1) {
                    const int l = __ffs(um) - 1;
                    const int rr = __shfl_sync(FULL, ls, l);
                    const unsigned sv = s.sup[rr][lane];
                    const unsigned key = sv > 0 ? (((255u - mult) << 21) | (min(sv, 0xffffu) << 5) | (31u - lane)) + 1u : 0u;
                    const unsigned mx = __reduce_max_sync(FULL, key);
                    if (mx == 0) continue;  // cannot happen: part guarantees supply
                    const int b = 31 - (int)((mx - 1u) & 31u);
                    if (lane == b) ++mult;
                    if (lane == l) { e = s.head[rr][b]; s.head[rr][b] = s.nxt[e]; --s.sup[rr][b]; }
                    __syncwarp();
                }
                const int pos = c * 8 + j;
                s.ocol[pos] = e >= 0 ? s.ecol[e] : sent;
                if constexpr (kVals) {
                    if (e >= 0) s.oval[pos] = s.eval[e];
                    else memset(&s.oval[pos], 0, sizeof(VT));
                }
                __syncwarp();
            }
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int c = CPL * lane + q;
            const uint64_t pc = c0 + (uint64_t)q * 32 + lane;
            *reinterpret_cast<uint4 *>(cols16 + pc * 8) = *reinterpret_cast<const uint4 *>(&s.ocol[c * 8]);
            if constexpr (kVals) {
                constexpr int kVec = (int)sizeof(VT) * 8 / 16;
#pragma unroll
                for (int t = 0; t < kVec; ++t)
                    reinterpret_cast<uint4 *>(vals + pc * 8)[t] = reinterpret_cast<const uint4 *>(&s.oval[c * 8])[t];
            }
        }
        __syncwarp();
    }
}

template <typename VT, int CPL>
static int launch_place(uint16_t *cols16, VT *vals, const uint32_t *tails, uint64_t n_steps, int n_cols, int sms, cudaStream_t st) {
    constexpr int PW = 8;   // warps per CTA (the scratch of a 256-chunk step is 21 KB)
    const size_t smem = sizeof(PlaceScratch<VT, CPL>) * PW;
    auto kern = place_step_kernel<VT, CPL, PW>;
    VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)std::min<uint64_t>((n_steps + PW - 1) / PW, (uint64_t)sms * (CPL >= 8 ? 1 : 2));
    VS_REQUIRE(smem <= 227 * 1024, VS_ERR_UNSUPPORTED, "placement scratch does not fit shared memory");
    kern<<<blocks, PW * 32, smem, st>>>(cols16, vals, tails, n_steps, n_cols);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

// The scan kernel reads a chunk's "last chunk of its row" flag from bit 15 of the chunk's first entry
// (columns are < 32768), so it needs no row pointers and no separate mask stream.
__global__ void mark_tails_kernel(uint16_t *cols16, const uint32_t *tails, uint64_t n_chunks, int cpl_shift) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // logical chunk
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; c < n_chunks; c += stride)
        if ((tails[c >> 5] >> (c & 31)) & 1u) cols16[ws_phys_chunk(c, cpl_shift) * 8] |= 0x8000u;
}


int build_ws_index(vs_index *idx, const void *d_crow, int crow_dtype, const void *d_col, int col_dtype,
                   const void *d_val, int val_dtype, cudaStream_t st, int64_t col_shift) {
    const int64_t N = idx->n_rows;
    cudaDeviceProp prop;
    VS_CUDA(cudaGetDeviceProperties(&prop, idx->device));
    idx->n_ctas = prop.multiProcessorCount;
    idx->warps_per_cta = kScanWarps;
    idx->n_parts = idx->n_ctas * idx->warps_per_cta;
    const int P = idx->n_parts;

    uint64_t *d_cptr = nullptr, *d_nwin = nullptr;   // d_nwin[0] = windows, d_nwin[1] = entries kept by the column shift
    int *d_err = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    auto cleanup = [&]() { cudaFree(d_cptr); cudaFree(d_nwin); cudaFree(d_err); cudaFree(d_tmp); };

    VS_CUDA(cudaMalloc(&d_cptr, sizeof(uint64_t) * (size_t)(N + 1)));
    VS_CUDA(cudaMalloc(&d_nwin, 2 * sizeof(uint64_t)));
    VS_CUDA(cudaMalloc(&d_err, sizeof(int)));
    VS_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    VS_CUDA(cudaMemsetAsync(d_nwin, 0, 2 * sizeof(uint64_t), st));
    const int64_t nnz_in = idx->nnz;   // entries of the input arrays (row pointers are checked against it)
    {
        int64_t n = N + 1;
        row_chunks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_crow, crow_dtype, d_col, col_dtype, col_shift, N, nnz_in, d_cptr,
                                                                       col_shift > 0 ? (unsigned long long *)(d_nwin + 1) : nullptr, d_err);
    }
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cptr, d_cptr, (int64_t)(N + 1), st);
    VS_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cptr, d_cptr, (int64_t)(N + 1), st);

    VS_CUDA(cudaMalloc(&idx->part_row_begin, sizeof(uint32_t) * (size_t)(P + 1)));
    VS_CUDA(cudaMalloc(&idx->part_win_begin, sizeof(uint32_t) * (size_t)(P + 1)));
    part_rows_kernel<<<(P + 1 + 255) / 256, 256, 0, st>>>(d_cptr, N, P, idx->part_row_begin);
    // chunks per lane per scan step (scan.cu): 8 binary, 2 fp16/bf16 values, 1 fp32 values
    idx->cpl_shift = idx->kind == 2 ? 3 : (idx->store_dtype == VS_F32 ? 0 : 1);
    const int cpl_shift = idx->cpl_shift;
    part_windows_kernel<<<1, 32, 0, st>>>(d_cptr, idx->part_row_begin, P, cpl_shift, idx->part_win_begin, d_nwin);
    uint64_t h_nwin[2] = {0, 0};
    VS_CUDA(cudaMemcpyAsync(h_nwin, d_nwin, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VS_CUDA(cudaStreamSynchronize(st));
    const uint64_t n_windows = h_nwin[0];
    if (col_shift > 0) idx->nnz = (int64_t)h_nwin[1];   // what survives `[:, shift:]`
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
                                                 idx->n_cols, nnz_in, d_cptr, idx->part_row_begin,               \
                                                 idx->part_win_begin,                                           \
                                                 P, cpl_shift, col_shift, (uint16_t *)idx->cols, (VT *)idx->vals, idx->tails, \
                                                 idx->row_chunk, d_err)
        if (idx->kind == 2) VS_FILL(NoVal);
        else if (idx->store_dtype == VS_F32) VS_FILL(float);
        else if (idx->store_dtype == VS_F16) VS_FILL(__half);
        else VS_FILL(__nv_bfloat16);
#undef VS_FILL
    }
    if (N > 0 && idx->bank_aware && n_windows) {
        const uint64_t n_steps = n_windows >> cpl_shift;
        int rc;
        if (idx->kind == 2) rc = launch_place<NoVal, 8>((uint16_t *)idx->cols, (NoVal *)nullptr, idx->tails, n_steps, (int)idx->n_cols, idx->n_ctas, st);
        else if (idx->store_dtype == VS_F32) rc = launch_place<float, 1>((uint16_t *)idx->cols, (float *)idx->vals, idx->tails, n_steps, (int)idx->n_cols, idx->n_ctas, st);
        else if (idx->store_dtype == VS_F16) rc = launch_place<__half, 2>((uint16_t *)idx->cols, (__half *)idx->vals, idx->tails, n_steps, (int)idx->n_cols, idx->n_ctas, st);
        else rc = launch_place<__nv_bfloat16, 2>((uint16_t *)idx->cols, (__nv_bfloat16 *)idx->vals, idx->tails, n_steps, (int)idx->n_cols, idx->n_ctas, st);
        if (rc != VS_OK) return rc;
    }
    if (n_windows) mark_tails_kernel<<<2048, 256, 0, st>>>((uint16_t *)idx->cols, idx->tails, n_windows * 32ull, cpl_shift);
    int h_err = 0;
    VS_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    VS_CUDA(cudaStreamSynchronize(st));
    VS_CUDA(cudaGetLastError());
    cleanup();
    VS_REQUIRE(h_err != 1, VS_ERR_INVALID, "crow_indices are not non-decreasing");
    VS_REQUIRE(h_err != 3, VS_ERR_INVALID, "crow_indices must start at 0 and stay within [0, nnz]");
    VS_REQUIRE(h_err != 2, VS_ERR_INVALID, "col_indices outside [0, n_cols)");

    idx->stream_bytes = (int64_t)(n_windows * 32ull * (16 + 8 * val_elem) + n_windows * 4);
    idx->device_bytes = idx->stream_bytes + (int64_t)(sizeof(uint32_t) * (size_t)(N + 1 + 2 * (P + 1)));
    return VS_OK;
}

// ---- diagnostic: shared-memory wavefronts the scan's gathers cost on this index (distinct addresses per bank; equal
// columns broadcast).  One warp per step; out[0] += wavefronts, out[1] += gather instructions.
__global__ void gather_wavefronts_kernel(const uint16_t *cols16, uint64_t n_steps, int cpl, unsigned long long *out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long wf = 0, ng = 0;
    for (uint64_t step = warp; step < n_steps; step += n_warps)
        for (int i = 0; i < cpl; ++i) {
            const uint16_t *ch = cols16 + ((step * cpl + i) * 32 + lane) * 8;
            for (int j = 0; j < 8; ++j) {
                const unsigned col = ch[j] & (j == 0 ? 0x7fffu : 0xffffu);
                const unsigned same_col = __match_any_sync(0xffffffffu, col);
                const bool leader = (__ffs(same_col) - 1) == lane;   // one access per distinct address
                const unsigned same_bank = __match_any_sync(0xffffffffu, leader ? (col & 31u) : 0xffffu + lane);
                const int m = leader ? __popc(same_bank) : 1;
                wf += __reduce_max_sync(0xffffffffu, (unsigned)m);
                ng += 1;
            }
        }
    if (lane == 0) { atomicAdd(&out[0], wf); atomicAdd(&out[1], ng); }
}

int debug_gather_wavefronts(const vs_index *idx, unsigned long long *d_out2, cudaStream_t st) {
    VS_CUDA(cudaMemsetAsync(d_out2, 0, 16, st));
    const uint64_t n_steps = idx->n_windows >> idx->cpl_shift;
    if (n_steps) gather_wavefronts_kernel<<<1024, 256, 0, st>>>((const uint16_t *)idx->cols, n_steps, 1 << idx->cpl_shift, d_out2);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

// ---- export: WS -> CSR (int64 crow/col, fp32 val), for SparseIndex.save (upstream index.py:181-202)
__global__ void export_len_kernel(const WsView idx, const uint32_t *row_chunk, uint64_t *len) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > idx.n_rows) return;
    if (r == idx.n_rows) { len[r] = 0; return; }
    const uint16_t *c16 = (const uint16_t *)idx.cols;
    // the row's chunk count comes from its tail bit: walk (logical) chunks until the tail
    uint64_t n = 0;
    for (uint64_t ch = row_chunk[r];; ++ch) {
        const uint64_t a = ws_phys_chunk(ch, idx.cpl_shift) * 8ull;
        for (int j = 0; j < 8; ++j) n += ((c16[a + j] & 0x7fffu) != (uint16_t)idx.n_cols);
        if ((idx.tails[ch >> 5] >> (ch & 31)) & 1u) break;
    }
    len[r] = n;
}

__global__ void export_fill_kernel(const WsView idx, const uint32_t *row_chunk, const uint64_t *crow,
                                   int64_t *out_crow, int64_t *out_col, float *out_val) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > idx.n_rows) return;
    out_crow[r] = (int64_t)crow[r];
    if (r == idx.n_rows) return;
    const uint16_t *c16 = (const uint16_t *)idx.cols;
    const uint64_t ch0 = row_chunk[r];
    uint64_t o = crow[r], n = crow[r + 1] - crow[r];
    for (uint64_t j = 0, w = 0; w < n; ++j) {
        const uint64_t at = ws_phys_chunk(ch0 + (j >> 3), idx.cpl_shift) * 8ull + (j & 7ull);
        uint16_t c = c16[at] & 0x7fffu;   // bit 15 of a chunk's first entry is the tail flag
        if (c == (uint16_t)idx.n_cols) continue;
        out_col[o + w] = c;
        float v = 1.0f;
        if (idx.kind == 1) {
            if (idx.store_dtype == VS_F32) v = ((const float *)idx.vals)[at];
            else if (idx.store_dtype == VS_F16) v = __half2float(((const __half *)idx.vals)[at]);
            else v = __bfloat162float(((const __nv_bfloat16 *)idx.vals)[at]);
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
