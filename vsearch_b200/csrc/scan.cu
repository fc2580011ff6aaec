// scan.cu -- K1 (valued CSR) / K2 (binary bag-of-token) passage-major scan with the top-k
// fused in (K5), for sm_100a.  Replaces `torch.matmul(q, vector.t())` + `scores.topk(k)`
// (upstream src/ir/retriever/index.py:91-92) for SparseIndex / BoTIndex.
//
// One persistent CTA per SM, kScanWarps (24) warps.  Per query ("pass"):
//   1. the dense fp32 query vector (V+1 slots, slot V = 0 for padding) is pulled into shared
//      memory with 1-D bulk TMA copies (cp.async.bulk + mbarrier);
//   2. every warp streams ITS part of the WS index (index.cuh): one 512-byte window per step,
//      16 bytes per lane, D windows in flight per warp (register ring) -- loads never depend on
//      row pointers; each lane gathers q[col] for its 8 entries from shared memory;
//   3. a segmented warp scan over the window's tail mask turns lane partials into row scores;
//   4. rows whose rank key beats the CTA threshold go to a per-warp staging buffer, which is
//      flushed under a shared-memory lock into the CTA candidate buffer; when that fills, the
//      flushing warp alone radix-selects it down to k and raises the threshold while the other
//      31 warps keep streaming;
//   5. at the end of the pass the CTA selects its exact top-k and writes k keys to HBM.
// The [B, N] score matrix is never written.  A second tiny kernel (merge.cu) merges the
// per-CTA lists.
#include "index.cuh"
#include "topk.cuh"

namespace vs {

constexpr int kScanThreads = kScanWarps * 32;
using ScanGeom = TopkGeom<kScanWarps>;
constexpr int kSampleRegion = ScanGeom::kSample;          // sampling-phase keys per warp
constexpr int kSampleKeys = kScanWarps * kSampleRegion;   // <= kCapMax

struct ScanParams {
    const uint4 *cols;
    const void *vals;
    const uint32_t *tails;
    const uint32_t *part_win_begin;
    const uint32_t *part_row_begin;
    const float *q;        // [B, vpad] prepared queries
    uint64_t *cand;        // [B, n_ctas, k]
    float *scores_out;     // optional [B, N] (diagnostic vs_scores path), else nullptr
    int64_t n_rows;
    int B;
    int k;
    int cap;               // CTA candidate buffer entries (>= k + kStage)
    int vpad;              // floats per prepared query (multiple of 4, > V)
    int score_round;
    uint32_t sentinel;     // V | V << 16
};

// ---- per-chunk payload -----------------------------------------------------------------
template <int VT> struct Chunk;
template <> struct Chunk<0> { uint4 c; };
template <> struct Chunk<1> { uint4 c; uint4 v0, v1; };
template <> struct Chunk<2> { uint4 c; uint4 v; };
template <> struct Chunk<3> { uint4 c; uint4 v; };

template <int VT>
__device__ __forceinline__ void load_chunk(Chunk<VT> &ch, const uint4 *c, const uint4 *v) {
    ch.c = ldg_stream(c);
    if constexpr (VT == 1) {
        ch.v0 = ldg_stream(v);
        ch.v1 = ldg_stream(v + 1);
    } else if constexpr (VT >= 2) {
        ch.v = ldg_stream(v);
    }
}

__device__ __forceinline__ float half_lo(uint32_t x, int vt) {
    if (vt == 2) return __half2float(__ushort_as_half((unsigned short)(x & 0xffffu)));
    return __uint_as_float(x << 16);
}
__device__ __forceinline__ float half_hi(uint32_t x, int vt) {
    if (vt == 2) return __half2float(__ushort_as_half((unsigned short)(x >> 16)));
    return __uint_as_float(x & 0xffff0000u);
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// sum over the chunk's 8 entries of q[col] (* val).  qs = shared-space byte address of the query vector;
// address = qs + col * 4 is one LEA per entry after the 16-bit extract.
template <int VT>
__device__ __forceinline__ float chunk_dot(const Chunk<VT> &ch, const uint32_t qs) {
    const uint32_t w[4] = {ch.c.x, ch.c.y, ch.c.z, ch.c.w};
    float g[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        g[2 * i] = lds_f32(qs + (__byte_perm(w[i], 0, 0x4410) << 2));
        g[2 * i + 1] = lds_f32(qs + (__byte_perm(w[i], 0, 0x4432) << 2));
    }
    if constexpr (VT == 0) {
        return ((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]));
    } else if constexpr (VT == 1) {
        const uint32_t v[8] = {ch.v0.x, ch.v0.y, ch.v0.z, ch.v0.w, ch.v1.x, ch.v1.y, ch.v1.z, ch.v1.w};
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a = fmaf(g[2 * i], __uint_as_float(v[2 * i]), a);
            b = fmaf(g[2 * i + 1], __uint_as_float(v[2 * i + 1]), b);
        }
        return a + b;
    } else {
        const uint32_t v[4] = {ch.v.x, ch.v.y, ch.v.z, ch.v.w};
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a = fmaf(g[2 * i], half_lo(v[i], VT), a);
            b = fmaf(g[2 * i + 1], half_hi(v[i], VT), b);
        }
        return a + b;
    }
}

// Rare paths of a window, kept out of line so the streaming loop stays small.  They recompute their
// shared-memory pointers from the launch parameters instead of holding them in registers.
struct SmemLayout {
    uint64_t *cbuf;
    uint32_t *hist;
};
__device__ __forceinline__ SmemLayout smem_layout(uint8_t *smem, const ScanParams &p) {
    SmemLayout L;
    L.cbuf = reinterpret_cast<uint64_t *>(smem + (size_t)p.vpad * 4);
    L.hist = reinterpret_cast<uint32_t *>(L.cbuf + kCapMax);
    return L;
}

// One window of one warp: lane partial -> segmented scan -> row scores -> (SAMPLE) key into this warp's
// sampling region, or (!SAMPLE) threshold test + staging.
// PRE: every lane holds its chunk `cur`; T = tail mask of the window (warp-uniform).
template <int VT, bool ROUND, bool DIAG, bool SAMPLE>
__device__ __forceinline__ void process_window(const Chunk<VT> &cur, const uint32_t T, const uint32_t qs, const int lane,
                                               const uint32_t lt, float &carry, uint32_t &row, int &n_keys,
                                               const float tau_s, uint8_t *smem, CtaState *st, const ScanParams &p,
                                               const int b) {
    float v = chunk_dot<VT>(cur, qs);
    if (lane == 0) v += carry;
    // segmented inclusive scan: segments end at tail bits; `reach` = how many lanes back my segment extends
    const uint32_t before = T & lt;
    const int reach = lane - (32 - __clz(before));
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, v, d);
        if (reach >= d) v += o;
    }
    carry = __shfl_sync(0xffffffffu, v, 31);
    if (T >> 31) carry = 0.f;
    const int rank = __popc(before);
    const uint32_t rid = row + rank;
    row += __popc(T);
    const bool is_tail = (T >> lane) & 1u;
    float s = v;
    if constexpr (ROUND) s = round_score(v, p.score_round);
    if constexpr (DIAG) {
        if (is_tail) p.scores_out[(size_t)b * p.n_rows + rid] = s + 0.0f;
    }
    if constexpr (SAMPLE) {
        uint64_t *region = reinterpret_cast<uint64_t *>(smem + (size_t)p.vpad * 4) + (threadIdx.x >> 5) * kSampleRegion;
        if (is_tail) region[n_keys + rank] = make_key(s, rid);
        n_keys += __popc(T);
    } else {
        // cheap float pre-filter against the score part of the threshold (read once per group of windows);
        // exact 64-bit test only for survivors
        const bool maybe = is_tail && (s >= tau_s);
        if (__any_sync(0xffffffffu, maybe)) {
            const SmemLayout L = smem_layout(smem, p);
            const uint64_t key = make_key(s, rid);
            const uint64_t tau = *(volatile uint64_t *)&st->tau;
            private_insert<kScanWarps>(maybe && key > tau, key, L.cbuf, n_keys, lt);
        }
    }
}

template <int VT, int D, bool ROUND, bool DIAG>
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_kernel(const ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ CtaState st;
    const uint32_t qs = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    const int part = blockIdx.x * kScanWarps + warp;
    const uint32_t w_begin = p.part_win_begin[part];
    const int nwin = (int)(p.part_win_begin[part + 1] - w_begin);
    const uint32_t row0 = p.part_row_begin[part];
    const uint64_t chunk0 = (uint64_t)w_begin * 32ull + lane;
    const uint32_t q_bytes = (uint32_t)p.vpad * 4u;

    if (tid == 0) {
        mbar_init(&st.mbar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    uint32_t phase = 0;
    for (int b = 0; b < p.B; ++b) {
        // ---- stage the query vector: bulk TMA into shared memory
        if (tid == 0) {
            cta_state_reset(&st);
            fence_proxy_async();  // earlier generic-proxy reads of qs are ordered before the async writes
            mbar_arrive_expect_tx(&st.mbar, q_bytes);
            const uint8_t *src = reinterpret_cast<const uint8_t *>(p.q + (size_t)b * p.vpad);
            for (uint32_t off = 0; off < q_bytes; off += 16384u) {
                uint32_t n = min(16384u, q_bytes - off);
                bulk_g2s(smem + off, src + off, n, &st.mbar);
            }
        }
        __syncthreads();          // cnt/tau reset visible
        mbar_wait(&st.mbar, phase);
        phase ^= 1u;

        // ---- stream this warp's part.  The stream carries kStreamSlack windows of slack past its end, so the
        // prefetch ring never needs a bounds check: loads are unconditional, addresses are base + immediate.
        const uint4 *cp = p.cols + chunk0;
        const uint4 *vp = (VT == 1) ? (const uint4 *)p.vals + chunk0 * 2 : (const uint4 *)p.vals + chunk0;
        const uint32_t *tp = p.tails + w_begin;
        Chunk<VT> ring[D];
        uint32_t tring[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            load_chunk<VT>(ring[j], cp + j * 32, vp + j * (VT == 1 ? 64 : 32));
            tring[j] = ldg_stream_u32(tp + j);
        }
        float carry = 0.f;
        uint32_t row = row0;
        int n_keys = 0;   // phase A: keys in this warp's sample slice; phase B: keys in its private region
        int w = 0;
#define VS_STEP(SAMPLE, J)                                                                                         \
    {                                                                                                              \
        process_window<VT, ROUND, DIAG, SAMPLE>(ring[J], tring[J], qs, lane, lt, carry, row, n_keys, tau_s, smem,  \
                                                &st, p, b);                                                        \
        load_chunk<VT>(ring[J], cp + (D + J) * 32, vp + (D + J) * (VT == 1 ? 64 : 32));                            \
        tring[J] = ldg_stream_u32(tp + D + J);                                                                     \
    }
#define VS_ADVANCE()                                                                                               \
    {                                                                                                              \
        cp += D * 32;                                                                                              \
        vp += D * (VT == 1 ? 64 : 32);                                                                             \
        tp += D;                                                                                                   \
        w += D;                                                                                                    \
    }
        // ---- phase A: sampling.  Every row's key goes straight into this warp's region of cbuf (no threshold,
        // no lock) until the region cannot take D more windows; then ONE CTA-wide select sets the threshold.
        float tau_s = -INFINITY;
        while (w + D <= nwin && n_keys + D * 32 <= kSampleRegion) {
#pragma unroll
            for (int j = 0; j < D; ++j) VS_STEP(true, j)
            VS_ADVANCE()
        }
        {
            const SmemLayout L = smem_layout(smem, p);
            for (int i = n_keys + lane; i < kSampleRegion; i += 32) L.cbuf[warp * kSampleRegion + i] = 0ull;
            __syncthreads();
            cta_sample_select<kScanThreads>(L.cbuf, kSampleKeys, p.k, L.hist, &st);
            n_keys = 0;
        }
        // ---- phase B: steady state.  Once per group of D windows: one 64-bit "gate" load (float threshold + join
        // epoch), and a join when somebody asked for a re-selection or this warp's region cannot take D more windows.
        uint32_t epoch = 0;
        const SmemLayout L = smem_layout(smem, p);
        while (w + D <= nwin) {
            const uint64_t gate = gate_load(&st);
            tau_s = gate_tau_score(gate);
#pragma unroll
            for (int j = 0; j < D; ++j) VS_STEP(false, j)
            VS_ADVANCE()
            join_if_needed<kScanThreads, kScanWarps>(gate, epoch, D * 32, L.cbuf, n_keys, p.k, L.hist, &st);
        }
        tau_s = gate_tau_score(gate_load(&st));
#pragma unroll
        for (int j = 0; j < D - 1; ++j)
            if (w + j < nwin) process_window<VT, ROUND, DIAG, false>(ring[j], tring[j], qs, lane, lt, carry, row, n_keys,
                                                                     tau_s, smem, &st, p, b);
        finish_streaming<kScanThreads, kScanWarps>(epoch, L.cbuf, n_keys, p.k, L.hist, &st);
#undef VS_STEP
#undef VS_ADVANCE
        {
            // ---- exact top-k of this CTA's rows, written unsorted (merge.cu sorts)
            cta_write_topk<kScanThreads, kScanWarps>(L.cbuf, n_keys, p.k, L.hist, &st,
                                                     p.cand + ((size_t)b * gridDim.x + blockIdx.x) * (size_t)p.k);
        }
        __syncthreads();  // everyone is done with qs / cbuf before the next pass overwrites them
    }
}

// ---- query preparation: [B, ldq] any float dtype -> fp32 [B, vpad], zero padded, optionally
// rounded through the index dtype first (upstream index.py:89 `q.type(self.vector.dtype)`).
__global__ void prep_query_kernel(const void *q, int q_dtype, int64_t ldq, int64_t n_cols, int vpad, int round_mode,
                                  float *out) {
    const int64_t b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < vpad; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < n_cols) {
            if (q_dtype == VS_F32) v = ((const float *)q)[b * ldq + i];
            else if (q_dtype == VS_F16) v = __half2float(((const __half *)q)[b * ldq + i]);
            else v = __bfloat162float(((const __nv_bfloat16 *)q)[b * ldq + i]);
            v = round_score(v, round_mode);
        }
        out[b * (int64_t)vpad + i] = v;
    }
}

size_t scan_smem_bytes(int vpad, int cap) {
    (void)cap;
    return (size_t)vpad * 4 + (size_t)kCapMax * 8 + 256 * 4;
}

int scan_cap_for_k(int k) { (void)k; return kCapMax; }

int launch_prep_query(const void *d_q, int q_dtype, int64_t B, int64_t ldq, int64_t n_cols, int vpad, int round_mode,
                      float *d_out, cudaStream_t st) {
    if (B == 0) return VS_OK;
    dim3 grid((unsigned)((vpad + 255) / 256), (unsigned)B);
    VS_REQUIRE(B <= 65535, VS_ERR_INVALID, "query batch chunk too large");
    prep_query_kernel<<<grid, 256, 0, st>>>(d_q, q_dtype, ldq, n_cols, vpad, round_mode, d_out);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

template <int VT, int D, bool ROUND, bool DIAG>
static int launch_scan_t(const vs_index *idx, const ScanParams &p, size_t smem, cudaStream_t st) {
    auto kern = scan_topk_kernel<VT, D, ROUND, DIAG>;
    VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<idx->n_ctas, kScanThreads, smem, st>>>(p);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

template <int VT, int D>
static int launch_scan_v(const vs_index *idx, const ScanParams &p, size_t smem, cudaStream_t st) {
    if (p.scores_out) return launch_scan_t<VT, D, true, true>(idx, p, smem, st);  // diagnostic path
    if (p.score_round != VS_F32) return launch_scan_t<VT, D, true, false>(idx, p, smem, st);
    return launch_scan_t<VT, D, false, false>(idx, p, smem, st);
}

// d_qprep: [B, vpad] fp32; d_cand: [B, n_ctas, k] keys; d_scores_out optional [B, N]
int launch_scan(const vs_index *idx, const float *d_qprep, int vpad, int64_t B, int k, int score_round,
                uint64_t *d_cand, float *d_scores_out, cudaStream_t st) {
    ScanParams p;
    p.cols = idx->cols;
    p.vals = idx->vals;
    p.tails = idx->tails;
    p.part_win_begin = idx->part_win_begin;
    p.part_row_begin = idx->part_row_begin;
    p.q = d_qprep;
    p.cand = d_cand;
    p.scores_out = d_scores_out;
    p.n_rows = idx->n_rows;
    p.B = (int)B;
    p.k = k;
    p.cap = scan_cap_for_k(k);
    p.vpad = vpad;
    p.score_round = score_round;
    p.sentinel = (uint32_t)idx->n_cols | ((uint32_t)idx->n_cols << 16);
    size_t smem = scan_smem_bytes(vpad, p.cap);
    VS_REQUIRE(smem <= 227 * 1024, VS_ERR_UNSUPPORTED,
               "n_cols=%lld needs %zu bytes of shared memory (> 227 KB): vocabulary too large for the scan kernel",
               (long long)idx->n_cols, smem);
    static_assert(kStreamSlack >= 8, "prefetch ring deeper than the stream slack");
    static_assert(ScanGeom::kPrivate >= 2 * 4 * 32, "private region must hold two groups of windows");
    if (idx->kind == 2) return launch_scan_v<0, 4>(idx, p, smem, st);
    if (idx->store_dtype == VS_F32) return launch_scan_v<1, 2>(idx, p, smem, st);
    if (idx->store_dtype == VS_F16) return launch_scan_v<2, 3>(idx, p, smem, st);
    return launch_scan_v<3, 3>(idx, p, smem, st);
}

}  // namespace vs
