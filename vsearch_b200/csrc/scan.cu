// scan.cu -- K1 (valued CSR) / K2 (binary bag-of-token) passage-major scan with the top-k
// fused in (K5), for sm_100a.  Replaces `torch.matmul(q, vector.t())` + `scores.topk(k)`
// (upstream src/ir/retriever/index.py:91-92) for SparseIndex / BoTIndex.
//
// One persistent CTA per SM, kScanWarps (24) warps.  Per query ("pass"):
//   1. the dense fp32 query vector (V+1 slots, slot V = 0 for padding) is pulled into shared
//      memory with 1-D bulk TMA copies (cp.async.bulk + mbarrier);
//   2. every warp streams ITS part of the WS index (index.cuh): one 1024-byte double window (64
//      chunks) per step, two adjacent 16-byte chunks per lane, D steps in flight per warp (register
//      ring) -- loads never depend on row pointers, the row-end ("tail") flag of a chunk rides in
//      bit 15 of its first entry; each lane gathers q[col] for its 16 entries from shared memory;
//   3. ONE segmented warp scan per 64 chunks (6 shuffles) turns lane partials into row scores;
//   4. the first rows of a pass are sampled into shared memory and ONE CTA-wide radix select sets the threshold;
//      afterwards a float pre-filter rejects almost every row, rows whose rank key beats the threshold are appended
//      to the warp's PRIVATE region with plain stores, and a warp whose region runs low raises a join epoch that all
//      warps poll: one CTA-wide re-selection, no locks, no per-row atomics (topk.cuh);
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
        // entry 0 carries the chunk's tail flag in bit 15 (columns are < 32768)
        g[2 * i] = lds_f32(qs + ((i == 0 ? (w[i] & 0x7fffu) : __byte_perm(w[i], 0, 0x4410)) << 2));
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

// One step of one warp = 64 chunks; lane l holds chunks 2l (A) and 2l+1 (B).  Lane partials -> ONE segmented warp
// scan -> row scores -> (SAMPLE) keys into this warp's sampling slice, or (!SAMPLE) threshold test + private region.
template <int VT, bool ROUND, bool DIAG, bool SAMPLE>
__device__ __forceinline__ void process_pair(const Chunk<VT> &ca, const Chunk<VT> &cb, const uint32_t qs, const int lane,
                                             const uint32_t lt, float &carry, uint32_t &row, int &n_keys,
                                             const float tau_s, uint8_t *smem, CtaState *st, const ScanParams &p,
                                             const int b) {
    const bool tail_a = (ca.c.x & 0x8000u) != 0, tail_b = (cb.c.x & 0x8000u) != 0;
    const float a = chunk_dot<VT>(ca, qs), bsum = chunk_dot<VT>(cb, qs);
    const uint32_t TA = __ballot_sync(0xffffffffu, tail_a), TB = __ballot_sync(0xffffffffu, tail_b);
    const uint32_t F = TA | TB;  // lanes in which a row ends
    // v = what this lane hands on to its right neighbour: the part after its last row end
    float v = tail_b ? 0.f : (tail_a ? bsum : a + bsum);
    if (lane == 0 && !(F & 1u)) v += carry;  // lane 0 passes the carry through unless a row ends inside it
    // inclusive segmented scan of v; a flagged lane RESTARTS the sum (its v is already post-row-end)
    const uint32_t upto = F & (lt | (1u << lane));
    const int reach = upto ? lane - (31 - __clz(upto)) : lane;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, v, d);
        if (reach >= d) v += o;
    }
    const float rot = __shfl_sync(0xffffffffu, v, (lane + 31) & 31);  // left neighbour's running sum; lane 0 gets lane 31's
    const float incoming = (lane == 0) ? carry : rot;
    carry = rot;  // only lane 0's copy is used: the open row's partial sum at the end of this step
    float sa = incoming + a;                        // score of the row ending in chunk A
    float sb = tail_a ? bsum : incoming + a + bsum;  // score of the row ending in chunk B
    const int rank = __popc(TA & lt) + __popc(TB & lt);  // rows completed by the lanes before me
    const uint32_t rid_a = row + rank, rid_b = rid_a + (tail_a ? 1u : 0u);
    const int n_rows_step = __popc(TA) + __popc(TB);
    row += n_rows_step;
    if constexpr (ROUND) { sa = round_score(sa, p.score_round); sb = round_score(sb, p.score_round); }
    if constexpr (DIAG) {
        if (tail_a) p.scores_out[(size_t)b * p.n_rows + rid_a] = sa + 0.0f;
        if (tail_b) p.scores_out[(size_t)b * p.n_rows + rid_b] = sb + 0.0f;
    }
    if constexpr (SAMPLE) {
        uint64_t *region = reinterpret_cast<uint64_t *>(smem + (size_t)p.vpad * 4) + (threadIdx.x >> 5) * kSampleRegion;
        if (tail_a) region[n_keys + rank] = make_key(sa, rid_a);
        if (tail_b) region[n_keys + rank + (tail_a ? 1 : 0)] = make_key(sb, rid_b);
        n_keys += n_rows_step;
    } else {
        // cheap float pre-filter against the score part of the threshold (read once per group of steps);
        // exact 64-bit test only for survivors
        const bool maybe_a = tail_a && (sa >= tau_s), maybe_b = tail_b && (sb >= tau_s);
        if (__any_sync(0xffffffffu, maybe_a || maybe_b)) {
            const SmemLayout L = smem_layout(smem, p);
            const uint64_t tau = *(volatile uint64_t *)&st->tau;
            const uint64_t ka = make_key(sa, rid_a), kb = make_key(sb, rid_b);
            private_insert<kScanWarps>(maybe_a && ka > tau, ka, L.cbuf, n_keys, lt);
            private_insert<kScanWarps>(maybe_b && kb > tau, kb, L.cbuf, n_keys, lt);
        }
    }
}

// Single-chunk variant (one chunk per lane, 32 chunks per step): used for fp32 values, where a lane's two chunks
// (96 bytes of payload) would not fit the register budget twice over and the kernel is HBM-bound anyway.
template <int VT, bool ROUND, bool DIAG, bool SAMPLE>
__device__ __forceinline__ void process_window(const Chunk<VT> &cur, const uint32_t qs, const int lane, const uint32_t lt,
                                               float &carry, uint32_t &row, int &n_keys, const float tau_s, uint8_t *smem,
                                               CtaState *st, const ScanParams &p, const int b) {
    const bool is_tail = (cur.c.x & 0x8000u) != 0;
    const uint32_t T = __ballot_sync(0xffffffffu, is_tail);
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
        const bool maybe = is_tail && (s >= tau_s);
        if (__any_sync(0xffffffffu, maybe)) {
            const SmemLayout L = smem_layout(smem, p);
            const uint64_t key = make_key(s, rid);
            const uint64_t tau = *(volatile uint64_t *)&st->tau;
            private_insert<kScanWarps>(maybe && key > tau, key, L.cbuf, n_keys, lt);
        }
    }
}

template <int VT, int D, bool PAIR, bool ROUND, bool DIAG>
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

        // ---- stream this warp's part, 64 chunks (two per lane) per step.  The stream carries kStreamSlack windows
        // of slack past its end, so the prefetch ring never needs a bounds check: loads are unconditional,
        // addresses are base + immediate.  Parts are a whole number of 64-chunk steps (build_index.cu).
        constexpr int CPL = PAIR ? 2 : 1;       // chunks per lane per step
        constexpr int CPS = 32 * CPL;           // chunks per step
        constexpr int VS = (VT == 1) ? 2 : 1;   // uint4 of values per chunk
        const uint4 *cp = p.cols + (uint64_t)w_begin * 32ull + CPL * lane;
        const uint4 *vp = (const uint4 *)p.vals + ((uint64_t)w_begin * 32ull + CPL * lane) * VS;
        Chunk<VT> ra[D], rb[PAIR ? D : 1];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            load_chunk<VT>(ra[j], cp + j * CPS, vp + j * CPS * VS);
            if constexpr (PAIR) load_chunk<VT>(rb[j], cp + j * CPS + 1, vp + (j * CPS + 1) * VS);
        }
        float carry = 0.f;
        uint32_t row = row0;
        int n_keys = 0;   // phase A: keys in this warp's sample slice; phase B: keys in its private region
        int w = 0;
        const int nstep = PAIR ? (nwin >> 1) : nwin;
#define VS_PROCESS(SAMPLE, J)                                                                                      \
    {                                                                                                              \
        if constexpr (PAIR)                                                                                        \
            process_pair<VT, ROUND, DIAG, SAMPLE>(ra[J], rb[J], qs, lane, lt, carry, row, n_keys, tau_s, smem,     \
                                                  &st, p, b);                                                      \
        else                                                                                                       \
            process_window<VT, ROUND, DIAG, SAMPLE>(ra[J], qs, lane, lt, carry, row, n_keys, tau_s, smem, &st, p,  \
                                                    b);                                                            \
    }
#define VS_STEP(SAMPLE, J)                                                                                         \
    {                                                                                                              \
        VS_PROCESS(SAMPLE, J)                                                                                      \
        load_chunk<VT>(ra[J], cp + (D + J) * CPS, vp + (D + J) * CPS * VS);                                        \
        if constexpr (PAIR) load_chunk<VT>(rb[J], cp + (D + J) * CPS + 1, vp + ((D + J) * CPS + 1) * VS);          \
    }
#define VS_ADVANCE()                                                                                               \
    {                                                                                                              \
        cp += D * CPS;                                                                                             \
        vp += D * CPS * VS;                                                                                        \
        w += D;                                                                                                    \
    }
        // ---- phase A: sampling.  Every row's key goes straight into this warp's slice of cbuf (no threshold, no
        // lock) until the slice cannot take D more steps; then ONE CTA-wide select sets the threshold.
        float tau_s = -INFINITY;
        while (w + D <= nstep && n_keys + D * CPS <= kSampleRegion) {
#pragma unroll
            for (int j = 0; j < D; ++j) VS_STEP(true, j)
            VS_ADVANCE()
        }
        const SmemLayout L = smem_layout(smem, p);
        {
            for (int i = n_keys + lane; i < kSampleRegion; i += 32) L.cbuf[warp * kSampleRegion + i] = 0ull;
            __syncthreads();
            cta_sample_select<kScanThreads>(L.cbuf, kSampleKeys, p.k, L.hist, &st);
            n_keys = 0;
        }
        // ---- phase B: steady state.  Once per group of D steps: one 64-bit "gate" load (float threshold + join
        // epoch), and a join when somebody asked for a re-selection or this warp's region cannot take D more steps.
        uint32_t epoch = 0;
        while (w + D <= nstep) {
            const uint64_t gate = gate_load(&st);
            tau_s = gate_tau_score(gate);
#pragma unroll
            for (int j = 0; j < D; ++j) VS_STEP(false, j)
            VS_ADVANCE()
            join_if_needed<kScanThreads, kScanWarps>(gate, epoch, D * CPS, L.cbuf, n_keys, p.k, L.hist, &st);
        }
        tau_s = gate_tau_score(gate_load(&st));
#pragma unroll
        for (int j = 0; j < D - 1; ++j)
            if (w + j < nstep) VS_PROCESS(false, j)
        finish_streaming<kScanThreads, kScanWarps>(epoch, L.cbuf, n_keys, p.k, L.hist, &st);
#undef VS_STEP
#undef VS_PROCESS
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

template <int VT, int D, bool PAIR, bool ROUND, bool DIAG>
static int launch_scan_t(const vs_index *idx, const ScanParams &p, size_t smem, cudaStream_t st) {
    auto kern = scan_topk_kernel<VT, D, PAIR, ROUND, DIAG>;
    VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<idx->n_ctas, kScanThreads, smem, st>>>(p);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

template <int VT, int D, bool PAIR>
static int launch_scan_v(const vs_index *idx, const ScanParams &p, size_t smem, cudaStream_t st) {
    if (p.scores_out) return launch_scan_t<VT, D, PAIR, true, true>(idx, p, smem, st);  // diagnostic path
    if (p.score_round != VS_F32) return launch_scan_t<VT, D, PAIR, true, false>(idx, p, smem, st);
    return launch_scan_t<VT, D, PAIR, false, false>(idx, p, smem, st);
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
    static_assert(ScanGeom::kPrivate >= 3 * 64 + 32, "private region must take one group of steps");
    // <value type, steps in flight per warp, two chunks per lane?>; the placement at build uses the same mode
    if (idx->kind == 2) return launch_scan_v<0, 3, true>(idx, p, smem, st);
    if (idx->store_dtype == VS_F32) return launch_scan_v<1, 2, false>(idx, p, smem, st);
    if (idx->store_dtype == VS_F16) return launch_scan_v<2, 2, true>(idx, p, smem, st);
    return launch_scan_v<3, 2, true>(idx, p, smem, st);
}

}  // namespace vs
