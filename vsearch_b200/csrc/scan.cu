// scan.cu -- K1 (valued CSR) / K2 (binary bag-of-token) passage-major scan with the top-k
// fused in (K5), for sm_100a.  Replaces `torch.matmul(q, vector.t())` + `scores.topk(k)`
// (upstream src/ir/retriever/index.py:91-92) for SparseIndex / BoTIndex.
//
// One persistent CTA per SM, kScanWarps (20) warps.  Per query ("pass"):
//   1. the dense fp32 query vector (V+1 slots, slot V = 0 for padding) is staged in shared memory: 1-D bulk TMA copies of
//      the prepared row (cp.async.bulk + mbarrier), or zero + scatter of the caller's (token, weight) list;
//   2. every warp streams ITS part of the WS index (index.cuh) in steps of 32*C chunks, C consecutive (logical)
//      chunks per lane stored transposed so every load is a coalesced 512-byte window -- C = 8 for the binary index
//      (scan_bin_kernel: a whole step in flight per warp; the last step of a pass wraps around and loads step 0 for the
//      next query), 2 for 16-bit values, 1 for fp32 values (register ring of D steps); loads never depend on row
//      pointers, the row-end ("tail") flag of a chunk rides in bit 15 of its first entry; each lane gathers q[col] for
//      its entries from shared memory;
//   3. ONE segmented warp scan per step turns lane partials into row scores -- with 8 chunks per lane a row of ~15
//      chunks spans 2-3 lanes, so the scan needs 1-2 shuffle levels per 256 chunks (the levels nobody needs are
//      skipped) instead of 5 per 64;
//   4. thresholds.  Binary kernel: every candidate row is counted in a two-level score histogram (topk.cuh); a warp whose
//      private region fills asks the histogram for the k-th bucket's bound (no barrier), drops its keys below it and
//      goes on.  Valued kernels (K1): the first rows of a pass are sampled and ONE CTA-wide radix select sets the
//      threshold.  Behind both: a float pre-filter rejects almost every row, keys go to the warp's PRIVATE region with
//      plain stores, and the exact join protocol (one CTA-wide re-selection per epoch, no locks, no per-row atomics)
//      covers what a score bucket cannot separate;
//   5. end of pass: the binary kernel hands every key at or above its final bound (k .. 2k + 64) to the merge kernel;
//      the valued kernels select their exact top-k and write k keys.
// The [B, N] score matrix is never written.  A second tiny kernel (merge.cu) merges the per-CTA lists.
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
    int kout;              // keys per (query, CTA) list handed to the merge kernel (binary scan: > k, see scan_kout)
    int cap;               // CTA candidate buffer entries (>= k + kStage)
    int vpad;              // floats per prepared query (multiple of 4, > V)
    int score_round;
    uint32_t sentinel;     // V | V << 16
    unsigned long long *prof;  // diagnostic: per (CTA, pass) phase timestamps (globaltimer ns), or nullptr
    // sparse queries (vs_search_sparse): CSR-style lists instead of prepared dense rows; query b of this launch is
    // entries [q_ptr[b0 + b], q_ptr[b0 + b + 1]) of q_tok / q_w
    const void *q_ptr;
    int ptr_dtype;
    const int32_t *q_tok;
    const float *q_w;
    int64_t b0;
    int64_t n_cols;
    const int *use_inv;        // device flag (auto mode): != 0 -> the inverted lists serve this chunk, this kernel exits
};
constexpr int kProfSlots = 8;
__device__ __forceinline__ void prof_mark(const ScanParams &p, int b, int slot) {
    if (p.prof != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.prof[((size_t)blockIdx.x * p.B + b) * kProfSlots + slot] = t;
    }
}

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

// Stage query b into shared memory (fp32 [vpad], slot V.. = 0).  Dense input: 1-D bulk TMA copies of the prepared row,
// completion on the CTA's mbarrier.  Sparse input: zero the vector, scatter the (token, weight) list (weights rounded
// like the prepared rows; duplicate tokens add up, tokens outside [0, V) are ignored).  Called by all threads; returns
// with the vector visible to the whole CTA.  `between` runs after the copies are issued and before anybody waits.
template <int NT, typename F>
__device__ __forceinline__ void stage_query(uint8_t *smem, CtaState *st, const ScanParams &p, const int b, uint32_t &phase,
                                            F between) {
    const int tid = threadIdx.x;
    const uint32_t q_bytes = (uint32_t)p.vpad * 4u;
    if (p.q_ptr == nullptr) {
        if (tid == 0) {
            cta_state_reset(st);
            fence_proxy_async();  // earlier generic-proxy reads of the vector are ordered before the async writes
            mbar_arrive_expect_tx(&st->mbar, q_bytes);
            const uint8_t *src = reinterpret_cast<const uint8_t *>(p.q + (size_t)b * p.vpad);
            for (uint32_t off = 0; off < q_bytes; off += 16384u) {
                uint32_t n = min(16384u, q_bytes - off);
                bulk_g2s(smem + off, src + off, n, &st->mbar);
            }
        }
        between();
        __syncthreads();          // cnt/tau reset visible
        mbar_wait(&st->mbar, phase);
        phase ^= 1u;
    } else {
        if (tid == 0) cta_state_reset(st);
        float4 *q4 = reinterpret_cast<float4 *>(smem);
        for (int i = tid; i < (p.vpad >> 2); i += NT) q4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        between();
        __syncthreads();
        const int64_t lo = p.ptr_dtype == VS_I32 ? (int64_t)((const int32_t *)p.q_ptr)[p.b0 + b] : ((const int64_t *)p.q_ptr)[p.b0 + b];
        const int64_t hi = p.ptr_dtype == VS_I32 ? (int64_t)((const int32_t *)p.q_ptr)[p.b0 + b + 1] : ((const int64_t *)p.q_ptr)[p.b0 + b + 1];
        float *qf = reinterpret_cast<float *>(smem);
        for (int64_t j = lo + tid; j < hi; j += NT) {
            const int32_t t = p.q_tok[j];
            if (t >= 0 && (int64_t)t < p.n_cols) atomicAdd(&qf[t], round_score(p.q_w[j], p.score_round));
        }
        __syncthreads();
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

    if (tid == 0) {
        mbar_init(&st.mbar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (p.use_inv != nullptr && *p.use_inv != 0) return;   // auto mode chose the inverted lists for this chunk
    uint32_t phase = 0;
    for (int b = 0; b < p.B; ++b) {
        // ---- stage the query vector
        stage_query<kScanThreads>(smem, &st, p, b, phase, [] {});

        // ---- stream this warp's part, 64 chunks (two per lane) per step.  The stream carries kStreamSlack windows
        // of slack past its end, so the prefetch ring never needs a bounds check: loads are unconditional,
        // addresses are base + immediate.  Parts are a whole number of 64-chunk steps (build_index.cu).
        constexpr int CPL = PAIR ? 2 : 1;       // chunks per lane per step
        constexpr int CPS = 32 * CPL;           // chunks per step
        constexpr int VS = (VT == 1) ? 2 : 1;   // uint4 of values per chunk
        // transposed steps (index.cuh): the lane's chunk #i of a step sits at step base + i*32 + lane
        const uint4 *cp = p.cols + (uint64_t)w_begin * 32ull + lane;
        const uint4 *vp = (const uint4 *)p.vals + ((uint64_t)w_begin * 32ull + lane) * VS;
        Chunk<VT> ra[D], rb[PAIR ? D : 1];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            load_chunk<VT>(ra[j], cp + j * CPS, vp + j * CPS * VS);
            if constexpr (PAIR) load_chunk<VT>(rb[j], cp + j * CPS + 32, vp + (j * CPS + 32) * VS);
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
        if constexpr (PAIR) load_chunk<VT>(rb[J], cp + (D + J) * CPS + 32, vp + ((D + J) * CPS + 32) * VS);        \
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
            uint64_t *out = p.cand + ((size_t)b * gridDim.x + blockIdx.x) * (size_t)p.kout;
            cta_write_topk<kScanThreads, kScanWarps>(L.cbuf, n_keys, p.k, L.hist, &st, out);
            for (int i = p.k + tid; i < p.kout; i += kScanThreads) out[i] = 0ull;   // lists are kout long for every kernel family
        }
        __syncthreads();  // everyone is done with qs / cbuf before the next pass overwrites them
    }
}

// =====================================================================================================================
// K2: binary bag-of-token scan, 8 chunks per lane per step (256 chunks = 4 KB of column ids per warp step).
// Lane l owns the LOGICAL chunks 8l .. 8l+7 of the step (stored transposed: load #i of the warp is one coalesced
// 512-byte window), so a row of ~15 chunks spans 2-3 lanes: one segmented warp scan per 256 chunks, and the levels
// of that scan which no lane needs are skipped (usually 1-2 shuffles instead of 5).  The whole next step is in flight
// while the current one is processed: chunk i's registers are refilled right after its gathers.
// Per lane and step:  d[i] = sum of q[col] over chunk i;  the lane's chunks are walked in order -- the part up to
// its FIRST row end is `head` (completed by what the lanes to the left hand over), the part after its LAST row end
// is `v` (handed to the right); rows that start AND end inside the lane (shorter than 8 chunks: rare) are scored
// locally in a warp-uniform side loop.
constexpr int kBinC = 8;

struct StepFlags {
    uint32_t F;       // lanes with at least one row end
    int nend;         // row ends in this lane
    int rank;         // row ends in the lanes before this one
    int total;        // row ends in the step
    bool multi;       // some lane holds more than one row end
};

__device__ __forceinline__ StepFlags step_flags(const uint4 (&r)[kBinC], const uint32_t lt) {
    StepFlags f;
    uint32_t fm = 0;
#pragma unroll
    for (int i = 0; i < kBinC; ++i) fm |= ((r[i].x >> 15) & 1u) << i;
    f.nend = __popc(fm);
    f.F = __ballot_sync(0xffffffffu, fm != 0);
    f.multi = __any_sync(0xffffffffu, f.nend > 1);
    if (!f.multi) {
        f.rank = __popc(f.F & lt);
        f.total = __popc(f.F);
    } else {
        f.rank = 0;
        f.total = 0;
#pragma unroll
        for (int bit = 0; bit < 4; ++bit) {   // nend <= 8
            const uint32_t m = __ballot_sync(0xffffffffu, (f.nend >> bit) & 1);
            f.rank += __popc(m & lt) << bit;
            f.total += __popc(m) << bit;
        }
    }
    return f;
}

__device__ __forceinline__ float bin_chunk_dot(const uint4 &c, const uint32_t qs) {
    float g[8];
    g[0] = lds_f32(qs + ((c.x & 0x7fffu) << 2));   // entry 0 carries the row-end flag in bit 15
    g[1] = lds_f32(qs + (__byte_perm(c.x, 0, 0x4432) << 2));
    g[2] = lds_f32(qs + (__byte_perm(c.y, 0, 0x4410) << 2));
    g[3] = lds_f32(qs + (__byte_perm(c.y, 0, 0x4432) << 2));
    g[4] = lds_f32(qs + (__byte_perm(c.z, 0, 0x4410) << 2));
    g[5] = lds_f32(qs + (__byte_perm(c.z, 0, 0x4432) << 2));
    g[6] = lds_f32(qs + (__byte_perm(c.w, 0, 0x4410) << 2));
    g[7] = lds_f32(qs + (__byte_perm(c.w, 0, 0x4432) << 2));
    return ((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]));
}

struct BinSmem {
    uint64_t *cbuf;      // kCapMax keys: shared top-k set + per-warp private regions (topk.cuh)
    uint32_t *hist;      // 256 words: radix-select scratch of the exact fallback
    uint32_t *fine;      // kHistFine
    uint32_t *coarse;    // kHistCoarse
};
__device__ __forceinline__ BinSmem bin_smem(uint8_t *smem, const ScanParams &p) {
    BinSmem L;
    L.cbuf = reinterpret_cast<uint64_t *>(smem + (size_t)p.vpad * 4);
    L.hist = reinterpret_cast<uint32_t *>(L.cbuf + kCapMax);
    L.fine = L.hist + 256;
    L.coarse = L.fine + kHistFine;
    return L;
}

// one row score -> (DIAG) the score matrix; float pre-filter, exact 64-bit test, count in the score histogram, append
// to the warp's private region
template <bool ROUND, bool DIAG>
__device__ __forceinline__ void emit_row(const bool valid, float s, const uint32_t rid, const float tau_s, int &n_keys,
                                         const uint32_t lt, uint8_t *smem, CtaState *st, const ScanParams &p, const int b) {
    if constexpr (ROUND) s = round_score(s, p.score_round);
    if constexpr (DIAG) {
        if (valid) p.scores_out[(size_t)b * p.n_rows + rid] = s + 0.0f;
    }
    const bool maybe = valid && (s >= tau_s);
    if (__any_sync(0xffffffffu, maybe)) {
        const BinSmem L = bin_smem(smem, p);
        const uint64_t tau = *(volatile uint64_t *)&st->tau;   // exact k-th key of the last CTA-wide join (0: none yet)
        const uint64_t key = make_key(s, rid);
        const bool ins = maybe && key > tau;
        hist_count(L.coarse, L.fine, ins, (uint32_t)(key >> 32));
        private_insert<kScanWarps>(ins, key, L.cbuf, n_keys, lt);
    }
}

// MULTI: some lane of the step holds more than one row end (rows shorter than 8 chunks).  A row that starts AND ends
// inside a lane is scored on the spot (warp-collective emit per chunk); the common case compiles without it.
template <bool ROUND, bool DIAG, bool MULTI>
__device__ __forceinline__ void process_step8(uint4 (&r)[kBinC], const uint4 *next, const StepFlags &f, const uint32_t qs,
                                              const int lane, const uint32_t lt, float &carry, uint32_t &row, int &n_keys,
                                              const float tau_s, uint8_t *smem, CtaState *st, const ScanParams &p,
                                              const int b) {
    // walk the lane's chunks in order
    float run = 0.f, head = 0.f;
    bool seen = false;
    [[maybe_unused]] uint32_t ord = 0;   // MULTI: row ends of this lane so far
#pragma unroll
    for (int i = 0; i < kBinC; ++i) {
        const bool end_i = (r[i].x & 0x8000u) != 0;   // bit 15 of the chunk's first entry: last chunk of its row
        run += bin_chunk_dot(r[i], qs);
        r[i] = ldg_stream(next + i * 32);   // same registers: chunk i of the next step
        if constexpr (MULTI) {
            emit_row<ROUND, DIAG>(end_i && seen, run, row + (uint32_t)f.rank + ord, tau_s, n_keys, lt, smem, st, p, b);
            ord += end_i ? 1u : 0u;
        }
        if (end_i) {
            if (!seen) head = run;
            seen = true;
            run = 0.f;
        }
    }
    float v = run;   // what this lane hands to its right neighbour: the part after its last row end
    const uint32_t F = f.F;
    if (lane == 0 && !(F & 1u)) v += carry;  // lane 0 passes the carry through unless a row ends inside it
    // inclusive segmented scan of v; a flagged lane RESTARTS the sum.  reach = lanes back to the nearest flagged lane.
    const uint32_t upto = F & (lt | (1u << lane));
    const int reach = upto ? lane - (31 - __clz(upto)) : lane;
    // level d is needed iff some lane >= d has reach >= d, i.e. ~F holds a run of d ones above bit 0 (warp-uniform)
    const uint32_t z1 = (~F) >> 1;
    if (z1) {
        float o = __shfl_up_sync(0xffffffffu, v, 1);
        if (reach >= 1) v += o;
        const uint32_t z2 = z1 & (z1 >> 1);
        if (z2) {
            o = __shfl_up_sync(0xffffffffu, v, 2);
            if (reach >= 2) v += o;
            const uint32_t z4 = z2 & (z2 >> 2);
            if (z4) {
                o = __shfl_up_sync(0xffffffffu, v, 4);
                if (reach >= 4) v += o;
                const uint32_t z8 = z4 & (z4 >> 4);
                if (z8) {
                    o = __shfl_up_sync(0xffffffffu, v, 8);
                    if (reach >= 8) v += o;
                    if (z8 & (z8 >> 8)) {
                        o = __shfl_up_sync(0xffffffffu, v, 16);
                        if (reach >= 16) v += o;
                    }
                }
            }
        }
    }
    const float rot = __shfl_sync(0xffffffffu, v, (lane + 31) & 31);  // left neighbour's running sum; lane 0 gets lane 31's
    const float incoming = (lane == 0) ? carry : rot;
    carry = rot;  // only lane 0's copy is used: the open row's partial sum at the end of this step
    // the row ending at this lane's first row end
    emit_row<ROUND, DIAG>(seen, incoming + head, row + (uint32_t)f.rank, tau_s, n_keys, lt, smem, st, p, b);
    row += (uint32_t)f.total;
}

template <bool ROUND, bool DIAG>
__global__ void __launch_bounds__(kScanThreads, 1) scan_bin_kernel(const ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ CtaState st;
    const uint32_t qs = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    const int part = blockIdx.x * kScanWarps + warp;
    const uint32_t w_begin = p.part_win_begin[part];
    const int nstep = (int)(p.part_win_begin[part + 1] - w_begin) / kBinC;   // parts are whole steps (build_index.cu)
    const uint32_t row0 = p.part_row_begin[part];
    constexpr int SC = 32 * kBinC;   // chunks per step

    if (tid == 0) {
        mbar_init(&st.mbar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (p.use_inv != nullptr && *p.use_inv != 0) return;   // auto mode chose the inverted lists for this chunk
    uint32_t phase = 0;
    const uint4 *cp = p.cols + (uint64_t)w_begin * 32ull + lane;
    uint4 r[kBinC];   // the step being processed; refilled chunk by chunk with the next step's (process_step8)
    for (int b = 0; b < p.B; ++b) {
        prof_mark(p, b, 0);
        // the first step's loads go out before anybody waits for the query vector
        const BinSmem L = bin_smem(smem, p);
        stage_query<kScanThreads>(smem, &st, p, b, phase, [&] {
            // (passes after the first find step 0 already in r[]: the last step of the previous pass wrapped around,
            // so the loads were in flight under that pass's final write and this query's staging)
            if (b == 0 || nstep == 0) {
#pragma unroll
                for (int i = 0; i < kBinC; ++i) r[i] = ldg_stream(cp + i * 32);
            }
            for (int i = tid; i < kHistFine + kHistCoarse; i += kScanThreads) L.fine[i] = 0u;   // coarse follows fine
        });
        prof_mark(p, b, 1);

        float carry = 0.f;
        uint32_t row = row0;
        int n_keys = 0;   // keys in this warp's private region
        uint32_t epoch = 0;
        // ---- stream.  No sampling phase: while the pre-filter is still open (-inf) every row is a candidate and is
        // counted in the score histogram; a warp whose region cannot take the next step's rows asks the histogram for
        // the threshold (warp_refresh: no barrier), and only if that does not make room -- equal scores en masse,
        // adversarial row order -- the CTA falls back to the exact re-selection (join protocol, topk.cuh).
        for (int s = 0; s < nstep; ++s) {
            const uint64_t gate = gate_load(&st);
            const StepFlags f = step_flags(r, lt);
            if (gate_epoch(gate) != epoch || n_keys + f.total > ScanGeom::kPrivate) {
                if (gate_epoch(gate) == epoch) warp_refresh<kScanWarps>(L.coarse, L.fine, p.k, L.cbuf, n_keys, &st);
                join_if_needed<kScanThreads, kScanWarps>(gate, epoch, f.total, L.cbuf, n_keys, p.k, L.hist, &st);
            }
            const float tau_s = gate_tau_score(gate_load(&st));
            const uint4 *nx = s + 1 < nstep ? cp + (size_t)(s + 1) * SC : cp;   // the last step wraps around: step 0, for the next query
            if (f.multi)
                process_step8<ROUND, DIAG, true>(r, nx, f, qs, lane, lt, carry, row, n_keys, tau_s, smem, &st, p, b);
            else
                process_step8<ROUND, DIAG, false>(r, nx, f, qs, lane, lt, carry, row, n_keys, tau_s, smem, &st, p, b);
        }
        if (warp == 0) prof_mark(p, b, 4);
        finish_streaming<kScanThreads, kScanWarps>(epoch, L.cbuf, n_keys, p.k, L.hist, &st);
        prof_mark(p, b, 5);
        // ---- end of pass: everything at or above the final histogram threshold goes to the merge kernel (between k
        // and kout keys); more than kout (a bucket full of equal scores) -> the exact CTA-wide select
        if (tid == 0) st.scratch = 0;
        __syncthreads();
        {
            uint64_t *out = p.cand + ((size_t)b * gridDim.x + blockIdx.x) * (size_t)p.kout;
            const uint64_t tmin = (uint64_t)hist_threshold(L.coarse, L.fine, p.k) << 32;
            const int n_sh = (int)st.cnt;
            const uint64_t *priv = L.cbuf + kSharedKeys + warp * ScanGeom::kPrivate;
            for (int i = tid; i < n_sh; i += kScanThreads) {
                const uint64_t x = L.cbuf[i];
                if (x != 0ull && x >= tmin) { const uint32_t o = atomicAdd(&st.scratch, 1u); if (o < (uint32_t)p.kout) out[o] = x; }
            }
            for (int i = lane; i < n_keys; i += 32) {
                const uint64_t x = priv[i];
                if (x >= tmin) { const uint32_t o = atomicAdd(&st.scratch, 1u); if (o < (uint32_t)p.kout) out[o] = x; }
            }
            __syncthreads();
            const uint32_t total = *(volatile uint32_t *)&st.scratch;
            if (total > (uint32_t)p.kout) {   // uniform
                __syncthreads();
                cta_write_topk<kScanThreads, kScanWarps>(L.cbuf, n_keys, p.k, L.hist, &st, out);
                for (int i = p.k + tid; i < p.kout; i += kScanThreads) out[i] = 0ull;
            } else {
                for (int i = (int)total + tid; i < p.kout; i += kScanThreads) out[i] = 0ull;
            }
        }
        __syncthreads();  // everyone is done with qs / cbuf before the next pass overwrites them
        prof_mark(p, b, 6);
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
static size_t scan_bin_smem_bytes(int vpad) {
    return (size_t)vpad * 4 + (size_t)kCapMax * 8 + 256 * 4 + (size_t)(kHistFine + kHistCoarse) * 4;
}
// Length of the per-(query, CTA) candidate lists.  The binary scan and the inverted-list kernel hand over every key at
// or above their final histogram threshold -- between k and, for smooth score distributions, about 2k of them --
// instead of paying a CTA-wide exact select per pass; the valued scan kernels write exactly k and pad.
int scan_kout(const vs_index *idx, int k) { (void)idx; return 2 * k + 64; }

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

// Queries: d_qprep [B, vpad] fp32 prepared rows, or (sq != nullptr) CSR-style lists.  d_cand: [B, n_ctas, scan_kout]
// keys; d_scores_out optional [B, N]; d_use_inv optional device flag (auto mode).
int launch_scan(const vs_index *idx, const float *d_qprep, const SparseQueries *sq, int vpad, int64_t B, int k,
                int score_round, uint64_t *d_cand, float *d_scores_out, const int *d_use_inv, cudaStream_t st) {
    ScanParams p;
    p.q_ptr = sq ? sq->ptr : nullptr;
    p.ptr_dtype = sq ? sq->ptr_dtype : VS_I64;
    p.q_tok = sq ? sq->tok : nullptr;
    p.q_w = sq ? sq->w : nullptr;
    p.b0 = sq ? sq->b0 : 0;
    p.n_cols = idx->n_cols;
    p.use_inv = d_use_inv;
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
    p.kout = scan_kout(idx, k);
    p.cap = scan_cap_for_k(k);
    p.vpad = vpad;
    p.score_round = score_round;
    p.sentinel = (uint32_t)idx->n_cols | ((uint32_t)idx->n_cols << 16);
    p.prof = idx->scan_prof;
    size_t smem = scan_smem_bytes(vpad, p.cap);
    VS_REQUIRE(smem <= 227 * 1024, VS_ERR_UNSUPPORTED,
               "n_cols=%lld needs %zu bytes of shared memory (> 227 KB): vocabulary too large for the scan kernel",
               (long long)idx->n_cols, smem);
    static_assert(kStreamSlack >= 8, "prefetch ring deeper than the stream slack");
    static_assert(ScanGeom::kPrivate >= 3 * 64 + 32, "private region must take one group of steps");
    static_assert(ScanGeom::kPrivate >= 32 * kBinC, "private region must take the rows of one binary step");
    if (idx->kind == 2) {   // binary: 8 chunks per lane per step (idx->cpl_shift == 3)
        smem = scan_bin_smem_bytes(vpad);
        VS_REQUIRE(smem <= 227 * 1024, VS_ERR_UNSUPPORTED,
                   "n_cols=%lld needs %zu bytes of shared memory (> 227 KB): vocabulary too large for the scan kernel",
                   (long long)idx->n_cols, smem);
        auto launch = [&](auto kern) -> int {
            VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<idx->n_ctas, kScanThreads, smem, st>>>(p);
            VS_CUDA(cudaGetLastError());
            return VS_OK;
        };
        if (p.scores_out) return launch(scan_bin_kernel<true, true>);   // diagnostic path
        if (p.score_round != VS_F32) return launch(scan_bin_kernel<true, false>);
        return launch(scan_bin_kernel<false, false>);
    }
    // <value type, steps in flight per warp, two chunks per lane?>; the placement at build uses the same mode
    if (idx->store_dtype == VS_F32) return launch_scan_v<1, 2, false>(idx, p, smem, st);
    if (idx->store_dtype == VS_F16) return launch_scan_v<2, 2, true>(idx, p, smem, st);
    return launch_scan_v<3, 2, true>(idx, p, smem, st);
}

}  // namespace vs
