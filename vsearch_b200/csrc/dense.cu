// dense.cu -- K4: dense index scoring on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), top-k fused into the
// epilogue.  Replaces `torch.matmul(q, vector.t())` + `scores.topk(k)` (upstream src/ir/retriever/index.py:91-92)
// for the strided `Index.vector` [N, D] (bf16 or fp16 storage, fp32 accumulate).
//
//   S[b, n] = sum_d Q[b, d] * X[n, d]      A = Q tile [128 queries, 64] , B = X tile [256 passages, 64], both K-major
//
// Persistent, warp-specialised (320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of the A / B k-blocks (128-byte swizzle) into a
//               shared-memory ring, completion on mbarriers
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.kind::f16 (K = 16, x4 per k-block) into one of two
//               256-column TMEM accumulators; tcgen05.commit frees the smem stage / signals the epilogue
//   warps 2-9   epilogue: warp w reads TMEM lane quarter w % 4 (thread = query row), columns 0..127 (warps 2-5) or
//               128..255 (warps 6-9), in ONE tcgen05.ld round trip per tile; S is never written:
//               mode 0 (sample)    emit every score of the tile as a rank key (small sample prefix of the index)
//               mode 1 (filter)    compare against the per-query threshold; candidates are appended to the
//                                  query's list (one global atomic each; they are rare)
// Two variants: dense_topk_pair_kernel (default; cta_group::2, a cluster of two CTAs shares one M=256 x N=256 UMMA,
// static item order) and dense_topk_kernel (cta_group::1, M=128 x N=256, dynamic tile scheduler; used when there
// is a single 128-query tile).  DESIGN.md section 4 records what bounds the kernel and how that was measured.
//
// Host side (search_dense): threshold from an exact top-k of a sample prefix -> filtered sweeps over the rest of the
// index (the threshold is tightened once after the first ~1M rows) -> exact top-k of the survivors with the K6 merge
// kernel.  The whole call is enqueued without reading anything back; a candidate list that overflows (adversarial
// ordering) raises a status word that is polled once at the end, and such a call is redone through the checked path
// (counts read after every sweep, an overflowed sweep repeated with the k-th best of what it stored).  Row-sharded
// (search_dense_step): the same sweeps in three steps, the ranks' [B, k] keys pooled after each, so every rank sweeps
// 1/W of the sample and filters with the thresholds of the whole index.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <vector>

#include "index.cuh"

namespace vs {

constexpr int kBM = 128, kBN = 256, kBK = 64, kStages = 4;
constexpr int kEpiWarps = 8;                       // two per TMEM lane quarter: each takes half of the tile's 256 columns
constexpr int kDenseThreads = 64 + kEpiWarps * 32;  // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
constexpr uint32_t kStageBytesA = kBM * kBK * 2, kStageBytesB = kBN * kBK * 2;
constexpr uint32_t kStageBytes = kStageBytesA + kStageBytesB;  // 48 KB
constexpr int kTmemCols = 512;

struct DenseArgs {
    int n_tiles_m, n_tiles_n, k_blocks;
    int64_t n_rows;        // real passages (rows >= n_rows of the padded matrix are ignored)
    int64_t n_queries;     // real queries
    int64_t row_offset;    // first passage row of this sweep (sample sweeps start at 0)
    int mode;              // 0 sample: write keys [n_queries, sample_ld]; 1 filter
    int score_round;
    uint32_t idesc;
    uint64_t *sample_keys; int64_t sample_ld;
    const uint64_t *tau;   // [n_queries] threshold keys (mode 1)
    uint64_t *cand; uint32_t *cand_cnt; int64_t cand_cap;  // [n_queries, cand_cap], [n_queries]
    unsigned long long *work_counter;   // dynamic tile scheduler (zeroed before every launch)
    int dbg;               // timing experiments only (VSEARCH_B200_DENSE_DBG): 1 skip epilogue, 2 skip MMA, 4 skip TMA, 8 skip candidate path
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// tcgen05.ld without the wait (the registers must not be read before tc_wait_ld + reg_fence32)
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ties every later use of r[] to a point after the preceding tc_wait_ld (volatile asms keep their order)
__device__ __forceinline__ void reg_fence32(uint32_t (&r)[32]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
    asm volatile("" : "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
    asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]));
    asm volatile("" : "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}

// A tcgen05.ld queues behind every tcgen05.mma already issued on the SM (the MMA warp keeps the ring of k-blocks full:
// ~1 us of queued tensor work), so the epilogue's cost is the NUMBER of ld -> wait round trips, not the bytes: a serial
// ld/wait per 32 columns made the epilogue (16 us per tile), not TMA + MMA (6 us), the bottleneck (measured with the
// epilogue switched off, profiles/README.md).  The filtered sweep therefore pulls 128 columns per wait.
//
// Candidates are rare (k / rows-seen per score), so the filter is two-level and stays in registers: a lane whose
// 32-column chunk maximum reaches the (conservative, un-rounded) threshold builds the bit mask of its qualifying
// columns -- one compare per score -- and hands each one to an out-of-line exact test (rounding, key compare, one
// global atomic, one 8-byte store).  The epilogue has to be SMALL as well as short: every variant that inlined the key
// handling per column (2-6 K instructions) left the warps waiting on instruction fetch, a local-memory copy of the
// scores has no L1 to live in next to 214 KB of shared memory, and re-reading a hot chunk from TMEM costs another
// 1-2 us round trip behind the queued MMAs (all measured, profiles/README.md).
__device__ __noinline__ void dense_candidate(const float raw, const int64_t row, const int64_t n_rows, const int score_round,
                                             const uint64_t tau, uint32_t *cnt_q, uint64_t *dst, const uint32_t cap) {
    if (row >= n_rows) return;
    const uint64_t key = make_key(round_score(raw, score_round), (uint32_t)row);
    if (key < tau) return;
    // Past the end of the list the second half is reused as a ring: after an overflow the list then holds the first AND
    // the last cap/2 candidates of the sweep, so the retry threshold is tight whichever way the rows are ordered.
    const uint32_t pos = atomicAdd(cnt_q, 1u);
    dst[pos < cap ? pos : (cap >> 1) + (pos & ((cap >> 1) - 1u))] = key;
}

__device__ __forceinline__ void scan32(const uint32_t (&x)[32], const int64_t nb, const DenseArgs &a, const float tau_raw,
                                       const uint64_t tau, uint32_t *cnt_q, uint64_t *dst) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) m |= (__uint_as_float(x[j]) >= tau_raw) ? (1u << j) : 0u;
#pragma unroll 1
    while (m) {
        const uint32_t bit = m & (0u - m);
        m ^= bit;
        uint32_t v = x[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) v = (bit & (1u << j)) ? x[j] : v;   // dynamic register read as a select chain
        dense_candidate(__uint_as_float(v), nb + (__ffs(bit) - 1), a.n_rows, a.score_round, tau, cnt_q, dst, (uint32_t)a.cand_cap);
    }
}

__device__ __forceinline__ float max32(const uint32_t (&x)[32]) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(x[j]));
    return mx;
}

// ---- epilogue of one accumulator tile: thread = query row `q`, columns = passages n0 .. n0 + kBN ----------------
// `taddr` / `n0` already point at this warp's 128-column half of the tile
__device__ __forceinline__ void epilogue_tile(const DenseArgs &a, const uint32_t taddr, const int64_t q, const int64_t n0,
                                              const float *s_tau) {
    const bool q_ok = q < a.n_queries;
    if (a.mode == 0) {
        // ---- sample sweep: every score of the tile becomes a rank key
#pragma unroll 1
        for (int c = 0; c < kBN / 64; ++c) {
            uint32_t r[32];
            tc_ld32(taddr + c * 32, r);
            const int64_t nb = n0 + c * 32;
            if (q_ok) {
                uint64_t *dst = a.sample_keys + q * a.sample_ld + (nb - a.row_offset);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int64_t n = nb + j;
                    const float s = round_score(__uint_as_float(r[j]), a.score_round);
                    dst[j] = (n < a.n_rows) ? make_key(s, (uint32_t)n) : 0ull;
                }
            }
        }
    } else {
        // ---- filtered sweep: 128 accumulator columns per tcgen05.ld round trip; per 32-column chunk a float pre-filter
        // against the threshold score staged in shared memory, the rare chunk with a candidate goes out of line.
        const float tau_s = q_ok ? s_tau[q] : INFINITY;
        const uint64_t tau = q_ok ? a.tau[q] : ~0ull;
        uint64_t *dst = a.cand + q * a.cand_cap;
        uint32_t *cnt_q = a.cand_cnt + q;
        // scores are rounded to the index dtype before they are ranked; raw >= tau_raw is implied by round(raw) >= tau_s
        // (bf16 / fp16 rounding moves a value by < 2^-8 relative), the exact test happens on the few that pass
        const float tau_raw = a.score_round != VS_F32 ? tau_s - fabsf(tau_s) * 0.0078125f - 1e-30f : tau_s;
        uint32_t ra[32], rb[32], rc[32], rd[32];
        {   // one round trip for the warp's 128 columns
            const uint32_t th = taddr;
            tc_ld32_issue(th, ra); tc_ld32_issue(th + 32, rb); tc_ld32_issue(th + 64, rc); tc_ld32_issue(th + 96, rd);
            tc_wait_ld(); reg_fence32(ra); reg_fence32(rb); reg_fence32(rc); reg_fence32(rd);
            const float m0 = max32(ra), m1 = max32(rb), m2 = max32(rc), m3 = max32(rd);
            if (!(a.dbg & 8)) {
                if (m0 >= tau_raw) scan32(ra, n0, a, tau_raw, tau, cnt_q, dst);
                if (m1 >= tau_raw) scan32(rb, n0 + 32, a, tau_raw, tau, cnt_q, dst);
                if (m2 >= tau_raw) scan32(rc, n0 + 64, a, tau_raw, tau, cnt_q, dst);
                if (m3 >= tau_raw) scan32(rd, n0 + 96, a, tau_raw, tau, cnt_q, dst);
            }
        }
        static_assert(kBN == 256 && kEpiWarps == 8, "each epilogue warp owns 128 columns");
    }
}

__global__ void __launch_bounds__(kDenseThreads, 1)
dense_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x, const DenseArgs a) {
    extern __shared__ __align__(1024) uint8_t dsmem[];
    // dynamic smem: [stages x (A 16 KB | B 32 KB)] [barriers] ; 1024-byte alignment is required by SWIZZLE_128B
    uint8_t *base = reinterpret_cast<uint8_t *>(((uintptr_t)dsmem + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(base);
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + kStages * kStageBytes);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + kStages * 8;
    const uint32_t bar_tfull = bar_empty + kStages * 8, bar_tempty = bar_tfull + 2 * 8;
    const uint32_t bar_sfull = bar_tempty + 2 * 8, bar_sempty = bar_sfull + 2 * 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 8);
    volatile long long *sched_slot = reinterpret_cast<volatile long long *>(bars + 2 * kStages + 10);   // [2] work items
    float *s_tau = reinterpret_cast<float *>(bars + 2 * kStages + 12);   // [n_queries <= 4096] threshold scores (mode 1)
    if (a.mode == 1)
        for (int64_t i = threadIdx.x; i < a.n_queries; i += kDenseThreads) {
            const uint64_t t = a.tau[i];
            s_tau[i] = t ? key_score(t) : -INFINITY;
        }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_x) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(bars + s, 1); mbar_init(bars + kStages + s, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bars + 2 * kStages + i, 1);          // tfull: one tcgen05.commit
            mbar_init(bars + 2 * kStages + 2 + i, kEpiWarps);      // tempty: the epilogue warps
            mbar_init(bars + 2 * kStages + 4 + i, 1);      // sfull: the scheduler (producer thread)
            mbar_init(bars + 2 * kStages + 6 + i, 1 + kEpiWarps);      // sempty: MMA thread + the epilogue warps
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // work item = (passage tile, query tile), query tile fastest
    const long long n_work = (int64_t)a.n_tiles_n * a.n_tiles_m;

    if (warp == 0) {
        // ================= TMA producer + tile scheduler =================
        // Work items (passage tile, query tile; query tile fastest) are handed out by a global atomic counter, so
        // the CTAs that run concurrently always work on neighbouring items: the X tile a CTA needs is being read
        // by its neighbours right now and comes from L2 (a static round-robin drifts apart and re-reads X from HBM
        // ~9x: measured, profiles/README.md).  The item is published to the MMA / epilogue warps through a 2-slot ring.
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int ss = 0; uint32_t sphase = 0;
            long long wk = (long long)atomicAdd(a.work_counter, 1ull);
            for (;;) {
                const bool live = wk < n_work;
                long long wk_next = 0;
                if (live) wk_next = (long long)atomicAdd(a.work_counter, 1ull);  // latency hidden behind this tile's loads
                mbar_wait_u32(bar_sempty + ss * 8, sphase ^ 1u);
                sched_slot[ss] = live ? wk : -1ll;
                mbar_arrive(bar_sfull + ss * 8);
                if (++ss == 2) { ss = 0; sphase ^= 1u; }
                if (!live) break;
                const int nt = (int)(wk / a.n_tiles_m), mt = (int)(wk % a.n_tiles_m);
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait_u32(bar_empty + stage * 8, phase ^ 1u);
                    mbar_expect_tx(bar_full + stage * 8, kStageBytes);
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    tma_load_2d(sa, &tmap_q, kb * kBK, mt * kBM, bar_full + stage * 8);
                    tma_load_2d(sa + kStageBytesA, &tmap_x, kb * kBK, (int)(a.row_offset) + nt * kBN, bar_full + stage * 8);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                wk = wk_next;
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            int ss = 0; uint32_t sphase = 0;
            for (;;) {
                mbar_wait_u32(bar_sfull + ss * 8, sphase);
                const long long wk = sched_slot[ss];
                mbar_arrive(bar_sempty + ss * 8);
                if (++ss == 2) { ss = 0; sphase ^= 1u; }
                if (wk < 0) break;
                {
                    mbar_wait_u32(bar_tempty + acc * 8, acc_phase ^ 1u);  // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)acc * kBN;
                    for (int kb = 0; kb < a.k_blocks; ++kb) {
                        mbar_wait_u32(bar_full + stage * 8, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * kStageBytes;
                        const uint64_t da = umma_desc(sa), db = umma_desc(sa + kStageBytesA);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)   // +32 bytes (2 x 16 B) per UMMA_K inside the swizzle atom
                            tc_mma_f16(d, da + 2 * k, db + 2 * k, a.idesc, (kb | k) != 0);
                        tc_commit(bar_empty + stage * 8);      // smem stage free once these MMAs have read it
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit(bar_tfull + acc * 8);            // accumulator complete
                    acc ^= 1; if (acc == 0) acc_phase ^= 1u;
                }
            }
        }
    } else {
        // ================= epilogue (warp w <-> TMEM lane quarter w % 4; warps 2..5 columns 0..127, 6..9 the rest) ====
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_phase = 0;
        int ss = 0; uint32_t sphase = 0;
        for (;;) {
            mbar_wait_u32(bar_sfull + ss * 8, sphase);
            const long long wk = sched_slot[ss];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sempty + ss * 8);
            if (++ss == 2) { ss = 0; sphase ^= 1u; }
            if (wk < 0) break;
            const int nt = (int)(wk / a.n_tiles_m), mt = (int)(wk % a.n_tiles_m);
            {
                mbar_wait_u32(bar_tfull + acc * 8, acc_phase);
                tc_fence_after();
                const int64_t q = (int64_t)mt * kBM + quarter * 32 + lane;   // this thread's query row
                const int64_t n0 = a.row_offset + (int64_t)nt * kBN + half * (kBN / 2);   // first passage of this warp's half
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * kBN + half * (kBN / 2);
                epilogue_tile(a, taddr, q, n0, s_tau);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + acc * 8);
                acc ^= 1; if (acc == 0) acc_phase ^= 1u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- CTA-pair variant (cta_group::2): M = 256 queries x N = 256 passages per work item ------------------------
// Two CTAs of a cluster (one TPC) share one UMMA: each stages ITS 128 queries (A half) and ITS 128 passages (B half)
// per k-block -- 32 KB instead of 48 KB per CTA per 2 x 128 x 256 x 64 MACs, and a 6-deep ring instead of 4 --
// the leader's MMA thread issues tcgen05.mma.cta_group::2 (the tensor cores of both SMs read both B halves), each
// CTA's TMEM gets the accumulator rows of its own 128 queries, and each CTA runs its own epilogue.
//   full[s]    leader's barrier only: one expect_tx arrival for the 64 KB of BOTH CTAs (the peer's TMA signals it
//              through the cta_group::2 form with the peer bit of the barrier address cleared)
//   empty[s], tfull[a]   in both CTAs, signalled by tcgen05.commit ... multicast::cluster (mask 0b11)
//   tempty[a]  leader's barrier, 8 arrivals: the four epilogue warps of either CTA (remote arrive from the peer)
// Work items are dealt statically (item = pair + i * n_pairs, query tile fastest): both CTAs of a pair derive the
// same sequence without talking.  X tiles then come from HBM more than once (the pairs drift apart), which costs
// < 5 % of the sweep at these arithmetic intensities (measured with the single-CTA kernel, profiles/README.md).
constexpr int kPairStages = 6;
constexpr uint32_t kPairStageBytesA = kBM * kBK * 2, kPairStageBytesB = (kBN / 2) * kBK * 2;
constexpr uint32_t kPairStageBytes = kPairStageBytesA + kPairStageBytesB;   // 32 KB per CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into THIS CTA's shared memory, completion bytes on the LEADER's barrier (same offset, peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar & 0xFEFFFFFFu) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // arrive on the leader CTA's barrier at this offset
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDenseThreads, 1)
dense_topk_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x, const DenseArgs a) {
    extern __shared__ __align__(1024) uint8_t dsmem[];
    uint8_t *base = reinterpret_cast<uint8_t *>(((uintptr_t)dsmem + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(base);
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + kPairStages * kPairStageBytes);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + kPairStages * 8;
    const uint32_t bar_tfull = bar_empty + kPairStages * 8, bar_tempty = bar_tfull + 2 * 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kPairStages + 4);
    float *s_tau = reinterpret_cast<float *>(bars + 2 * kPairStages + 6);   // [n_queries <= 4096] threshold scores (mode 1)
    if (a.mode == 1)
        for (int64_t i = threadIdx.x; i < a.n_queries; i += kDenseThreads) {
            const uint64_t t = a.tau[i];
            s_tau[i] = t ? key_score(t) : -INFINITY;
        }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_x) : "memory");
        for (int s = 0; s < kPairStages; ++s) { mbar_init(bars + s, 1); mbar_init(bars + kPairStages + s, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bars + 2 * kPairStages + i, 1);      // tfull: one multicast commit
            mbar_init(bars + 2 * kPairStages + 2 + i, 2 * kEpiWarps);  // tempty (leader's is used): epilogue warps of both CTAs
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / peer TMA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles_m2 = (a.n_tiles_m + 1) / 2;                       // query tiles of 256
    const long long n_work = (long long)a.n_tiles_n * n_tiles_m2;
    const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0) {
        // ================= TMA producer (both CTAs: own A half, own B half) =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long wk = pair; wk < n_work; wk += n_pairs) {
                const int nt = (int)(wk / n_tiles_m2), mt2 = (int)(wk % n_tiles_m2);
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait_u32(bar_empty + stage * 8, phase ^ 1u);
                    if (a.dbg & 4) {
                        if (leader) mbar_arrive(bar_full + stage * 8);
                    } else {
                    if (leader) mbar_expect_tx(bar_full + stage * 8, 2 * kPairStageBytes);
                    const uint32_t sa = smem_base + stage * kPairStageBytes;
                    tma_load_2d_pair(sa, &tmap_q, kb * kBK, mt2 * 2 * kBM + (int)rank * kBM, bar_full + stage * 8);
                    tma_load_2d_pair(sa + kPairStageBytesA, &tmap_x, kb * kBK,
                                     (int)(a.row_offset) + nt * kBN + (int)rank * (kBN / 2), bar_full + stage * 8);
                    }
                    if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (long long wk = pair; wk < n_work; wk += n_pairs) {
                mbar_wait_u32(bar_tempty + acc * 8, acc_phase ^ 1u);  // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)acc * kBN;
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait_u32(bar_full + stage * 8, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * kPairStageBytes;
                    const uint64_t da = umma_desc(sa), db = umma_desc(sa + kPairStageBytesA);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        if (!(a.dbg & 2)) tc_mma_f16_pair(d, da + 2 * k, db + 2 * k, a.idesc, (kb | k) != 0);
                    tc_commit_pair(bar_empty + stage * 8);
                    if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
                }
                tc_commit_pair(bar_tfull + acc * 8);
                acc ^= 1; if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ================= epilogue (each CTA: its own 128 queries; warp w <-> TMEM lane quarter w % 4, column half) ===
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long wk = pair; wk < n_work; wk += n_pairs) {
            const int nt = (int)(wk / n_tiles_m2), mt2 = (int)(wk % n_tiles_m2);
            mbar_wait_u32(bar_tfull + acc * 8, acc_phase);
            tc_fence_after();
            const int64_t q = (int64_t)mt2 * 2 * kBM + (int64_t)rank * kBM + quarter * 32 + lane;
            const int64_t n0 = a.row_offset + (int64_t)nt * kBN + half * (kBN / 2);
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * kBN + half * (kBN / 2);
            if (!(a.dbg & 1)) epilogue_tile(a, taddr, q, n0, s_tau);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(bar_tempty + acc * 8);
            acc ^= 1; if (acc == 0) acc_phase ^= 1u;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's shared memory / barriers stay valid until both CTAs are done
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- operand preparation ---------------------------------------------------------------------------------
// rows of `in` ([rows, ld] f32 / f16 / bf16) -> 16-bit storage [rows_pad, d_pad] (zero padded)
__global__ void dense_convert_kernel(const void *in, int in_dtype, int64_t rows, int64_t dim, int64_t ld,
                                     uint16_t *out, int out_dtype, int64_t rows_pad, int64_t d_pad) {
    const int64_t total = rows_pad * d_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d_pad, c = i - r * d_pad;
        float v = 0.f;
        if (r < rows && c < dim) {
            if (in_dtype == VS_F32) v = ((const float *)in)[r * ld + c];
            else if (in_dtype == VS_F16) v = __half2float(((const __half *)in)[r * ld + c]);
            else v = __bfloat162float(((const __nv_bfloat16 *)in)[r * ld + c]);
        }
        out[i] = (out_dtype == VS_F16) ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
    }
}

// Start of a filtered sweep: every query's candidate list is seeded with its current top-k (sorted keys, one block
// per query), the threshold is the k-th of them.
// Every sweep starts from the top-k of the rows BEFORE it (`sorted_keys`) and their k-th key as the threshold.  A retry
// after an overflow keeps those seeds -- seeding rows of the sweep itself would append them a second time -- and only
// raises the threshold to the k-th best of what the failed attempt had stored (`retry_keys`).
// tau_ext: [B] thresholds from elsewhere (the other ranks of a row-sharded index, search_dense_step), or nullptr
__global__ void dense_seed_lists_kernel(const uint64_t *sorted_keys, int k, uint64_t *cand, int64_t cap, uint32_t *cnt,
                                        uint64_t *tau, const uint64_t *retry_keys, const uint64_t *tau_ext = nullptr) {
    const int64_t q = blockIdx.x;
    for (int i = threadIdx.x; i < k; i += blockDim.x) cand[q * cap + i] = sorted_keys[q * k + i];
    if (threadIdx.x == 0) {
        cnt[q] = (uint32_t)k;
        uint64_t t = sorted_keys[q * k + (k - 1)];
        if (retry_keys && retry_keys[q * k + (k - 1)] > t) t = retry_keys[q * k + (k - 1)];
        if (tau_ext && tau_ext[q] > t) t = tau_ext[q];
        tau[q] = t;
    }
}

// status word of a call that does not stop to look at its survivor counts: bit `bit` is set when a list went past its
// capacity (the caller polls the word once, at the end)
__global__ void dense_overflow_kernel(const uint32_t *cnt, int64_t n, uint32_t cap, uint32_t *status, uint32_t bit) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && cnt[i] > cap) atomicOr(status, bit);
}

// merged [B, k] keys of all ranks (global ids) -> per-query threshold: every row that scores at least as much as the
// k-th best of the union can still make the global top-k; ids are left out of the comparison (local and global ids
// do not compare), so rows that tie with the k-th score pass
__global__ void dense_tau_from_keys_kernel(const uint64_t *merged, int k, int64_t n, uint64_t *tau_ext) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) tau_ext[q] = merged[q * k + (k - 1)] & 0xffffffff00000000ull;
}

// ---- fp32 storage (upstream's Index(fp16=False): an fp32 [N, D] matrix and an fp32 GEMM, index.py:36-44, 88-94) ----------
// The tensor cores score a bf16 copy; the fp32 matrix stays beside it.  Rounding both operands to bf16 (unit roundoff
// u = 2^-8) changes a score by at most (2u + u^2) * sum_i |q_i x_i| <= E = 2^-7 * |q| * |x| (Cauchy-Schwarz; the fp32
// accumulation errors of either side are three orders below).  With s_k = the k-th best bf16 score, every row of the
// true fp32 top-k has a bf16 score >= s_k - 2E: one more filtered sweep with that threshold collects a superset of the
// true top-k, its members are re-scored exactly from the fp32 rows on the CUDA cores, and the merge ranks those.
constexpr float kBf16PairErr = 1.02f * 0.0078125f;   // 2^-7 with a 2 % margin for the second-order terms

__global__ void row_norm_max_kernel(const float *x, int64_t n_rows, int64_t dim, unsigned int *max_norm_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float best = 0.f;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        float s = 0.f;
        for (int64_t c = lane; c < dim; c += 32) { const float v = x[r * dim + c]; s = fmaf(v, v, s); }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        best = fmaxf(best, s);
    }
    if (lane == 0) atomicMax(max_norm_bits, __float_as_uint(sqrtf(best) * 1.000001f));   // non-negative floats order like their bits
}

__global__ void copy_f32_kernel(const void *in, int in_dtype, int64_t rows, int64_t dim, int64_t ld, float *out) {
    const int64_t total = rows * dim;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / dim, c = i - r * dim;
        float v;
        if (in_dtype == VS_F32) v = ((const float *)in)[r * ld + c];
        else if (in_dtype == VS_F16) v = __half2float(((const __half *)in)[r * ld + c]);
        else v = __bfloat162float(((const __nv_bfloat16 *)in)[r * ld + c]);
        out[i] = v;
    }
}

__device__ __forceinline__ float load_q(const void *q, int dtype, int64_t i) {
    if (dtype == VS_F32) return ((const float *)q)[i];
    if (dtype == VS_F16) return __half2float(((const __half *)q)[i]);
    return __bfloat162float(((const __nv_bfloat16 *)q)[i]);
}

// one block per query: threshold key of the superset sweep = (k-th best bf16 score - 2E, lowest rank for that score)
__global__ void dense_relax_kernel(const uint64_t *sorted_keys, int k, const void *q, int q_dtype, int64_t ldq, int64_t dim,
                                   float max_norm, uint64_t *tau, uint32_t *cnt) {
    __shared__ float s_part[8];
    const int64_t b = blockIdx.x;
    float s = 0.f;
    for (int64_t c = threadIdx.x; c < dim; c += blockDim.x) { const float v = load_q(q, q_dtype, b * ldq + c); s = fmaf(v, v, s); }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += s_part[i];
        const float E = kBf16PairErr * sqrtf(tot) * 1.000001f * max_norm;
        const uint64_t kth = sorted_keys[b * k + (k - 1)];
        // fewer than k rows in the index cannot happen (k <= N is checked); kth == 0 would mean an empty list
        const float thr = kth ? key_score(kth) - 2.0f * E : -INFINITY;
        // the sweep appends keys > tau: take the predecessor of the lowest key with score `thr`
        tau[b] = make_key(thr, 0xffffffffu) - 1ull;
        cnt[b] = 0u;
    }
}

// exact fp32 scores of the candidates: one warp per (query, candidate), the candidate's fp32 row gathered from HBM
__global__ void __launch_bounds__(256) dense_rescore_kernel(const float *x32, int64_t dim, const void *q, int q_dtype, int64_t ldq,
                                                            uint64_t *cand, const uint32_t *cnt, int64_t cap) {
    const int lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x;
    const uint32_t n = min(cnt[b], (uint32_t)cap);
    for (uint32_t j = blockIdx.y * 8 + (threadIdx.x >> 5); j < n; j += gridDim.y * 8) {
        const uint32_t id = key_id(cand[b * cap + j]);
        const float *row = x32 + (int64_t)id * dim;
        float s = 0.f;
        for (int64_t c = lane; c < dim; c += 32) s = fmaf(load_q(q, q_dtype, b * ldq + c), row[c], s);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) cand[b * cap + j] = make_key(s, id);
    }
}

static PFN_cuTensorMapEncodeTiled get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled)p;
    }
    return fn;
}

// 2-D K-major tensor map: [rows, d_pad] 16-bit, box = 64 columns x box_rows rows, 128-byte swizzle
static int make_tmap(CUtensorMap *map, const void *ptr, int dtype, int64_t rows, int64_t d_pad, int box_rows) {
    PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
    VS_REQUIRE(enc != nullptr, VS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)d_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dtype == VS_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VS_REQUIRE(r == CUDA_SUCCESS, VS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VS_OK;
}

int build_dense_index(vs_index *idx, const void *d_x, int x_dtype, int64_t ld, cudaStream_t st) {
    idx->d_pad = (idx->dim + kBK - 1) / kBK * kBK;
    idx->n_pad = (idx->n_rows + kBN - 1) / kBN * kBN;
    if (idx->n_pad == 0) idx->n_pad = kBN;
    idx->mma_dtype = idx->store_dtype == VS_F32 ? VS_BF16 : idx->store_dtype;   // what the tensor cores read
    VS_CUDA(cudaMalloc(&idx->dense, (size_t)idx->n_pad * idx->d_pad * 2));
    dense_convert_kernel<<<2048, 256, 0, st>>>(d_x, x_dtype, idx->n_rows, idx->dim, ld, (uint16_t *)idx->dense,
                                               idx->mma_dtype, idx->n_pad, idx->d_pad);
    VS_CUDA(cudaGetLastError());
    idx->device_bytes = idx->n_pad * idx->d_pad * 2;
    if (idx->store_dtype == VS_F32) {   // fp32 semantics: the exact rows for the re-score, and the largest row norm
        unsigned int *d_bits = nullptr;
        VS_CUDA(cudaMalloc(&idx->dense32, idx->n_rows ? (size_t)idx->n_rows * idx->dim * 4 : 16));
        VS_CUDA(cudaMalloc(&d_bits, 4));
        VS_CUDA(cudaMemsetAsync(d_bits, 0, 4, st));
        copy_f32_kernel<<<2048, 256, 0, st>>>(d_x, x_dtype, idx->n_rows, idx->dim, ld, idx->dense32);
        row_norm_max_kernel<<<1024, 256, 0, st>>>(idx->dense32, idx->n_rows, idx->dim, d_bits);
        unsigned int h_bits = 0;
        cudaError_t e = cudaMemcpyAsync(&h_bits, d_bits, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(d_bits);
        VS_CUDA(e);
        memcpy(&idx->max_row_norm, &h_bits, 4);
        idx->device_bytes += idx->n_rows * idx->dim * 4;
    }
    VS_CUDA(cudaStreamSynchronize(st));
    int rc = make_tmap(reinterpret_cast<CUtensorMap *>(idx->tmap_x), idx->dense, idx->mma_dtype, idx->n_pad, idx->d_pad, kBN);
    if (rc) return rc;
    rc = make_tmap(reinterpret_cast<CUtensorMap *>(idx->tmap_x_half), idx->dense, idx->mma_dtype, idx->n_pad, idx->d_pad, kBN / 2);   // CTA-pair kernel: B halves
    if (rc) return rc;
    cudaDeviceProp prop;
    VS_CUDA(cudaGetDeviceProperties(&prop, idx->device));
    idx->n_ctas = prop.multiProcessorCount;
    idx->stream_bytes = idx->device_bytes;
    return VS_OK;
}

int launch_merge(const uint64_t *d_in, int64_t P, int64_t stride_p, int64_t stride_b, int64_t B, int k_in, int k_out,
                 int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, cudaStream_t st);
int launch_merge_counted(const uint64_t *d_in, const uint32_t *d_counts, int64_t P, int64_t stride_p, int64_t stride_b,
                         int64_t B, int k_in, int k_out, int64_t id_offset, int64_t *d_ids, float *d_scores,
                         uint64_t *d_keys, cudaStream_t st, const int *d_alt_flag = nullptr, int k_in_alt = 0);

// workspace carve for one query chunk (behind the call's 1 KB status block)
struct DenseWs {
    uint16_t *q16;        // [b_pad, d_pad]
    uint64_t *sample;     // [Bc, sample_rows]
    uint64_t *tau;        // [Bc]
    uint64_t *tau_ext;    // [Bc]     thresholds learnt from the other ranks (search_dense_step)
    uint64_t *tau_sorted; // [Bc, k]  top-k of the rows before the current sweep
    uint64_t *tau_retry;  // [Bc, k]  top-k of what an overflowed attempt stored / merged keys of all ranks
    uint32_t *cnt;        // [Bc]
    unsigned long long *work_counter;
    uint64_t *cand;       // [Bc, cap]
    size_t bytes;
};
constexpr int64_t kDenseQueryChunk = 4096;
constexpr int64_t kDenseCandCap = 1 << 16;   // per-query survivor list (keys)
constexpr size_t kDenseStatusBytes = 1024;   // head of the workspace: the call's status word
constexpr uint32_t kDenseListOverflow = 1u, kDenseSupersetOverflow = 2u;

static int64_t dense_sample_rows(const vs_index *idx, int k, int n_ranks = 1) {
    // sample prefix: ~64*k rows (>= 16 K) -- of the whole index: a rank of a row-sharded index sweeps its share and the
    // ranks pool what they found --, at least k rows, a whole number of tiles, at most the index
    int64_t s = (int64_t)k * 64;
    if (s < 16384) s = 16384;
    s = (s + n_ranks - 1) / n_ranks;
    if (s < k) s = k;
    s = (s + kBN - 1) / kBN * kBN;
    return s < idx->n_pad ? s : idx->n_pad;
}

static DenseWs carve_dense(const vs_index *idx, void *base, int64_t Bc, int k) {
    DenseWs w;
    auto al = [](size_t x) { return (x + 1023) / 1024 * 1024; };
    const int64_t b_pad = (Bc + 2 * kBM - 1) / (2 * kBM) * (2 * kBM);   // whole 256-query pair tiles
    size_t o = kDenseStatusBytes;
    uint8_t *p = (uint8_t *)base;
    w.q16 = (uint16_t *)(p + o); o += al((size_t)b_pad * idx->d_pad * 2);
    w.sample = (uint64_t *)(p + o); o += al((size_t)Bc * dense_sample_rows(idx, k) * 8);
    w.tau = (uint64_t *)(p + o); o += al((size_t)Bc * 8);
    w.tau_ext = (uint64_t *)(p + o); o += al((size_t)Bc * 8);
    w.tau_sorted = (uint64_t *)(p + o); o += al((size_t)Bc * k * 8);
    w.tau_retry = (uint64_t *)(p + o); o += al((size_t)Bc * k * 8);
    w.cnt = (uint32_t *)(p + o); o += al((size_t)Bc * 4);
    w.work_counter = (unsigned long long *)(p + o); o += al(8);
    w.cand = (uint64_t *)(p + o); o += al((size_t)Bc * kDenseCandCap * 8);
    w.bytes = o;
    return w;
}

size_t dense_workspace_bytes(const vs_index *idx, int64_t B, int k) {
    int64_t Bc = B < kDenseQueryChunk ? B : kDenseQueryChunk;
    return carve_dense(idx, nullptr, Bc, k).bytes + 1024;
}

static bool dense_use_pair(const DenseArgs &a) {
    static const bool off = getenv("VSEARCH_B200_DENSE_PAIR") && atoi(getenv("VSEARCH_B200_DENSE_PAIR")) == 0;
    return !off && a.n_tiles_m >= 2;   // a single 128-query tile would leave the peer CTA's half of the MMA empty
}

static int launch_dense(const vs_index *idx, const CUtensorMap &tq, const CUtensorMap &tx, const CUtensorMap &tx_half,
                        DenseArgs a, cudaStream_t st) {
    if (dense_use_pair(a)) {
        const size_t smem = (size_t)kPairStages * kPairStageBytes + 256 + (size_t)kDenseQueryChunk * 4 + 1024;
        VS_CUDA(cudaFuncSetAttribute(dense_topk_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t n_work = (int64_t)a.n_tiles_n * ((a.n_tiles_m + 1) / 2);
        int64_t pairs = idx->n_ctas / 2;
        if (pairs > n_work) pairs = n_work;
        a.idesc = (a.idesc & ~(0x1fu << 24)) | ((uint32_t)((2 * kBM) >> 4) << 24);   // M = 256 across the pair
        if (getenv("VSEARCH_B200_DEBUG")) fprintf(stderr, "[vsearch_b200] dense pair kernel: %lld pairs, %lld items, idesc %08x\n", (long long)pairs, (long long)n_work, a.idesc);
        dense_topk_pair_kernel<<<(unsigned)(2 * pairs), kDenseThreads, smem, st>>>(tq, tx_half, a);
        VS_CUDA(cudaGetLastError());
        return VS_OK;
    }
    const size_t smem = (size_t)kStages * kStageBytes + 256 + (size_t)kDenseQueryChunk * 4 + 1024;
    VS_CUDA(cudaFuncSetAttribute(dense_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_work = (int64_t)a.n_tiles_n * a.n_tiles_m;
    int grid = (int64_t)idx->n_ctas < n_work ? idx->n_ctas : (int)n_work;
    VS_CUDA(cudaMemsetAsync(a.work_counter, 0, 8, st));
    dense_topk_kernel<<<grid, kDenseThreads, smem, st>>>(tq, tx, a);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

// ---- one chunk (<= kDenseQueryChunk queries) of a dense search: the pieces both drivers below are made of
struct DenseRun {
    vs_index *idx;
    cudaStream_t st;
    int k;
    int64_t Bc;
    DenseWs w;
    DenseArgs a;
    const CUtensorMap *tq, *tx, *tx_half;
    const uint8_t *qsrc;
    int q_dtype;
    int64_t ldq;
    uint32_t *status;
};

static int dense_begin(DenseRun &r, vs_index *idx, const uint8_t *qsrc, int q_dtype, int64_t Bc, int64_t ldq, int k, int score_round,
                       void *ws_base, bool convert, cudaStream_t st) {
    r.idx = idx; r.st = st; r.k = k; r.Bc = Bc; r.qsrc = qsrc; r.q_dtype = q_dtype; r.ldq = ldq;
    r.status = (uint32_t *)ws_base;
    r.w = carve_dense(idx, ws_base, Bc, k);
    const int64_t b_pad = (Bc + 2 * kBM - 1) / (2 * kBM) * (2 * kBM);
    if (convert) {
        dense_convert_kernel<<<1024, 256, 0, st>>>(qsrc, q_dtype, Bc, idx->dim, ldq, r.w.q16, idx->mma_dtype, b_pad, idx->d_pad);
        VS_CUDA(cudaGetLastError());
    }
    if (idx->tmap_q_ptr != (const void *)r.w.q16 || idx->tmap_q_rows != b_pad) {   // a new workspace or batch height
        int rc = make_tmap(reinterpret_cast<CUtensorMap *>(idx->tmap_q), r.w.q16, idx->mma_dtype, b_pad, idx->d_pad, kBM);
        if (rc) return rc;
        idx->tmap_q_ptr = r.w.q16; idx->tmap_q_rows = b_pad;
    }
    r.tq = reinterpret_cast<const CUtensorMap *>(idx->tmap_q);
    r.tx = reinterpret_cast<const CUtensorMap *>(idx->tmap_x);
    r.tx_half = reinterpret_cast<const CUtensorMap *>(idx->tmap_x_half);
    const uint32_t fmt = idx->mma_dtype == VS_F16 ? 0u : 1u;
    DenseArgs &a = r.a;
    a = DenseArgs{};
    a.n_tiles_m = (int)((Bc + kBM - 1) / kBM);
    a.k_blocks = (int)(idx->d_pad / kBK);
    a.n_rows = idx->n_rows; a.n_queries = Bc; a.row_offset = 0;
    a.score_round = score_round;
    a.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    a.sample_keys = r.w.sample; a.tau = r.w.tau; a.cand = r.w.cand; a.cand_cnt = r.w.cnt; a.cand_cap = kDenseCandCap;
    a.work_counter = r.w.work_counter;
    a.dbg = getenv("VSEARCH_B200_DENSE_DBG") ? atoi(getenv("VSEARCH_B200_DENSE_DBG")) : 0;
    return VS_OK;
}

// sample sweep: every score of rows [0, s_rows) becomes a key, their exact top-k goes to the given outputs
static int dense_sample_pass(DenseRun &r, int64_t s_rows, int64_t id_offset, int64_t *ids, float *scores, uint64_t *keys) {
    DenseArgs a = r.a;
    a.mode = 0; a.n_tiles_n = (int)(s_rows / kBN); a.sample_ld = s_rows; a.row_offset = 0;
    int rc = launch_dense(r.idx, *r.tq, *r.tx, *r.tx_half, a, r.st);
    if (rc) return rc;
    return launch_merge(r.w.sample, 1, 0, s_rows, r.Bc, (int)s_rows, r.k, id_offset, ids, scores, keys, r.st);
}

// filtered sweep over rows [row, row + rows): every list starts with the current top-k (w.tau_sorted), rows at or above the
// threshold (its k-th key, raised by `retry` / `tau_ext` where given) are appended
static int dense_filtered_pass(DenseRun &r, int64_t row, int64_t rows, const uint64_t *retry, const uint64_t *tau_ext) {
    dense_seed_lists_kernel<<<(unsigned)r.Bc, 128, 0, r.st>>>(r.w.tau_sorted, r.k, r.w.cand, kDenseCandCap, r.w.cnt, r.w.tau, retry, tau_ext);
    VS_CUDA(cudaGetLastError());
    if (rows <= 0) return VS_OK;
    DenseArgs a = r.a;
    a.mode = 1; a.row_offset = row; a.n_tiles_n = (int)(rows / kBN);
    return launch_dense(r.idx, *r.tq, *r.tx, *r.tx_half, a, r.st);
}

static int dense_flag_overflow(DenseRun &r, uint32_t bit) {
    dense_overflow_kernel<<<(unsigned)((r.Bc + 255) / 256), 256, 0, r.st>>>(r.w.cnt, r.Bc, (uint32_t)kDenseCandCap, r.status, bit);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

static int dense_survivor_topk(DenseRun &r, int64_t id_offset, int64_t *ids, float *scores, uint64_t *keys) {
    return launch_merge_counted(r.w.cand, r.w.cnt, 1, 0, kDenseCandCap, r.Bc, (int)kDenseCandCap, r.k, id_offset, ids, scores, keys, r.st);
}

// fp32 storage: superset sweep + exact re-score, run after the bf16 top-k of the chunk is in w.tau_sorted.  `checked`:
// stop and read the survivor counts (the fallback driver); otherwise they only raise the status word.
static int dense_finish_exact32(DenseRun &r, bool checked, int64_t id_offset, int64_t *ids, float *scores, uint64_t *keys) {
    vs_index *idx = r.idx;
    dense_relax_kernel<<<(unsigned)r.Bc, 256, 0, r.st>>>(r.w.tau_sorted, r.k, r.qsrc, r.q_dtype, r.ldq, idx->dim, idx->max_row_norm, r.w.tau, r.w.cnt);
    DenseArgs e = r.a;
    e.score_round = VS_F32; e.sample_ld = 0; e.dbg = 0;
    e.mode = 1; e.row_offset = 0; e.n_tiles_n = (int)(idx->n_pad / kBN);
    int rc = launch_dense(idx, *r.tq, *r.tx, *r.tx_half, e, r.st);
    if (rc) return rc;
    unsigned gy = 16;
    if (checked) {
        std::vector<uint32_t> h_cnt((size_t)r.Bc);
        VS_CUDA(cudaMemcpyAsync(h_cnt.data(), r.w.cnt, (size_t)r.Bc * 4, cudaMemcpyDeviceToHost, r.st));
        VS_CUDA(cudaStreamSynchronize(r.st));
        uint32_t mx = 0;
        for (int64_t i = 0; i < r.Bc; ++i) mx = h_cnt[i] > mx ? h_cnt[i] : mx;
        VS_REQUIRE(mx <= (uint32_t)kDenseCandCap, VS_ERR_UNSUPPORTED,
                   "fp32 dense search: %u passages lie within the bf16 error bound of a query's k-th score (limit %lld)", mx,
                   (long long)kDenseCandCap);
        gy = (mx + 7) / 8;
        gy = gy < 1 ? 1 : (gy > 64 ? 64 : gy);
    } else if ((rc = dense_flag_overflow(r, kDenseSupersetOverflow)) != VS_OK) {
        return rc;
    }
    dense_rescore_kernel<<<dim3((unsigned)r.Bc, gy), 256, 0, r.st>>>(idx->dense32, idx->dim, r.qsrc, r.q_dtype, r.ldq, r.w.cand, r.w.cnt, kDenseCandCap);
    VS_CUDA(cudaGetLastError());
    return dense_survivor_topk(r, id_offset, ids, scores, keys);
}

// One chunk, start to finish.  checked = false: nothing is read back -- a survivor list that ran past its capacity only
// raises the status word, which search_dense polls once at the end of the call.  checked = true (the fallback for such
// a call: adversarial row order): every sweep's counts are read (SYNC) and an overflowed sweep is redone with the
// k-th best of what it stored as the new, strictly tighter threshold.
static int dense_chunk(vs_index *idx, const uint8_t *qsrc, int q_dtype, int64_t Bc, int64_t ldq, int k, int score_round,
                       int64_t id_offset, int64_t *ids, float *scores, uint64_t *keys, void *ws_base, bool checked, cudaStream_t st) {
    const bool exact32 = idx->store_dtype == VS_F32;   // bf16 sweep -> superset with an error margin -> exact fp32 re-score
    DenseRun r;
    int rc = dense_begin(r, idx, qsrc, q_dtype, Bc, ldq, k, exact32 ? VS_F32 : score_round, ws_base, true, st);
    if (rc) return rc;
    // in exact32 mode the bf16 pipeline delivers its top-k as keys (local ids) into w.tau_sorted
    int64_t *const p_ids = exact32 ? nullptr : ids;
    float *const p_scores = exact32 ? nullptr : scores;
    uint64_t *const p_keys = exact32 ? r.w.tau_sorted : keys;
    const int64_t p_off = exact32 ? 0 : id_offset;

    // ---- pass 1 (sample sweep): exact top-k of the first S1 rows -> threshold tau1
    const int64_t s1 = dense_sample_rows(idx, k);
    const int slot = idx->timer_n < VS_TIMER_SLOTS ? idx->timer_n : -1;
    if (slot >= 0) VS_CUDA(cudaEventRecord(idx->ev0[slot], st));
    if (s1 >= idx->n_rows) {   // the sample IS the index (small index): its top-k is the answer
        rc = dense_sample_pass(r, s1, p_off, p_ids, p_scores, p_keys);
        if (slot >= 0) VS_CUDA(cudaEventRecord(idx->ev1[slot], st));
        idx->timer_n += 1;
        if (rc) return rc;
        return exact32 ? dense_finish_exact32(r, checked, id_offset, ids, scores, keys) : VS_OK;
    }
    rc = dense_sample_pass(r, s1, 0, nullptr, nullptr, r.w.tau_sorted);
    if (rc) return rc;
    // ---- passes 2 and 3 (filtered sweeps).  Pass 2 covers the next ~64 x S1 rows with tau1 and tightens the
    // threshold to the k-th best of everything seen so far; pass 3 covers the rest of the index with that
    // threshold, so only ~k * N / (65 * S1) rows per query survive it.
    int64_t row = s1;
    for (int pass = 2; row < idx->n_pad; ++pass) {
        int64_t rows = (pass == 2) ? 64 * s1 : idx->n_pad - row;
        if (rows > idx->n_pad - row) rows = idx->n_pad - row;
        for (int attempt = 0;; ++attempt) {
            rc = dense_filtered_pass(r, row, rows, attempt ? r.w.tau_retry : nullptr, nullptr);
            if (rc) return rc;
            if (!checked) {
                if ((rc = dense_flag_overflow(r, kDenseListOverflow)) != VS_OK) return rc;
                break;
            }
            std::vector<uint32_t> h_cnt((size_t)Bc);
            VS_CUDA(cudaMemcpyAsync(h_cnt.data(), r.w.cnt, (size_t)Bc * 4, cudaMemcpyDeviceToHost, st));
            VS_CUDA(cudaStreamSynchronize(st));
            uint32_t mx = 0;
            for (int64_t i = 0; i < Bc; ++i) mx = h_cnt[i] > mx ? h_cnt[i] : mx;
            if (mx <= (uint32_t)kDenseCandCap) break;
            VS_REQUIRE(attempt < 8, VS_ERR_UNSUPPORTED, "dense candidate lists keep overflowing");
            // (lists that did not overflow keep their own count: only their first cnt[q] entries are valid)
            rc = dense_survivor_topk(r, 0, nullptr, nullptr, r.w.tau_retry);
            if (rc) return rc;
        }
        row += rows;
        if (row < idx->n_pad && (rc = dense_survivor_topk(r, 0, nullptr, nullptr, r.w.tau_sorted)) != VS_OK) return rc;
    }
    if (slot >= 0) VS_CUDA(cudaEventRecord(idx->ev1[slot], st));
    idx->timer_n += 1;
    // exact top-k of the survivors (only the first cnt[q] entries of each list are valid)
    rc = dense_survivor_topk(r, p_off, p_ids, p_scores, p_keys);
    if (rc) return rc;
    return exact32 ? dense_finish_exact32(r, checked, id_offset, ids, scores, keys) : VS_OK;
}

static int dense_all_chunks(vs_index *idx, const void *d_q, int q_dtype, int64_t B, int64_t ldq, int k, int score_round,
                            int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, void *ws_base, bool checked,
                            cudaStream_t st) {
    for (int64_t b0 = 0; b0 < B; b0 += kDenseQueryChunk) {
        const int64_t Bc = (B - b0) < kDenseQueryChunk ? (B - b0) : kDenseQueryChunk;
        const uint8_t *qsrc = (const uint8_t *)d_q + (size_t)b0 * ldq * (q_dtype == VS_F32 ? 4 : 2);
        int rc = dense_chunk(idx, qsrc, q_dtype, Bc, ldq, k, score_round, id_offset, d_ids ? d_ids + b0 * k : nullptr,
                             d_scores ? d_scores + b0 * k : nullptr, d_keys ? d_keys + b0 * k : nullptr, ws_base, checked, st);
        if (rc) return rc;
    }
    return VS_OK;
}

// d_q: device queries [B, ldq]; outputs like the sparse path (ids/scores or keys).  The whole call is enqueued without
// looking at intermediate results; SYNC once at the end (the status word: did any survivor list overflow?).
int search_dense(vs_index *idx, const void *d_q, int q_dtype, int64_t B, int64_t ldq, int k, int score_round,
                 int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, void *d_ws, cudaStream_t st) {
    VS_REQUIRE(idx->n_rows + id_offset < 0xffffffffll, VS_ERR_UNSUPPORTED, "global ids must fit 32 bits");
    void *ws_base = (void *)(((uintptr_t)d_ws + 1023) / 1024 * 1024);
    VS_CUDA(cudaMemsetAsync(ws_base, 0, 4, st));
    int rc = dense_all_chunks(idx, d_q, q_dtype, B, ldq, k, score_round, id_offset, d_ids, d_scores, d_keys, ws_base, false, st);
    if (rc) return rc;
    uint32_t status = 0;
    VS_CUDA(cudaMemcpyAsync(&status, ws_base, 4, cudaMemcpyDeviceToHost, st));
    VS_CUDA(cudaStreamSynchronize(st));
    VS_REQUIRE(!(status & kDenseSupersetOverflow), VS_ERR_UNSUPPORTED,
               "fp32 dense search: more than %lld passages lie within the bf16 error bound of a query's k-th score",
               (long long)kDenseCandCap);
    if (status & kDenseListOverflow)   // adversarial row order: redo the call with per-sweep checks and retries
        return dense_all_chunks(idx, d_q, q_dtype, B, ldq, k, score_round, id_offset, d_ids, d_scores, d_keys, ws_base, true, st);
    return VS_OK;
}

// keys with local ids -> global ids (empty keys stay empty)
__global__ void dense_keys_offset_kernel(const uint64_t *in, uint64_t *out, int64_t n, uint32_t id_offset) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const uint64_t x = in[i]; out[i] = x ? ((x & 0xffffffff00000000ull) | (uint64_t)(uint32_t)~(key_id(x) + id_offset)) : 0ull; }
}

// Row-sharded dense search, one rank's part in three steps with an all-gather of [B, k] keys after each (sharded.py):
//   step 0  sample sweep over this rank's share of the sample prefix        -> d_keys_out: its top-k (global ids)
//   step 1  d_gathered = every rank's step-0 keys [n_ranks, B, k]: their merged k-th score is a threshold no rank could
//           have found alone; filtered sweep over the next 64 x share rows   -> d_keys_out: top-k of what this rank has seen
//   step 2  the same with the step-1 keys, over the rest of the rows         -> d_keys_out: this rank's final top-k;
//           *d_status (device) != 0 when a survivor list overflowed: the caller redoes the search through vs_search.
// The workspace (same pointer in all three steps) carries the state.  B <= 4096, 16-bit index.  Nothing is read back.
int search_dense_step(vs_index *idx, int step, const void *d_q, int q_dtype, int64_t B, int64_t ldq, int k, int score_round,
                      int64_t id_offset, int n_ranks, const uint64_t *d_gathered, uint64_t *d_keys_out, uint32_t *d_status,
                      void *d_ws, cudaStream_t st) {
    VS_REQUIRE(idx->kind == 0 && idx->store_dtype != VS_F32, VS_ERR_UNSUPPORTED, "stepwise search is for 16-bit dense indices");
    VS_REQUIRE(B >= 1 && B <= kDenseQueryChunk, VS_ERR_INVALID, "stepwise search takes 1..%lld queries per call", (long long)kDenseQueryChunk);
    VS_REQUIRE(step >= 0 && step <= 2 && n_ranks >= 1, VS_ERR_INVALID, "bad step / n_ranks");
    VS_REQUIRE(idx->n_rows + id_offset < 0xffffffffll, VS_ERR_UNSUPPORTED, "global ids must fit 32 bits");
    VS_REQUIRE(step == 0 || d_gathered != nullptr, VS_ERR_INVALID, "steps 1 and 2 need the gathered keys");
    void *ws_base = (void *)(((uintptr_t)d_ws + 1023) / 1024 * 1024);
    DenseRun r;
    int rc = dense_begin(r, idx, (const uint8_t *)d_q, q_dtype, B, ldq, k, score_round, ws_base, step == 0, st);
    if (rc) return rc;
    const int64_t s_loc = dense_sample_rows(idx, k, n_ranks);
    int64_t r2 = 64 * s_loc;
    if (r2 > idx->n_pad - s_loc) r2 = idx->n_pad - s_loc;
    const int slot = idx->timer_n < VS_TIMER_SLOTS ? idx->timer_n : -1;
    const unsigned nblk = (unsigned)((B * k + 255) / 256);
    if (step == 0) {
        VS_CUDA(cudaMemsetAsync(ws_base, 0, 4, st));
        if (slot >= 0) VS_CUDA(cudaEventRecord(idx->ev0[slot], st));
        rc = dense_sample_pass(r, s_loc, 0, nullptr, nullptr, r.w.tau_sorted);
        if (rc) return rc;
    } else {
        // the k-th best of all ranks' keys -> external threshold
        rc = launch_merge(d_gathered, n_ranks, B * k, k, B, k, k, 0, nullptr, nullptr, r.w.tau_retry, st);
        if (rc) return rc;
        dense_tau_from_keys_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(r.w.tau_retry, k, B, r.w.tau_ext);
        VS_CUDA(cudaGetLastError());
        const int64_t row = step == 1 ? s_loc : s_loc + r2;
        const int64_t rows = step == 1 ? r2 : idx->n_pad - row;
        rc = dense_filtered_pass(r, row, rows, nullptr, r.w.tau_ext);
        if (rc) return rc;
        if ((rc = dense_flag_overflow(r, kDenseListOverflow)) != VS_OK) return rc;
        if (step == 2) {
            if (slot >= 0) VS_CUDA(cudaEventRecord(idx->ev1[slot], st));
            idx->timer_n += 1;
        }
        rc = dense_survivor_topk(r, 0, nullptr, nullptr, r.w.tau_sorted);
        if (rc) return rc;
    }
    dense_keys_offset_kernel<<<nblk, 256, 0, st>>>(r.w.tau_sorted, d_keys_out, B * k, (uint32_t)id_offset);
    VS_CUDA(cudaGetLastError());
    if (step == 2 && d_status) VS_CUDA(cudaMemcpyAsync(d_status, ws_base, 4, cudaMemcpyDeviceToDevice, st));
    return VS_OK;
}

}  // namespace vs
