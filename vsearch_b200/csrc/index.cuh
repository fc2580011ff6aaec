// index.cuh -- device-resident index formats.
//
// Sparse / binary ("warp-stream", WS) format.  One scan pass reads it front to back:
//   * entries are uint16 column ids packed 8 to a 16-byte CHUNK; every row owns
//     ceil(len/8) chunks (>= 1), padded with the sentinel column V whose query slot is 0;
//   * 32 consecutive chunks form a WINDOW = one coalesced 512-byte warp load;
//   * a scan STEP of one warp covers 32 * C chunks, C = chunks per lane (8 for the binary index, 2 for
//     16-bit values, 1 for fp32 values).  In LOGICAL order (the order rows run in) lane l of a step owns
//     chunks l*C .. l*C+C-1, so a row of ~15 chunks spans only 2-3 lanes and ONE segmented warp scan
//     serves 32*C chunks; in MEMORY the step is stored transposed -- logical chunk l*C+i sits at
//     physical position i*32+l -- so that load #i of all 32 lanes is one coalesced 512-byte window
//     (ws_phys_chunk());
//   * the stream is cut into n_parts contiguous PARTS (one per resident warp of the
//     persistent scan grid: #SMs x 24), each a whole number of steps holding whole
//     rows, balanced by chunk count;
//   * bit 15 of a chunk's FIRST entry is set when the chunk is the LAST chunk of its row (so
//     n_cols <= 32,767): the scan kernel turns per-lane partial dot products into row scores
//     with one segmented warp scan and reads no side stream and no row pointers;
//   * tails[w] keeps the same flags as one bit per LOGICAL chunk for the builders (bank
//     placement, export); the scan does not read it;
//   * values (fp32 / fp16 / bf16) sit in a parallel array with the same chunk geometry;
//     the binary bag-of-token index has none.
// Algorithmic bytes of a pass: nnz*(2+b_val) + (N+1)*4.  The format streams
// n_windows*32*(16 + 8*b_val) bytes; the overhead is the row padding (<= 7
// entries per row) and is reported by vs_index_info(stream_bytes).
#pragma once
#include "common.cuh"

// windows of readable slack allocated past the end of the cols/vals/tails streams (prefetch never checks bounds)
constexpr int kStreamSlack = 16;
// warps per persistent scan CTA (= stream parts per SM).  20 warps x 96 registers: the binary kernel keeps a whole
// 8-chunk step in flight per lane (32 registers) and does not spill; 24 x 80 spills and measures 10 % slower;
// 32 x 64 spills the prefetch ring.
#ifndef VS_SCAN_WARPS
#define VS_SCAN_WARPS 20
#endif
constexpr int kScanWarps = VS_SCAN_WARPS;

struct vs_index {
    int device = 0;
    int kind = 0;         // 0 dense, 1 sparse (valued), 2 binary
    int store_dtype = VS_F32;
    int64_t n_rows = 0, n_cols = 0, nnz = 0;

    // ---- WS format
    bool bank_aware = true;   // bank-aware entry placement at build (VSEARCH_B200_BANK_AWARE=0 disables: A/B runs)
    int cpl_shift = 0;        // log2(chunks per lane per scan step): 3 binary, 1 fp16/bf16 values, 0 fp32 values
    int n_ctas = 0;           // persistent scan grid (= #SMs at build time)
    int warps_per_cta = kScanWarps;
    int n_parts = 0;          // n_ctas * warps_per_cta
    uint64_t n_windows = 0;
    uint4 *cols = nullptr;            // n_windows * 32 chunks
    void *vals = nullptr;             // f32: 2 x uint4 per chunk; f16/bf16: 1 x uint4 per chunk
    uint32_t *tails = nullptr;        // n_windows
    uint32_t *part_win_begin = nullptr;  // n_parts + 1
    uint32_t *part_row_begin = nullptr;  // n_parts + 1
    uint32_t *row_chunk = nullptr;       // N + 1: first chunk of each row in the stream (export, rerank)

    // ---- K3 block-partitioned token-major inverted lists (built lazily from the WS stream on first use)
    bool inv_built = false;
    bool inv_has_long = false;   // some (block, token) list is longer than kLongList: K3 runs its long-list variant
    // rows are cut into n_blocks blocks of blk_rows rows; each block has its own token-major lists (inverted.cu)
    int blk_rows = 0, n_blocks = 0, blocks_per_cta = 0;
    uint64_t *post_ptr = nullptr;   // n_cols + 1: global postings per token (exclusive prefix; cost model)
    uint32_t *blk_ptr = nullptr;    // n_blocks x (n_cols + 1): list offsets inside the block
    uint64_t *blk_base = nullptr;   // n_blocks + 1: first posting of each block
    uint16_t *post_row = nullptr;   // nnz: block-local row ids, grouped by (block, token)
    void *post_val = nullptr;       // nnz values in store_dtype (valued index only)
    int64_t inv_bytes = 0;

    // ---- dense
    int64_t dim = 0, d_pad = 0, n_pad = 0;   // logical width; width / rows padded to the GEMM tile
    void *dense = nullptr;                   // [n_pad, d_pad] bf16 or fp16, K-major
    int mma_dtype = VS_BF16;                 // dtype of `dense` (store_dtype, or bf16 when the index keeps fp32 semantics)
    float *dense32 = nullptr;                // store_dtype == VS_F32: the exact rows [n_rows, dim] for the re-score
    float max_row_norm = 0.f;                // ... and the largest row L2 norm (error bound of the bf16 sweep)
    // TMA descriptors (CUtensorMap, kept as bytes: this header does not see the driver API) of `dense` with full- and
    // half-height boxes, encoded once at build; the one of the converted query block is re-encoded only when the
    // caller's workspace or the batch height changes
    alignas(64) unsigned char tmap_x[128] = {}, tmap_x_half[128] = {}, tmap_q[128] = {};
    const void *tmap_q_ptr = nullptr;
    int64_t tmap_q_rows = 0;

    int64_t device_bytes = 0;
    int64_t stream_bytes = 0;
    unsigned long long *scan_prof = nullptr;   // diagnostic (vs_debug_scan_profile): phase timestamps of the binary scan

    // ---- timing hook
    cudaEvent_t ev0[VS_TIMER_SLOTS] = {}, ev1[VS_TIMER_SLOTS] = {};
    int last_mode = VS_MODE_SCAN;  // kernel family the last search used, when the host chose it (forced scan / dense)
    bool last_mode_on_device = false;   // the last search let the device decide: the answer is in *d_last_mode
    int *d_last_mode = nullptr;    // device int written by inv_decide_kernel
    int timer_n = 0;         // launches recorded since the last reset (may exceed the ring)
};

// Plain-pointer view of an index passed BY VALUE to kernels (vs_index itself carries host-only state).
struct WsView {
    const uint4 *cols; const void *vals; const uint32_t *tails;
    const uint32_t *part_win_begin; const uint32_t *part_row_begin; const uint32_t *row_chunk;
    int64_t n_rows, n_cols;
    int kind, store_dtype, cpl_shift;
};
// logical chunk index (row order) -> physical chunk position in cols / vals: inside its step of 32 << cpl_shift
// chunks, logical l*C+i is stored at i*32+l
__host__ __device__ __forceinline__ uint64_t ws_phys_chunk(uint64_t c, int cpl_shift) {
    const uint64_t m = (32ull << cpl_shift) - 1ull, r = c & m;
    return (c & ~m) | ((r & ((1ull << cpl_shift) - 1ull)) << 5) | (r >> cpl_shift);
}
inline WsView ws_view(const vs_index *i) {
    WsView v;
    v.cols = i->cols; v.vals = i->vals; v.tails = i->tails;
    v.part_win_begin = i->part_win_begin; v.part_row_begin = i->part_row_begin; v.row_chunk = i->row_chunk;
    v.n_rows = i->n_rows; v.n_cols = i->n_cols; v.kind = i->kind; v.store_dtype = i->store_dtype;
    v.cpl_shift = i->cpl_shift;
    return v;
}

// CSR-style sparse queries on the device: query b of a launch = entries [ptr[b0 + b], ptr[b0 + b + 1]) of tok / w
struct SparseQueries {
    const void *ptr; int ptr_dtype;    // VS_I32 | VS_I64 offsets
    const int32_t *tok; const float *w;
    int64_t b0;
};

namespace vs {
int build_ws_index(vs_index *idx, const void *d_crow, int crow_dtype, const void *d_col, int col_dtype,
                   const void *d_val, int val_dtype, cudaStream_t st, int64_t col_shift = 0);
// index from CSR arrays already on the device (col: VS_I32 | VS_I64 | VS_U16); columns < col_shift are dropped and the
// rest renumbered (n_cols = width AFTER the shift).  SYNC.
int create_csr_from_device(int device, int64_t n_rows, int64_t n_cols, int64_t nnz, const void *d_crow, int crow_dtype,
                           const void *d_col, int col_dtype, const void *d_val, int val_dtype, int store_dtype,
                           int64_t col_shift, cudaStream_t st, vs_index **out);
int export_ws_csr(const vs_index *idx, int64_t *d_crow, int64_t *d_col, float *d_val, cudaStream_t st);
int build_inverted(vs_index *idx, cudaStream_t st);
int debug_gather_wavefronts(const vs_index *idx, unsigned long long *d_out2, cudaStream_t st);
int build_dense_index(vs_index *idx, const void *d_x, int x_dtype, int64_t ld, cudaStream_t st);
}  // namespace vs
