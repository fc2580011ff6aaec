// index.cuh -- device-resident index formats.
//
// Sparse / binary ("warp-stream", WS) format.  One scan pass reads it front to back:
//   * entries are uint16 column ids packed 8 to a 16-byte CHUNK; every row owns
//     ceil(len/8) chunks (>= 1), padded with the sentinel column V whose query slot is 0;
//   * 32 consecutive chunks form a WINDOW = one coalesced 512-byte warp load;
//   * the stream is cut into n_parts contiguous PARTS (one per resident warp of the
//     persistent scan grid: #SMs x 32), each a whole number of windows holding whole rows,
//     balanced by chunk count;
//   * tails[w] has bit i set when chunk i of window w is the LAST chunk of its row: the
//     scan kernel turns per-lane partial dot products into row scores with one segmented
//     warp scan and no row-pointer lookups;
//   * values (fp32 / fp16 / bf16) sit in a parallel array with the same chunk geometry;
//     the binary bag-of-token index has none.
// Algorithmic bytes of a pass: nnz*(2+b_val) + (N+1)*4.  The format streams
// n_windows*32*(16 + 8*b_val) + n_windows*4 bytes; the overhead is the row padding (<= 7
// entries per row) and is reported by vs_index_info(stream_bytes).
#pragma once
#include "common.cuh"

struct vs_index {
    int device = 0;
    int kind = 0;         // 0 dense, 1 sparse (valued), 2 binary
    int store_dtype = VS_F32;
    int64_t n_rows = 0, n_cols = 0, nnz = 0;

    // ---- WS format
    int n_ctas = 0;           // persistent scan grid (= #SMs at build time)
    int warps_per_cta = 32;
    int n_parts = 0;          // n_ctas * warps_per_cta
    uint64_t n_windows = 0;
    uint4 *cols = nullptr;            // n_windows * 32 chunks
    void *vals = nullptr;             // f32: 2 x uint4 per chunk; f16/bf16: 1 x uint4 per chunk
    uint32_t *tails = nullptr;        // n_windows
    uint32_t *part_win_begin = nullptr;  // n_parts + 1
    uint32_t *part_row_begin = nullptr;  // n_parts + 1
    uint32_t *row_chunk = nullptr;       // N + 1: first chunk of each row in the stream (export, rerank)

    // ---- dense
    int64_t dim = 0;
    void *dense = nullptr;

    int64_t device_bytes = 0;
    int64_t stream_bytes = 0;

    // ---- timing hook
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int last_launches = 0;
    bool timed = false;
};

namespace vs {
int build_ws_index(vs_index *idx, const void *d_crow, int crow_dtype, const void *d_col, int col_dtype,
                   const void *d_val, int val_dtype, cudaStream_t st);
int export_ws_csr(const vs_index *idx, int64_t *d_crow, int64_t *d_col, float *d_val, cudaStream_t st);
}  // namespace vs
