/* vsearch_b200.h -- C ABI of the B200-native index-scoring engine.
 *
 * The upstream project (jzhoubu/vsearch) is pure Python and has no FFI of its
 * own; the narrowest seam is the `Index` class family in
 * src/ir/retriever/index.py.  Each entry point below names the reference
 * interface it replaces (paths relative to the upstream repo).  The Python
 * mirror of those classes (vsearch_b200/index.py) binds this ABI with ctypes;
 * INTEGRATION.md shows the stub a maintainer would add upstream.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *  - every function returns a vs_status (0 = OK) and never throws; the message
 *    of the last failure on the calling thread is vs_last_error();
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the
 *    functions do not synchronise it unless stated ("SYNC"): a search on a sparse /
 *    bag-of-token index returns as soon as its kernels are enqueued -- host inputs are
 *    staged through the caller's workspace (no allocation; pageable memory is consumed
 *    before the call returns, pinned memory must stay valid until the stream has passed
 *    the copy), and the scan-vs-inverted choice of the auto mode is made on the device;
 *  - pointers named d_* must be device pointers on the index's device; pointers
 *    named hd_* may be host or device (the library checks);
 *  - a handle may be used from one host thread at a time.
 */
#ifndef VSEARCH_B200_H
#define VSEARCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_ABI_VERSION 2

typedef enum {
    VS_OK = 0,
    VS_ERR_INVALID = 1,      /* bad argument (also: k > N, the reference's RuntimeError, index.py:92) */
    VS_ERR_UNSUPPORTED = 2,  /* valid request this build cannot serve (e.g. V too large for shared memory) */
    VS_ERR_CUDA = 3,         /* a CUDA runtime call failed; see vs_last_error() */
    VS_ERR_NOMEM = 4
} vs_status;

typedef enum {
    VS_F32 = 0, VS_F16 = 1, VS_BF16 = 2, VS_I32 = 3, VS_I64 = 4, VS_U16 = 5, VS_U32 = 6, VS_NONE = 7,
    VS_F64 = 8   /* .npz members only (scipy's default value dtype): converted to f32 / f16 while loading */
} vs_dtype;

/* which kernel family serves vs_search on a sparse / binary index */
typedef enum {
    VS_MODE_AUTO = 0,     /* engine picks scan vs inverted from the query sparsity */
    VS_MODE_SCAN = 1,     /* K1/K2: passage-major stream, query vector in shared memory */
    VS_MODE_INVERTED = 2  /* K3: token-major posting lists of the query's non-zero tokens */
} vs_mode;

typedef struct vs_index vs_index; /* opaque; owns device memory */

const char *vs_last_error(void);
int vs_abi_version(void);

/* ---- index construction --------------------------------------------------------------
 * Replaces SparseIndex._scipy_csr_to_torch_csr + `.to(device)` (index.py:144-161,179) and
 * Index.move_to_device (index.py:54-57): takes the CSR triple of the [N, V] index
 * (crow: N+1 entries, col/val: nnz entries; hd_* = host or device memory) and builds the
 * compact device format (uint16 columns in 16-byte chunks, optional values, per-warp
 * stream partition).  val_dtype == VS_NONE (hd_val == NULL) builds the binary
 * bag-of-token index (BoTIndex, index.py:205-218; all values 1).  store_dtype picks the
 * on-device value type: VS_F32 / VS_F16 / VS_BF16 (ignored for binary).  SYNC. */
int vs_index_create_csr(int device, int64_t n_rows, int64_t n_cols, int64_t nnz,
                        const void *hd_crow, int crow_dtype,   /* VS_I32 | VS_I64 */
                        const void *hd_col, int col_dtype,     /* VS_I32 | VS_I64 */
                        const void *hd_val, int val_dtype,     /* VS_F32 | VS_F16 | VS_BF16 | VS_NONE */
                        int store_dtype, void *stream, vs_index **out);

/* Dense index: `Index.vector` strided [N, D] (index.py:25-44,88-94).  store_dtype VS_BF16 | VS_F16: the vectors are
 * scored as stored on the tensor cores (fp32 accumulate).  VS_F32 (upstream's Index(fp16=False): an fp32 matrix and an
 * fp32 GEMM): fp32 semantics -- the tensor cores sweep a bf16 copy, every passage within the bf16 error bound
 * 2 * 2^-7 * |q| * max|x| of the k-th score is re-scored exactly from its fp32 row, and those scores are ranked; the
 * index keeps both copies (6 bytes per element).  SYNC. */
int vs_index_create_dense(int device, int64_t n_rows, int64_t dim, const void *hd_x, int x_dtype,
                          int64_t ld, int store_dtype, void *stream, vs_index **out);

int vs_index_destroy(vs_index *idx);

/* shape / layout queries (Index.__str__, index.py:117-126) */
int vs_index_info(const vs_index *idx, int64_t *n_rows, int64_t *n_cols, int64_t *nnz,
                  int *kind /* 0 dense, 1 sparse, 2 binary */, int *store_dtype, int64_t *device_bytes,
                  int64_t *stream_bytes /* bytes one scan pass reads */);

/* Export the index back to a CSR triple (for SparseIndex.save, index.py:181-202).
 * d_crow int64[N+1], d_col int64[nnz], d_val float32[nnz] device buffers. */
int vs_index_export_csr(const vs_index *idx, int64_t *d_crow, int64_t *d_col, float *d_val, void *stream);

/* ---- search ---------------------------------------------------------------------------
 * Replaces Index.search (index.py:88-94): scores = q @ vector.t(); scores.topk(k).
 *   hd_q      [B, ldq] queries (host or device), q_dtype VS_F32 | VS_F16 | VS_BF16
 *   k         1 <= k <= min(N, VS_MAX_K); k > N -> VS_ERR_INVALID (reference: RuntimeError)
 *   d_ids     int64 [B, k]  out: passage ids + id_offset, ranked (score desc, id asc)
 *   d_scores  float [B, k]  out: scores (fp32 accumulate)
 *   score_round  VS_F32 none | VS_F16 | VS_BF16: round scores to that type BEFORE ranking,
 *             mirroring `q.type(vector.dtype)` / scores in the index dtype (index.py:89)
 *   d_workspace  >= vs_search_workspace_bytes(idx, B, k) bytes, 256-byte aligned
 * The [B, N] score matrix is never written. */
/* Synchronisation: sparse / bag-of-token indices -- none (the first auto / inverted search on a handle builds the
 * inverted lists once: SYNC that one time).  Dense index -- the whole call is enqueued without reading anything back,
 * then SYNC once at its end: one status word says whether a survivor list overflowed (adversarial row order), in which
 * case the call is redone with per-sweep checks. */
#define VS_MAX_K 2048
size_t vs_search_workspace_bytes(const vs_index *idx, int64_t B, int k);
int vs_search(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, int k, int mode,
              int score_round, int64_t id_offset, int64_t *d_ids, float *d_scores,
              void *d_workspace, size_t workspace_bytes, void *stream);

/* Same search, but returns the packed 64-bit rank keys (ordered score bits << 32 | ~global id),
 * sorted descending: what each rank contributes to the one all-gather of the row-sharded path
 * (no reference counterpart: upstream searches one device, index.py:179). */
int vs_search_keys(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, int k, int mode,
                   int score_round, int64_t id_offset, uint64_t *d_keys,
                   void *d_workspace, size_t workspace_bytes, void *stream);

/* Row-sharded DENSE search, one rank's part, in three steps with an all-gather of the [B, k] keys after each (no
 * reference counterpart: upstream searches one device, index.py:179; sharded.py drives it).  A rank on its own learns
 * its thresholds from its own rows only; pooled, the ranks get the thresholds of the whole index from a sample 1/n_ranks
 * the size each:
 *   step 0  sample sweep over this rank's share of the sample prefix   -> d_keys: its top-k (global ids)
 *   step 1  d_gathered_keys = all ranks' step-0 keys [n_ranks, B, k]; filtered sweep over the next 64 x share rows
 *                                                                      -> d_keys: top-k of what this rank has seen
 *   step 2  the same with the step-1 keys, over the rest of the rows   -> d_keys: this rank's final top-k;
 *           *d_status (device uint32, may be NULL) != 0: a survivor list overflowed, redo the search through vs_search_keys.
 * Same workspace pointer (vs_search_workspace_bytes(idx, B, k)) in all three steps: it carries the state.  Device
 * queries, B <= 4096, 16-bit index.  No synchronisation. */
int vs_search_dense_step(const vs_index *idx, int step, const void *d_q, int q_dtype, int64_t B, int64_t ldq, int k,
                         int score_round, int64_t id_offset, int n_ranks, const uint64_t *d_gathered_keys, uint64_t *d_keys,
                         uint32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream);

/* Same search for SPARSE queries given as CSR-style (token, weight) lists -- what the reference's query sparsifier
 * produces (utils/sparse.py:8-19: the a=768 activation budget, encoder/vdr.py:159-169) -- without the dense [B, V]
 * detour: query b = sum_j hd_qw[j] * e(hd_qtok[j]) over j in [hd_qptr[b], hd_qptr[b+1]).
 *   hd_qptr   B+1 offsets (VS_I32 | VS_I64), hd_qtok int32 columns, hd_qw float weights; all host or all device
 *   tokens outside [0, n_cols) and zero weights are ignored, duplicate tokens of a query add up
 *   outputs: d_ids + d_scores (as vs_search) and / or d_keys (as vs_search_keys); unused ones NULL
 * 64 (token, weight) pairs are 512 bytes against the 118 KB of a dense fp32 query row.  The inverted lists take the
 * pairs as they are; the scan kernels scatter them into their shared-memory query vector.  Sparse / bag-of-token
 * indices only.  No synchronisation. */
int vs_search_sparse(const vs_index *idx, const void *hd_qptr, int ptr_dtype, const int32_t *hd_qtok, const float *hd_qw,
                     int64_t B, int k, int mode, int score_round, int64_t id_offset, int64_t *d_ids, float *d_scores,
                     uint64_t *d_keys, void *d_workspace, size_t workspace_bytes, void *stream);

/* Diagnostic: the dense score matrix itself (upstream index.py:91), d_scores_full float [B, N].
 * Not used by the search path; lets tests separate scoring errors from selection errors. */
int vs_scores(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, int score_round,
              float *d_scores_full, void *d_workspace, size_t workspace_bytes, void *stream);

/* Merge P candidate lists per query (e.g. the all-gathered [P, B, k] keys of P shards) into the
 * global top-k.  d_keys_in[p * stride_p + b * stride_b + j], j < k_in. */
int vs_merge_keys(int device, const uint64_t *d_keys_in, int64_t P, int64_t stride_p, int64_t stride_b,
                  int64_t B, int k_in, int k_out, int64_t *d_ids, float *d_scores, void *stream);

/* Rerank stage: scores of GIVEN rows, d_scores[b, j] = <q_b, row d_ids[b, j]> (ids outside [0, N) -> -inf).
 * Replaces the re-embedding + bmm of the reference's `retrieve(rerank=True)` (src/ir/retriever/retriever.py:137-141)
 * when the candidates' parametric vectors already sit in a device-resident sparse index: B*k row gathers instead of
 * an encoder pass.  Sparse / bag-of-token indices.  Workspace >= vs_score_rows_workspace_bytes(idx, B) (the whole batch
 * is prepared at once: not covered by vs_search_workspace_bytes for large B).  No synchronisation. */
size_t vs_score_rows_workspace_bytes(const vs_index *idx, int64_t B);
int vs_score_rows(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, const int64_t *d_ids, int k,
                  int score_round, float *d_scores, void *d_workspace, size_t workspace_bytes, void *stream);

/* which kernel family (VS_MODE_SCAN | VS_MODE_INVERTED) served the last search on this handle.  SYNC when that search
 * let the device choose (auto / inverted): the decision is read back.  A forced VS_MODE_INVERTED whose queries the
 * lists cannot serve (> 4096 non-zeros or >= 2^32 postings in one query) is answered by the scan and reports so. */
int vs_index_last_mode(const vs_index *idx, int *mode);

/* timing hook for bench.py: every scan/score kernel launch made through this handle is bracketed by a
 * CUDA event pair on the launching stream (a ring of VS_TIMER_SLOTS pairs).  Returns the summed device
 * time (ms) and the number of launches recorded since the last reset; reset != 0 clears the ring
 * afterwards.  SYNC (waits for the recorded events). */
#define VS_TIMER_SLOTS 256
int vs_kernel_timer(vs_index *idx, int reset, float *total_ms, int *launches);

/* Diagnostic: while d_buf != NULL the binary scan kernel (K2) stores %globaltimer at its phase boundaries,
 * d_buf[(cta * B + query) * 8 + slot] (slots: 0 pass start, 1 query vector staged, 2 sampling streamed, 3 threshold
 * selected, 4 warp 0 finished streaming, 5 all warps finished, 6 top-k written); the buffer must hold
 * n_ctas * B * 8 entries for every search issued while it is set.  The inverted-list kernel (K3) stores phase DURATIONS
 * instead, d_buf[(query * n_ctas + cta) * 16 + slot] in ns (slots: 0 setup, 1 zero the accumulator, 2 accumulate, 3
 * first-block histogram, 4 block select, 5 refresh/compact, 6 final write, 7 total, 8-10 the first block's select split
 * into first rows / refreshes / rest), so the buffer must hold n_ctas * B * 16 entries there.  NULL switches it off. */
int vs_debug_scan_profile(vs_index *idx, unsigned long long *d_buf);

/* Diagnostic: what the bank-aware entry placement achieved on this sparse / binary index.  d_out2[0] = shared-memory
 * wavefronts the scan's query gathers cost per pass (largest number of distinct addresses on one bank, summed over the
 * gather instructions), d_out2[1] = gather instructions per pass (device uint64[2]).  1.0 per gather is conflict-free. */
int vs_debug_gather_wavefronts(const vs_index *idx, unsigned long long *d_out2, void *stream);

/* ---- query sparsifier ---------------------------------------------------------------------------------------------
 * In place on a device fp32 batch d_q [B, ld]: keep the k largest of the first n_cols entries of every row (ties ->
 * lower column) and zero the rest -- upstream utils/sparse.py:8-19 (build_topk_mask / topk_sparsify), the `a=768`
 * activation budget of retrieve (retriever.py:134).  d_bow_ids (optional, int32 [B, bow_ld], values id - bow_shift
 * outside [0, n_cols) ignored): the row's own token columns survive as well, the logical_or(bow_mask, topk_mask) of
 * encoder/vdr.py:159-169.  k = 0 keeps only those; k >= n_cols keeps everything. */
int vs_sparsify_topk(int device, float *d_q, int64_t B, int64_t ld, int n_cols, int k, const int32_t *d_bow_ids, int bow_ld,
                     int bow_shift, void *stream);

/* ---- dense -> CSR on the GPU -----------------------------------------------------------------------------------------
 * The non-zeros of a dense device matrix d_x [n_rows, ld] (f32 | f16 | bf16; first n_cols columns) as a CSR triple,
 * columns ascending.  Replaces `vectors.to_sparse_csr()` of the reference's Retriever.build_index
 * (src/ir/retriever/retriever.py:299-305), and turns a sparsified query batch (vs_sparsify_topk) into the
 * (token, weight) lists vs_search_sparse takes.  Two calls, like vs_bot_from_tokens: d_col == NULL writes the row
 * lengths into d_row_nnz_or_crow[n_rows]; the caller scans them into row pointers and calls again with
 * d_row_nnz_or_crow = crow (int64 [n_rows + 1]), d_col (int32 [nnz]) and d_val (float [nnz]).  No synchronisation. */
int vs_dense_to_csr(int device, const void *d_x, int x_dtype, int64_t n_rows, int64_t ld, int n_cols,
                    int64_t *d_row_nnz_or_crow, int32_t *d_col, float *d_val, void *stream);

/* ---- bag-of-token rows from token-id batches, on the GPU -----------------------------------------------------------
 * Replaces Retriever._build_bot_vectors (reference src/ir/retriever/retriever.py:208-253: dense [batch, vocab]
 * scatter of ones, `[:, num_shift:]`, to_sparse_coo / cat / to_sparse_csr).  Row r of the result = the distinct
 * token ids of passage r (all of them, or the first `max_token` distinct ones in sequence order when max_token > 0,
 * reference get_first_unique_n), minus ids < num_shift, renumbered id - num_shift, ascending.
 *   d_token_ids  device [n_rows, ld] int32 | int64 (ids outside [0, vocab_size) are ignored)
 *   d_lengths    device int32 [n_rows] valid ids per row, or NULL (all ld)
 * Two calls: d_col == NULL writes the row lengths into d_row_nnz_or_crow[n_rows]; the caller scans them into row
 * pointers and calls again with d_row_nnz_or_crow = crow (int64 [n_rows + 1]) and d_col (int32 [nnz]) to fill.
 * Enqueues on `stream`, does not synchronise. */
int vs_bot_from_tokens(int device, const void *d_token_ids, int ids_dtype, int64_t n_rows, int64_t ld,
                       const int32_t *d_lengths, int vocab_size, int num_shift, int max_token,
                       int64_t *d_row_nnz_or_crow, int32_t *d_col, void *stream);

/* ---- native .npz shard reader (host only, no CUDA calls) ------------------------------------------------------
 * Replaces, for the big members of an index shard, scipy.sparse.load_npz + vstack + astype on one Python thread
 * (reference src/ir/retriever/index.py:172-176).  A shard is a zip of .npy members written by
 * scipy.sparse.save_npz (index.py:195-197): indices / indptr (int32 | int64), data (float32 | float16), plus the
 * small format / shape / _is_array members.  Calls on one handle may run concurrently from several threads (each
 * opens its own file descriptor), which is how the loader inflates every member of every shard in parallel. */
typedef struct vs_npz vs_npz;
int vs_npz_open(const char *path, vs_npz **out);
int vs_npz_close(vs_npz *z);
/* name without ".npy".  dtype = VS_* code (VS_NONE for dtypes the search path does not use); shape4 gets up to
 * four extents (0-filled); n_elems = product of the extents. */
int vs_npz_member_info(vs_npz *z, const char *name, int *dtype, int *ndim, int64_t *shape4, int64_t *n_elems);
/* Inflate elements [skip_elems, skip_elems + n_elems) of member `name` straight into dst (host memory), converted
 * to dst_dtype (integer -> VS_I32 | VS_I64 | VS_U32 | VS_U16 with add_offset added: the row-pointer offset of a
 * shard when shards are concatenated by rows; float -> VS_F32 | VS_F16: the reference's fp16=True astype).
 * Streaming: 2 MB of scratch per call, no inflated copy of the member. */
int vs_npz_read(vs_npz *z, const char *name, void *dst, int dst_dtype, int64_t skip_elems, int64_t n_elems,
                int64_t add_offset);

/* Loader: row shards written by scipy.sparse.save_npz (reference SparseIndex.save, index.py:195-197) -> one device
 * index, replacing SparseIndex.init_index (index.py:163-179: glob -> load_npz -> [:, shift:] -> vstack -> astype ->
 * torch CSR -> .to(device)) without any host copy of the index.  `paths` are concatenated by rows in the order given
 * (the caller sorts them lexicographically like upstream: index10 before index2).  Every (shard, member) pair is
 * inflated by one of `threads` host threads (<= 0: all cores), narrowed on the fly (columns -> uint16 when the file's
 * vocabulary fits, values -> float16 when value_dtype == VS_F16 -- upstream's fp16=True --, row pointers offset by the
 * shards before) into pinned staging buffers and copied asynchronously to its place in the device CSR arrays.
 *   shift            columns < shift are dropped and the rest renumbered (done by the index build on the GPU)
 *   value_dtype      device value type of a valued index: VS_F32 | VS_F16 | VS_BF16
 *   binary_if_ones   != 0: an all-ones `data` member gives the binary bag-of-token index (BoTIndex)
 *   shard_rows       optional out, int64[n_paths]: rows of each file (global id offsets of row-sharded ranks)
 * A rank of the row-sharded layout passes only ITS files (upstream writes one file per shard,
 * examples/inference_sparse/README.md:86-107).  SYNC. */
int vs_index_load_npz(int device, const char *const *paths, int n_paths, int shift, int value_dtype, int binary_if_ones,
                      int threads, void *stream, vs_index **out, int64_t *shard_rows);

/* Writer: a zip of deflated .npy members, byte-compatible with scipy.sparse.save_npz / numpy.savez_compressed
 * (reference SparseIndex.save, index.py:195-197), compressed by `threads` threads (independent 4 MB deflate blocks
 * concatenated into one valid stream).  The caller supplies each member's .npy header bytes (numpy.lib.format) and
 * its raw C-order data; members are written in the order given.  ZIP64 when a member or offset needs it. */
typedef struct vs_npz_member_in {
    const char *name;            /* without ".npy" */
    const void *header; int64_t header_bytes;
    const void *data; int64_t data_bytes;
} vs_npz_member_in;
int vs_npz_write(const char *path, const vs_npz_member_in *members, int n_members, int level, int threads);

#ifdef __cplusplus
}
#endif
#endif /* VSEARCH_B200_H */
