// bot_build.cu -- bag-of-token index rows from token-id batches, on the GPU.
// Replaces upstream Retriever._build_bot_vectors (src/ir/retriever/retriever.py:208-253): a dense [batch, vocab] fp16
// matrix per batch, `emb[i, token_ids] = 1`, the `[:, num_shift:]` column slice, to_sparse_coo, cat, to_sparse_csr --
// 1,756 s for the 21M-passage corpus (build_binary_token_index.sh:10).  Here: one warp per passage marks its token ids
// in a vocabulary bitmap in shared memory (all of them, or the first `max_token` DISTINCT ones in sequence order, the
// reference's get_first_unique_n) and emits the set bits >= num_shift as ascending column ids `id - num_shift`.
// Two passes with the same kernel: count (row lengths) -> caller's exclusive scan -> fill.
#include "index.cuh"

namespace vs {

constexpr int kBotWarps = 8;

__device__ __forceinline__ int64_t load_token(const void *ids, int dtype, int64_t i) {
    return dtype == VS_I64 ? ((const int64_t *)ids)[i] : (int64_t)((const int32_t *)ids)[i];
}

template <bool FILL>
__global__ void __launch_bounds__(kBotWarps * 32) bot_rows_kernel(const void *ids, int ids_dtype, int64_t n_rows, int64_t ld,
                                                                  const int32_t *lengths, int vocab, int shift, int max_token,
                                                                  int64_t *row_nnz_or_crow, int32_t *col) {
    extern __shared__ uint32_t s_bits[];   // kBotWarps bitmaps of `words` words
    const int words = (vocab + 31) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *bits = s_bits + (size_t)warp * words;
    const uint32_t lt = (1u << lane) - 1u;
    for (int64_t r = (int64_t)blockIdx.x * kBotWarps + warp; r < n_rows; r += (int64_t)gridDim.x * kBotWarps) {
        for (int i = lane; i < words; i += 32) bits[i] = 0u;
        __syncwarp();
        int64_t len = lengths ? (int64_t)lengths[r] : ld;
        if (len > ld) len = ld;
        if (max_token <= 0) {
            for (int64_t t = lane; t < len; t += 32) {
                const int64_t id = load_token(ids, ids_dtype, r * ld + t);
                if (id >= 0 && id < vocab) atomicOr(&bits[id >> 5], 1u << (id & 31));
            }
        } else if (lane == 0) {   // first `max_token` distinct ids in sequence order
            int taken = 0;
            for (int64_t t = 0; t < len && taken < max_token; ++t) {
                const int64_t id = load_token(ids, ids_dtype, r * ld + t);
                if (id < 0 || id >= vocab) continue;
                const uint32_t b = 1u << (id & 31);
                if (!(bits[id >> 5] & b)) { bits[id >> 5] |= b; ++taken; }
            }
        }
        __syncwarp();
        // lane l owns the contiguous word range [l * per, (l + 1) * per): ascending ids across lanes
        const int per = (words + 31) >> 5;
        const int w0 = lane * per, w1 = min(words, w0 + per);
        const int first_word = shift >> 5;
        int mine = 0;
        for (int i = w0; i < w1; ++i) {
            uint32_t b = bits[i];
            if (i < first_word) b = 0u;
            else if (i == first_word) b &= ~((1u << (shift & 31)) - 1u);
            mine += __popc(b);
        }
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if constexpr (!FILL) {
            if (lane == 31) row_nnz_or_crow[r] = incl;
        } else {
            int64_t pos = row_nnz_or_crow[r] + (incl - mine);
            for (int i = w0; i < w1; ++i) {
                uint32_t b = bits[i];
                if (i < first_word) b = 0u;
                else if (i == first_word) b &= ~((1u << (shift & 31)) - 1u);
                while (b) {
                    const int bit = __ffs(b) - 1;
                    b &= b - 1;
                    col[pos++] = (int32_t)(i * 32 + bit - shift);
                }
            }
        }
        __syncwarp();
        (void)lt;
    }
}

}  // namespace vs

using namespace vs;

extern "C" int vs_bot_from_tokens(int device, const void *d_token_ids, int ids_dtype, int64_t n_rows, int64_t ld,
                                  const int32_t *d_lengths, int vocab_size, int num_shift, int max_token,
                                  int64_t *d_row_nnz_or_crow, int32_t *d_col, void *stream) {
    VS_REQUIRE(d_token_ids && d_row_nnz_or_crow && n_rows >= 0 && ld > 0, VS_ERR_INVALID, "vs_bot_from_tokens: bad argument");
    VS_REQUIRE(ids_dtype == VS_I32 || ids_dtype == VS_I64, VS_ERR_INVALID, "token ids must be int32 or int64");
    VS_REQUIRE(vocab_size > 0 && num_shift >= 0 && num_shift < vocab_size, VS_ERR_INVALID, "need 0 <= num_shift < vocab_size");
    VS_CUDA(cudaSetDevice(device));
    const size_t smem = (size_t)kBotWarps * ((vocab_size + 31) / 32) * 4;
    VS_REQUIRE(smem <= 200 * 1024, VS_ERR_UNSUPPORTED, "vocabulary of %d tokens does not fit the bitmap kernel", vocab_size);
    if (n_rows == 0) return VS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n_rows + kBotWarps - 1) / kBotWarps;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (d_col == nullptr) {
        VS_CUDA(cudaFuncSetAttribute(bot_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bot_rows_kernel<false><<<(unsigned)blocks, kBotWarps * 32, smem, st>>>(d_token_ids, ids_dtype, n_rows, ld, d_lengths, vocab_size,
                                                                           num_shift, max_token, d_row_nnz_or_crow, nullptr);
    } else {
        VS_CUDA(cudaFuncSetAttribute(bot_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bot_rows_kernel<true><<<(unsigned)blocks, kBotWarps * 32, smem, st>>>(d_token_ids, ids_dtype, n_rows, ld, d_lengths, vocab_size,
                                                                          num_shift, max_token, d_row_nnz_or_crow, d_col);
    }
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}
