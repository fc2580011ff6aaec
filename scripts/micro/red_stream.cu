// micro-benchmark: why is streaming an accumulator row that was just updated by RED.ADD.F32 slow?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
__global__ void red_kernel(float* acc, const uint32_t* doc, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x+threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x) atomicAdd(acc + doc[i], 1.0f);
}
__global__ void gen_kernel(uint32_t* doc, int64_t n, uint32_t N, uint32_t seed) {
    for (int64_t i = (int64_t)blockIdx.x*blockDim.x+threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x) {
        uint64_t x = (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull + seed; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        doc[i] = (uint32_t)(x % N);
    }
}
template <bool ZERO, int U>
__global__ void __launch_bounds__(1024,1) stream_kernel(float* acc, int64_t n_rows, int rows_per_cta, float* out) {
    float4* a4 = (float4*)acc;
    const int64_t rb = (int64_t)blockIdx.x * rows_per_cta, re = min(n_rows, rb + rows_per_cta);
    float m = 0.f;
    for (int64_t r = rb + (int64_t)threadIdx.x * 4; r < re; r += 1024 * 4 * U) {
        float4 v[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) { int64_t ru = r + (int64_t)u*4096; v[u] = ru < re ? a4[ru >> 2] : make_float4(0,0,0,0); }
        #pragma unroll
        for (int u = 0; u < U; ++u) { int64_t ru = r + (int64_t)u*4096; if (ZERO && ru < re) a4[ru >> 2] = make_float4(0,0,0,0);
            m = fmaxf(m, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w))); }
    }
    if (m == 12345.f) out[0] = m;
}
int main() {
    const int64_t N = 21015324, P = 5500000; const int ctas = 148;
    int rpc = (int)((N + ctas - 1) / ctas); rpc = (rpc + 3) / 4 * 4;
    float *acc, *out; uint32_t* doc;
    CK(cudaMalloc(&acc, (N + 4) * 4)); CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&doc, P * 4));
    gen_kernel<<<1024, 256>>>(doc, P, (uint32_t)N, 7);
    CK(cudaMemset(acc, 0, (N + 4) * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto fn, int reps) { float best = 1e9, tot = 0; for (int i = 0; i < reps; ++i) { cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; tot += ms; } printf("%-40s best %.1f us  avg %.1f us\n", name, best * 1e3, tot / reps * 1e3); };
    timeit("read only U=2 (no REDs before)", [&]{ stream_kernel<false,2><<<ctas,1024>>>(acc, N, rpc, out); }, 5);
    timeit("read+zero U=2 (no REDs before)", [&]{ stream_kernel<true,2><<<ctas,1024>>>(acc, N, rpc, out); }, 5);
    timeit("read+zero U=4", [&]{ stream_kernel<true,4><<<ctas,1024>>>(acc, N, rpc, out); }, 5);
    timeit("RED 5.5M random", [&]{ red_kernel<<<592,256>>>(acc, doc, P); }, 5);
    for (int i = 0; i < 3; ++i) {
        timeit("  RED then..", [&]{ red_kernel<<<592,256>>>(acc, doc, P); }, 1);
        timeit("  ..read+zero U=2 after RED", [&]{ stream_kernel<true,2><<<ctas,1024>>>(acc, N, rpc, out); }, 1);
    }
    for (int i = 0; i < 3; ++i) {
        timeit("  RED then..", [&]{ red_kernel<<<592,256>>>(acc, doc, P); }, 1);
        timeit("  ..read only U=2 after RED", [&]{ stream_kernel<false,2><<<ctas,1024>>>(acc, N, rpc, out); }, 1);
        timeit("  ..zero pass", [&]{ stream_kernel<true,2><<<ctas,1024>>>(acc, N, rpc, out); }, 1);
    }
    timeit("memset 84MB", [&]{ cudaMemsetAsync(acc, 0, N * 4); }, 5);
    // many small CTAs variant
    timeit("read+zero U=2, 1184 CTAs x 1024", [&]{ stream_kernel<true,2><<<1184,1024>>>(acc, N, (int)(((N + 1183) / 1184 + 3) / 4 * 4), out); }, 5);
    return 0;
}
