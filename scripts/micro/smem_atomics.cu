// smem_atomics.cu -- ceiling of the inverted-list accumulate (K3): shared-memory read-modify-write operations per
// second, one persistent CTA of 768 threads per SM, random rows of a 36,864-row accumulator (144 KB), 8 independent
// operations in flight per thread -- the shape of accumulate_slice() in csrc/inverted.cu without its posting loads.
//   variant 0  atomicAdd(float)      ATOMS.CAST.SPIN loop (what K3 issues: sm_100a has no native fp32 shared atomic add)
//   variant 1  atomicAdd(uint32)     ATOMS.ADD
//   variant 2  atomicAdd(uint64)     ATOMS.ADD.64 (8-byte accumulators: 18,432 rows)
//   variant 3  atomicOr(uint32)      ATOMS.OR
//   variant 4  plain load-add-store  (racy upper bound: LDS + FADD + STS)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu ; run: ./smem_atomics
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int kThreads = 768, kRows = 36864, kUnroll = 8;

template <int VAR>
__global__ void __launch_bounds__(kThreads, 1) rmw_kernel(int iters, unsigned long long *sink) {
    extern __shared__ __align__(16) uint8_t smem[];
    float *accf = reinterpret_cast<float *>(smem);
    uint32_t *accu = reinterpret_cast<uint32_t *>(smem);
    unsigned long long *acc64 = reinterpret_cast<unsigned long long *>(smem);
    for (int i = threadIdx.x; i < kRows; i += kThreads) accu[i] = 0;
    __syncthreads();
    uint32_t x = 0x9e3779b9u * (blockIdx.x * kThreads + threadIdx.x + 1);
    constexpr uint32_t rows = VAR == 2 ? kRows / 2 : kRows;
    for (int it = 0; it < iters; ++it) {
        uint32_t r[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            x = x * 1664525u + 1013904223u;
            r[u] = (uint32_t)(((uint64_t)x * rows) >> 32);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (VAR == 0) atomicAdd(&accf[r[u]], 1.25f);
            else if (VAR == 1) atomicAdd(&accu[r[u]], 5u);
            else if (VAR == 2) atomicAdd(&acc64[r[u]], 5ull);
            else if (VAR == 3) atomicOr(&accu[r[u]], 1u << (u & 31));
            else accf[r[u]] = accf[r[u]] + 1.25f;
        }
    }
    __syncthreads();
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < kRows; i += kThreads) s += accu[i];
    if (s == 0x1234567ull) sink[0] = s;
}

template <int VAR>
static double run(int sms, int iters) {
    unsigned long long *sink;
    cudaMalloc(&sink, 8);
    auto k = rmw_kernel<VAR>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kRows * 4);
    k<<<sms, kThreads, kRows * 4>>>(iters / 10, sink);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<sms, kThreads, kRows * 4>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(sink);
    return (double)sms * kThreads * (double)iters * kUnroll / (ms * 1e-3);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 20000;
    const char *names[5] = {"atomicAdd(float) CAS loop", "atomicAdd(uint32)", "atomicAdd(uint64)", "atomicOr(uint32)", "plain load-add-store"};
    double v[5] = {run<0>(sms, iters), run<1>(sms, iters), run<2>(sms, iters), run<3>(sms, iters), run<4>(sms, iters)};
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"threads_per_cta\": %d, \"rows\": %d", p.name, sms, kThreads, kRows);
    for (int i = 0; i < 5; ++i) printf(", \"%s\": {\"ops_per_s\": %.4g, \"ops_per_clk_per_sm_at_1.9GHz\": %.3f}", names[i], v[i], v[i] / sms / 1.9e9);
    printf("}\n");
    return 0;
}
