// Microbenchmark: DRAM throughput of sector-scattered reads.  A warp request = 32 lanes x 4 B (or 16 B), laid out as
// runs of `run` elements every `period` elements (like NCHW window rows).  8 independent requests in flight per warp.
#include <cuda_runtime.h>
#include <stdio.h>
template <typename T>
__global__ void k(const T* __restrict__ p, size_t n, int run, int period, float* out) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int runs_per_req = 32 / run;                        // run divides 32
    const size_t req_span = (size_t)runs_per_req * period;     // elements covered by one request
    const size_t lane_off = (size_t)(lane / run) * period + (lane % run);
    const size_t n_req = n / req_span;
    float acc = 0.f;
    for (size_t r = warp * 8; r + 8 <= n_req; r += nwarps * 8) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (r + u) * req_span + lane_off);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += *reinterpret_cast<float*>(&v[u]);
    }
    if (acc == 123.456f) out[0] = acc;
}
template <typename T> void run_case(const void* p, size_t bytes, int run, int period, float* out, const char* name) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    size_t n = bytes / sizeof(T);
    k<T><<<148 * 8, 256>>>((const T*)p, n, run, period, out);
    cudaEventRecord(a);
    k<T><<<148 * 8, 256>>>((const T*)p, n, run, period, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double useful = (double)bytes / period * run;
    printf("%s run %4zu B every %4zu B: useful %.0f GB/s, span %.0f GB/s (%.3f ms)\n", name, run * sizeof(T), period * sizeof(T),
           useful / ms / 1e6, bytes / ms / 1e6, ms);
}
int main() {
    size_t bytes = (size_t)8 << 30;
    void* p; cudaMalloc(&p, bytes); cudaMemset(p, 0, bytes);
    float* out; cudaMalloc(&out, 4);
    size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit: %zu\n", lim);
    for (int g = 0; g < 2; ++g) {
        if (g == 1) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32); cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("-- set granularity -> %zu\n", lim); }
        int c4[][2] = {{32, 32}, {8, 16}, {8, 32}, {8, 80}, {16, 32}, {16, 80}, {4, 8}, {4, 16}, {8, 20}, {8, 40}, {16, 40}};
        for (auto& c : c4) run_case<float>(p, bytes, c[0], c[1], out, "f32 ");
        int c16[][2] = {{32, 32}, {2, 4}, {2, 5}, {2, 10}, {2, 20}, {4, 5}, {4, 10}, {4, 20}, {4, 8}, {8, 16}, {1, 2}, {1, 4}};
        for (auto& c : c16) run_case<float4>(p, bytes, c[0], c[1], out, "f128");
    }
    return 0;
}
