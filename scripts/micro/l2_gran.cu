// Does cudaLimitMaxL2FetchGranularity make a 32-byte miss fill the whole 128-byte line?
// pass A reads sector 0 of every line of a 48 MB buffer (fits L2), pass B then reads sector s of every line; if the
// line was filled as a whole, pass B runs at L2-hit speed.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void rd(const float* __restrict__ p, size_t n_lines, int sector, float* out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t nt = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    // 8 lanes read one 32-byte sector; a warp reads sector `sector` of 4 consecutive lines
    for (size_t i = t; i < n_lines * 8; i += nt) {
        size_t line = i >> 3; int w = i & 7;
        acc += __ldcg(p + line * 32 + sector * 8 + w);
    }
    if (acc == 123.456f) out[0] = acc;
}
int main() {
    size_t bytes = (size_t)48 << 20, flush_bytes = (size_t)1 << 30;
    float *p, *out, *flush; cudaMalloc(&p, bytes); cudaMemset(p, 0, bytes); cudaMalloc(&out, 4); cudaMalloc(&flush, flush_bytes);
    size_t n_lines = bytes / 128;
    for (int gran : {0, 32, 64, 128}) {
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        size_t lim; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
        for (int sector : {0, 1, 2, 3}) {
            cudaMemset(flush, 1, flush_bytes);                      // evict
            cudaEvent_t a, b, c; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c);
            cudaEventRecord(a);
            rd<<<148 * 8, 256>>>(p, n_lines, 0, out);
            cudaEventRecord(b);
            rd<<<148 * 8, 256>>>(p, n_lines, sector, out);
            cudaEventRecord(c); cudaEventSynchronize(c);
            float ma, mb; cudaEventElapsedTime(&ma, a, b); cudaEventElapsedTime(&mb, b, c);
            printf("granularity %3zu: pass A (sector 0, cold) %.1f us; pass B (sector %d) %.1f us\n", lim, ma * 1e3, sector, mb * 1e3);
        }
    }
    return 0;
}
