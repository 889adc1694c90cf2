// FP64 issue rates on B200: DFMA (vector pipe), DMMA m8n8k4 (mma.sync f64), F2F.F64.F32, and FFMA for scale.
// Each warp runs a long chain of 8 independent accumulators; 148 x 4 CTAs x 256 threads.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
constexpr int ITERS = 4096;
__global__ void k_dfma(double* out, double a, double b) {
    double acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    double s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, float a, float b) {
    float acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], a, b);
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double* out, double a, double b) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_cvt(double* out, const float* in) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = in[threadIdx.x + i];
    double acc = 0;
    float f = v[0];
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { double d = (double)(v[i] + f); f = (float)__double2hiint(d); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + f;
}
template <class F> float run(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    const int grid = 148 * 4, thr = 256;
    double* dout; float* fout; float* fin;
    cudaMalloc(&dout, grid * thr * 8); cudaMalloc(&fout, grid * thr * 4); cudaMalloc(&fin, 4096);
    cudaMemset(fin, 0, 4096);
    const double n_thread_ops = (double)grid * thr * ITERS * 8;
    float ms = run([&] { k_dfma<<<grid, thr>>>(dout, 1.0000001, 1e-9); });
    printf("DFMA : %.3f ms  %.2f TFLOP/s  (%.1f lanes/clk/SM at 1.965 GHz)\n", ms, 2 * n_thread_ops / ms * 1e-9,
           n_thread_ops / (ms * 1e-3) / 148 / 1.965e9);
    ms = run([&] { k_ffma<<<grid, thr>>>(fout, 1.0000001f, 1e-9f); });
    printf("FFMA : %.3f ms  %.2f TFLOP/s  (%.1f lanes/clk/SM)\n", ms, 2 * n_thread_ops / ms * 1e-9,
           n_thread_ops / (ms * 1e-3) / 148 / 1.965e9);
    ms = run([&] { k_dmma<<<grid, thr>>>(dout, 1.0000001, 1e-9); });
    const double n_mma = (double)grid * (thr / 32) * ITERS * 8;
    printf("DMMA m8n8k4: %.3f ms  %.2f TFLOP/s  (%.2f clk/SM per mma)\n", ms, 2 * 256 * n_mma / ms * 1e-9,
           (ms * 1e-3) * 1.965e9 * 148 / n_mma);
    ms = run([&] { k_cvt<<<grid, thr>>>(dout, fin); });
    printf("F2F.F64.F32 chain: %.3f ms (%.1f cvt lanes/clk/SM, dependent)\n", ms, n_thread_ops / (ms * 1e-3) / 148 / 1.965e9);
    return 0;
}
