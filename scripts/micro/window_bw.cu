// Attainable throughput of the NCHW window gather pattern: a warp reads, with one LDG.128 per lane, the 16-byte chunks of
// a (rows x nxc chunks) window of a plane with W floats per row; CU consecutive planes (channels) are in flight.
// Windows move pseudo-randomly per group of 32 planes.  Buffer >> L2, so every line comes from DRAM once.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
template <int CU>
__global__ void k(const float4* __restrict__ p, size_t n_planes, int W4, int H, int rows, int nxc, float* out) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t plane4 = (size_t)W4 * H;
    const int nch = rows * nxc;
    float acc = 0.f;
    for (size_t g = warp; g * 32 + 32 <= n_planes; g += nwarps) {       // one "item": 32 planes, same window
        uint32_t h = (uint32_t)(g * 2654435761u);
        const int y0 = (h >> 8) % (H - rows + 1), x0 = (h >> 20) % (W4 - nxc + 1);
        const int q = lane < nch ? lane : 0;
        const int off = (y0 + q / nxc) * W4 + x0 + q % nxc;
        for (int c = 0; c < 32; c += CU) {
            float4 v[CU];
#pragma unroll
            for (int u = 0; u < CU; ++u) v[u] = __ldg(p + (g * 32 + c + u) * plane4 + off);
#pragma unroll
            for (int u = 0; u < CU; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
int main() {
    size_t bytes = (size_t)6 << 30;
    float4* p; cudaMalloc(&p, bytes); cudaMemset(p, 0, bytes);
    float* out; cudaMalloc(&out, 4);
    struct Cfg { int W, H, rows, nxc; const char* name; } cfgs[] = {
        {20, 20, 9, 3, "stride32 20x20 win 9x3ch"}, {40, 40, 8, 3, "stride16 40x40 win 8x3ch"}, {80, 80, 5, 2, "stride8 80x80 win 5x2ch"},
        {20, 20, 8, 4, "stride32 20x20 win 8x4ch"}, {40, 40, 8, 4, "stride16 40x40 win 8x4ch"}};
    for (auto& c : cfgs) {
        size_t plane_bytes = (size_t)c.W * c.H * 4, n_planes = bytes / plane_bytes;
        for (int ctas : {2, 4, 8}) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            k<8><<<148 * ctas, 256>>>(p, n_planes, c.W / 4, c.H, c.rows, c.nxc, out);
            cudaEventRecord(a);
            k<8><<<148 * ctas, 256>>>(p, n_planes, c.W / 4, c.H, c.rows, c.nxc, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            double win = (double)n_planes * c.rows * c.nxc * 16;
            printf("%-28s %2d warps/SM: window bytes %.0f GB/s, plane span %.0f GB/s (%.2f ms)\n", c.name, ctas * 8, win / ms / 1e6,
                   (double)bytes / ms / 1e6, ms);
        }
    }
    return 0;
}
