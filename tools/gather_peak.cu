// Microbenchmark: ceiling of dependent random 4-byte gathers from an L2-resident table on this
// GPU, at the step kernel's launch shape (64-thread CTAs, 64 registers) -- the practical
// roofline of the sphere-tracing march, whose every sample is one such gather.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_prof/gather_peak tools/gather_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(64) chase(const float *__restrict__ tab, unsigned mask, int iters, unsigned *out, int spread)
{
    unsigned idx[ILP];
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int j = 0; j < ILP; j++) idx[j] = (gid * 2654435761u + j * 40503u) & mask;
    float acc = 0.f;
    for (int i = 0; i < iters; i++) {
        float v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) v[j] = __ldg(tab + idx[j]);
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            acc += v[j];
            // next index depends on the loaded value; spread = 0: anywhere in the table,
            // spread > 0: within +-spread cells (neighbouring lanes stay in neighbouring lines)
            const unsigned h = (idx[j] * 1664525u + __float_as_uint(v[j]) + 1013904223u);
            idx[j] = spread ? ((idx[j] + (h >> 8) % (2 * spread + 1) - spread) & mask) : ((h >> 4) & mask);
        }
    }
    if (acc == 123.456f) out[0] = 1;
}

int main()
{
    const unsigned n = 1u << 20;  // 4 MB of float32: the indoor map's EDT size
    float *tab; unsigned *out;
    cudaMalloc(&tab, n * 4); cudaMalloc(&out, 4);
    float *h = (float *)malloc(n * 4);
    for (unsigned i = 0; i < n; i++) h[i] = (float)(rand() & 1023);
    cudaMemcpy(tab, h, n * 4, cudaMemcpyHostToDevice);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, clk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    for (int spread = 0; spread <= 64; spread = spread ? spread * 8 : 8) {
        for (int ctas_per_sm = 4; ctas_per_sm <= 32; ctas_per_sm *= 2) {
            for (int ilp = 1; ilp <= 4; ilp *= 2) {
                const int grid = p.multiProcessorCount * ctas_per_sm;
                float ms = 0;
                for (int rep = 0; rep < 2; rep++) {
                    cudaEventRecord(e0);
                    if (ilp == 1) chase<1><<<grid, 64>>>(tab, n - 1, iters, out, spread);
                    if (ilp == 2) chase<2><<<grid, 64>>>(tab, n - 1, iters, out, spread);
                    if (ilp == 4) chase<4><<<grid, 64>>>(tab, n - 1, iters, out, spread);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    cudaEventElapsedTime(&ms, e0, e1);
                }
                const double g = (double)grid * 64 * ilp * iters;
                printf("spread %3d  CTAs/SM %2d (warps/SM %2d)  ILP %d : %7.1f G gathers/s  %6.1f cycles/iteration  %.2f warp-gathers/clk/SM\n",
                       spread, ctas_per_sm, ctas_per_sm * 2, ilp, g / ms / 1e6, ms * 1e-3 * clk * 1e3 / iters,
                       g / 32 / (ms * 1e-3 * clk * 1e3) / p.multiProcessorCount);
            }
        }
    }
    return 0;
}
