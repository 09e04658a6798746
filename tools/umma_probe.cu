// Timing probe for small tcgen05.mma (kind::f16, M 128, K 16): how long a batch of NM MMAs takes
// from the first issue to the commit's mbarrier arrival, as a function of N and of the number of
// independent accumulators the batch rotates over.  One CTA, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_prof/umma_probe tools/umma_probe.cu && tools/_prof/umma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw64(uint32_t a)
{
    return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void probe(int N, int nacc, int NM, int reps, int nissue, long long *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bars[4];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t *)sm)[i] = 0x3c003c00u;   // halves = 1.0
    const uint32_t b = smem_u32(&bars[threadIdx.x >> 5]);
    if ((threadIdx.x & 31) == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < nissue) {
        const uint32_t wq = threadIdx.x >> 5;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t a = desc_sw64(smem_u32(sm)), w = desc_sw64(smem_u32(sm) + 32768);
        long long t_issue = 0, t_done = 0;
        uint32_t ph = 0;
        for (int r = 0; r < reps; r++) {
            const long long t0 = clock64();
            for (int i = 0; i < NM; i++) umma(tm + wq * 128u + (uint32_t)(nacc == 1 ? 0 : (i & (nacc - 1))) * (uint32_t)N, a, w, idesc, i >= nacc);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
            const long long t1 = clock64();
            uint32_t ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
            } while (!ok);
            ph ^= 1;
            const long long t2 = clock64();
            if (r > 0) { t_issue += t1 - t0; t_done += t2 - t0; }
        }
        out[2 * wq] = t_issue / (reps - 1);
        out[2 * wq + 1] = t_done / (reps - 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}
int main()
{
    long long *d, h[8];
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    const int Ns[] = {32, 64, 96, 128, 256};
    printf("M 128, K 16, kind::f16; cycles from first issue to [end of issue loop | commit arrival]\n");
    for (int NM : {1, 2, 3, 12, 24})
        for (int N : {32, 64})
            for (int nacc : {1, 2})
                for (int nissue : {1, 2, 4}) {
                    if (nacc > NM) continue;
                    probe<<<1, 128, 70000>>>(N, nacc, NM, 20, nissue, d);
                    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
                    printf("MMAs %2d  N %3d  accumulators %d  issuing warps %d: issue %5lld  done %5lld  (%4lld per MMA)  last warp done %5lld\n", NM, N, nacc, nissue, h[0], h[1], h[1] / NM, h[2 * nissue - 1]);
                }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
