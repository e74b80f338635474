// Micro-benchmark: cycles per tcgen05.mma (kind::f16, SS operands, K-major SW128) for several N.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../beetroots_b200/csrc/mlp_tc.cuh"

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int nslots) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2[4];
    __shared__ uint64_t bar3;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < (16384 * 4 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar3), 1); mbar_arrive(smem_u32(&bar3)); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar2[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = tptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t st = i & 3;
            if (MODE & 2) { mbar_wait(smem_u32(&bar3), 0); if (!(MODE & 4)) tc_fence_after(); }     // wait on an already-complete barrier
            for (int ks = 0; ks < 4; ++ks)
                umma_f16(tb + (i % nslots) * N, umma_desc_sw128(a0 + st * 16384 + ks * 32), umma_desc_sw128(b0 + ks * 32), idesc, 1u);
            if (MODE & 1) umma_commit(smem_u32(&bar2[st]));                         // commit per K-block (nobody waits)
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512)); }
}

template <int N, int MODE> void run(int grid, int nslots) {
    long long* d; cudaMalloc(&d, 16);
    const int iters = 2000, smem = 16384 * 4 + 32768 + 2048;
    cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<N, MODE><<<grid, 128, smem>>>(d, iters, nslots);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("mode=%d N=%3d grid=%3d slots=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  -> %.0f MAC/cyc/SM  (%s)\n", MODE, N, grid, nslots,
           h[0] / (4.0 * iters), h[1] / (4.0 * iters), 128.0 * N * 16 / (h[1] / (4.0 * iters)), cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    run<64,2>(148, 1); run<64,6>(148, 1); run<64,7>(148, 1);
    return 0;
}
