// Micro-benchmark: cycles per tcgen05.mma.cta_group::2 (kind::f16, SS operands, K-major SW128), M = 256, several N.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../beetroots_b200/csrc/common.cuh"
#include "../beetroots_b200/csrc/mlp_simt.cuh"
#include "../beetroots_b200/csrc/mlp_tc2.cuh"

template <int N, int M>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) k(long long* out, int iters, int bwalk, int traffic, const uint8_t* gsrc, int nissue) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < (16384 * 4 + 131072) / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    const uint32_t tb = tptr;
    __shared__ uint64_t bar_b;
    if (threadIdx.x == 1) { mbar_init(smem_u32(&bar_b), 1); }
    __syncthreads();
    if ((threadIdx.x == 0 || (threadIdx.x == 32 && nissue >= 2)) && cluster_ctarank() == 0) {
        const uint32_t tb = tptr + ((threadIdx.x == 32 && nissue == 2) ? 128u : 0u);
        const uint32_t mybar = threadIdx.x == 0 ? smem_u32(&bar) : smem_u32(&bar_b);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t st = i & 3;
            for (int ks = 0; ks < 4; ++ks)
                umma2_f16(tb, umma_desc_sw128(a0 + st * 16384 + ks * 32), umma_desc_sw128(b0 + (bwalk ? (i & 15) * 8192 : 0) + ks * 32), idesc, 1u);
        }
        long long t1 = clock64();
        umma2_commit_mc(mybar, 1);
        mbar_wait(mybar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    if (traffic && threadIdx.x >= 128) {
        volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(&tptr) + 1;
        uint32_t* p = reinterpret_cast<uint32_t*>(smem + 65536 + 131072);
        uint32_t n = 0;
        for (int r = 0; r < iters * 4; ++r) { p[(threadIdx.x - 128) + 128 * (r & 15)] = n++; }
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512)); }
}

template <int N, int M> void run(int grid, int bwalk = 0, int traffic = 0, int nissue = 1) {
    long long* d; cudaMalloc(&d, 16);
    const int iters = 2000, smem = 16384 * 4 + 131072 + 8192 + 2048;
    cudaFuncSetAttribute(k<N, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<N, M><<<grid, 256, smem>>>(d, iters, bwalk, traffic, nullptr, nissue);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("cg2 M=%d N=%3d grid=%3d bwalk=%d traffic=%d issuers=%d (per-issuer MMAs = iters*4): issue %.1f cyc/MMA, complete %.1f cyc/MMA  -> %.0f MAC/cyc/SM  (%s)\n", M, N, grid, bwalk, traffic, nissue,
           h[0] / (4.0 * iters), h[1] / (4.0 * iters), (double)M * N * 16 / 2 / (h[1] / (4.0 * iters)), cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    run<128, 256>(148, 1, 0, 1); run<128, 256>(148, 1, 0, 2); run<128, 256>(148, 1, 0, 3); run<128, 256>(148, 1, 1, 2);
    return 0;
}
