// Micro-benchmark: per-SM throughput of cp.async.bulk (global/L2 -> shared) vs. copies in flight and chunk size,
// all SMs streaming the same L2-resident 1.3 MB buffer (the weight stream of the forward-map kernel).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 1) k(const uint8_t* src, size_t src_bytes, int depth, uint32_t chunk, int iters, long long* out, int nprod, int lanes) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar[16 * 8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 128; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < nprod && lane < lanes) {
        const int pid = warp * lanes + lane;
        uint64_t* bar_ = bar + 16 * pid; uint8_t* sm_ = sm + (size_t)pid * depth * chunk;
        size_t off = ((size_t)blockIdx.x * 4096 + (size_t)pid * 65536) % src_bytes;
        const long long t0 = clock64();
        for (int i = 0; i < iters + depth; ++i) {
            const int s = i % depth;
            if (i >= depth) {                                       // wait for the copy that used this slot
                const uint32_t par = ((i / depth) - 1) & 1;
                uint32_t ok;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar_[s])), "r"(par) : "memory");
                } while (!ok);
            }
            if (i < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_[s])), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sm_ + (size_t)s * chunk)), "l"(src + off), "r"(chunk), "r"(smem_u32(&bar_[s])) : "memory");
                off += chunk;
                if (off + chunk > src_bytes) off = 0;
            }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && pid == 0) out[0] = t1 - t0;
    }
}
int main() {
    const size_t src_bytes = 1312 * 1024;
    uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    long long* out; cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (uint32_t chunk : {16384u, 4096u})
        for (int cfg = 0; cfg < 6; ++cfg) {
            const int nprod = (int[]){1, 2, 4, 1, 1, 2}[cfg], lanes = (int[]){1, 1, 1, 2, 4, 2}[cfg];
            const int depth = 2, iters = 2000, grid = 148;
            if ((size_t)depth * chunk * nprod * lanes > 196 * 1024) continue;
            k<<<grid, 256, 200 * 1024>>>(src, src_bytes, depth, chunk, iters, out, nprod, lanes);
            k<<<grid, 256, 200 * 1024>>>(src, src_bytes, depth, chunk, iters, out, nprod, lanes);
            long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("chunk %5u warps %d lanes %d depth %d: %6.1f cycles per round, %5.1f B/clk/SM  %s\n", chunk, nprod, lanes, depth,
                   (double)h / iters, (double)chunk * iters * nprod * lanes / h, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
