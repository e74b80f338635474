// Micro-benchmark: issue cost / throughput of descriptor-based TMA (cp.async.bulk.tensor.2d) vs the raw 1-D bulk copy,
// one issuing thread per CTA, L2-resident source.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bench tma_bench.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, const uint8_t* src, size_t src_bytes, int depth, int rows, int iters,
                                            int use_tensor, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar[16];
    const uint32_t chunk = (uint32_t)rows * 1024u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int total_rows = (int)(src_bytes / 1024);
        int row = (blockIdx.x * 4) % (total_rows - rows);
        const long long t0 = clock64();
        for (int i = 0; i < iters + depth; ++i) {
            const int s = i % depth;
            if (i >= depth) {
                const uint32_t par = ((i / depth) - 1) & 1;
                uint32_t ok;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(par) : "memory");
                } while (!ok);
            }
            if (i < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(chunk) : "memory");
                if (use_tensor)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(smem_u32(sm + (size_t)s * chunk)), "l"(&tm), "r"(0), "r"(row), "r"(smem_u32(&bar[s])) : "memory");
                else
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(sm + (size_t)s * chunk)), "l"(src + (size_t)row * 1024), "r"(chunk), "r"(smem_u32(&bar[s])) : "memory");
                row += rows;
                if (row + rows > total_rows) row = 0;
            }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
        // phase 2: 'depth' copies issued back to back, no waits in between; one expect_tx per copy or one for all
        for (int mode = 0; mode < 2; ++mode) {
            const long long u0 = clock64();
            if (mode == 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[15])), "r"(chunk * depth) : "memory");
            for (int s = 0; s < depth; ++s) {
                const uint32_t b = mode == 1 ? smem_u32(&bar[15]) : smem_u32(&bar[8 + (s & 3)]);
                if (mode == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sm + (size_t)s * chunk)), "l"(src + (size_t)s * chunk), "r"(chunk), "r"(b) : "memory");
            }
            const long long u1 = clock64();
            if (blockIdx.x == 0) out[1 + mode] = u1 - u0;
            // drain
            for (volatile int w = 0; w < 3000; ++w) { }
        }
    }
}
int main() {
    const size_t src_bytes = 1312 * 1024;
    uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    long long* out; cudaMalloc(&out, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cuInit(0);
    for (int rows : {32, 16, 4}) {
        CUtensorMap tm;
        cuuint64_t gdim[2] = {256, src_bytes / 1024};          // 256 x uint32 = 1024 B per row
        cuuint64_t gstride[1] = {1024};
        cuuint32_t box[2] = {256, (cuuint32_t)rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, src, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); return 1; }
        for (int use_tensor : {0, 1})
            for (int depth : {1, 2, 4}) {
                if ((size_t)depth * rows * 1024 > 196 * 1024) continue;
                const int iters = 2000;
                k<<<148, 128, 200 * 1024>>>(tm, src, src_bytes, depth, rows, iters, use_tensor, out);
                k<<<148, 128, 200 * 1024>>>(tm, src, src_bytes, depth, rows, iters, use_tensor, out);
                long long hh[3]; cudaMemcpy(hh, out, 24, cudaMemcpyDeviceToHost); long long h = hh[0];
                printf("%s chunk %2d KB depth %d: %7.1f cycles per copy, %5.1f B/clk/SM  %s\n", use_tensor ? "tensor-map TMA" : "raw bulk copy ", rows, depth,
                       (double)h / iters, (double)rows * 1024 * iters / h, cudaGetErrorString(cudaGetLastError()));
                printf("      back-to-back issue of %d copies: %.1f cycles per copy (expect_tx per copy), %.1f (one expect_tx for all)\n", depth, (double)hh[1] / depth, (double)hh[2] / depth);
            }
    }
    return 0;
}
