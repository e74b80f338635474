// Micro-benchmarks for the "columns as M" orientation of the forward-map kernel (round 2):
//   A. cycles per tcgen05.mma.cta_group::2 kind::f16 for M = 128 (64 rows per CTA) over N, with 1..4 independent accumulators
//      (is the ~60-cycle floor of a small MMA a dependent-accumulator latency or a per-instruction cost?)
//   B. where D lands in TMEM for M = 128 / cta_group::2 (lane / column of element (m, n)), which CTA supplies which rows of B,
//      and whether an N = 16 MMA at a column offset continues rows of a wider accumulator
//   C. activation epilogue with thread = (column, 16 consecutive features): tcgen05.ld x16 -> 2 x GELU -> bf16x2 -> two
//      16-byte swizzled st.shared per thread
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma3_bench umma3_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../beetroots_b200/csrc/common.cuh"
#include "../beetroots_b200/csrc/mlp_simt.cuh"
#include "../beetroots_b200/csrc/mlp_tc3.cuh"
#include "../beetroots_b200/csrc/mlp_tc4.cuh"

// ---------------------------------------------------------------- A: rate
template <int NACC, int PAT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
rate_k(long long* out, int iters, int M, int N) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < (131072 + 65536) / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    const uint32_t tb = tptr;
    if (threadIdx.x < 32 && cluster_ctarank() == 0) {
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 131072);
        const uint32_t cols = (uint32_t)(M == 128 ? N / 2 : N);      // TMEM columns per accumulator
        const uint32_t desc_lo0 = 1u << 16, desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t a_lo = desc_lo0 + ((a0 + (uint32_t)(i & 15) * 8192u) >> 4), b_lo = desc_lo0 + ((b0 + (uint32_t)(i & 3) * 16384u) >> 4);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t acc = PAT == 1 ? (uint32_t)(ks % NACC) : (uint32_t)(i & (NACC - 1));
                    umma2_f16_lo(tb + acc * cols, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, idesc, 1u);
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (leader) umma2_commit_mc(smem_u32(&bar), 1);
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512)); }
}
template <int NACC, int PAT>
static void run_rate(int M, int N) {
    const int nacc = NACC, pattern = PAT;
    long long* d; cudaMalloc(&d, 16);
    const int iters = 2000, smem = 131072 + 65536 + 2048;
    cudaFuncSetAttribute(rate_k<NACC, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_k<NACC, PAT><<<148, 256, smem>>>(d, iters, M, N);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double cyc = h[1] / (4.0 * iters);
    printf("A cg2 M=%3d N=%3d nacc=%d pattern=%d: issue %.1f complete %.1f cyc/MMA -> %.0f MAC/cyc/SM (%.0f %% of 4096)  (%s)\n", M, N, nacc, pattern,
           h[0] / (4.0 * iters), cyc, (double)M * N * 16 / 2 / cyc, 100.0 * M * N * 16 / 2 / cyc / 4096, cudaGetErrorString(e));
    cudaFree(d);
}

// ---------------------------------------------------------------- B: layout
// A: CTA c, local row r <-> m = 64 c + r;  A[m][0] = m + 1, A[m][1] = 1.   B: CTA c, local row j <-> tag = 100 c + j;
// B[.][0] = 1, B[.][1] = 256 (tag + 1)  ->  D = (m + 1) + 256 (tag + 1): every TMEM element names its (m, CTA of the B row, row j)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
layout_k(float* out, int N, int N2, int col2) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const uint32_t crank = cluster_ctarank();
    uint8_t* A = smem;             // [64][64] bf16 SW128
    uint8_t* B = smem + 8192;      // [128][64]
    uint8_t* B2 = smem + 8192 + 16384;
    for (int i = threadIdx.x; i < (8192 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.x < 64) {
        const int r = threadIdx.x;
        st_bf16(A, sw128_offset(r, 0), (float)(64 * crank + r + 1));
        st_bf16(A, sw128_offset(r, 1), 1.f);
    }
    {
        const int j = threadIdx.x;
        st_bf16(B, sw128_offset(j, 0), 1.f);
        st_bf16(B, sw128_offset(j, 1), 256.f * (float)(100 * crank + j + 1));
        st_bf16(B2, sw128_offset(j, 2), 0.5f);                        // continuation: adds 0.5 A[m][2]; A[m][2] = 0 -> set below
    }
    if (threadIdx.x < 64) st_bf16(A, sw128_offset(threadIdx.x, 2), 1.f);   // D2 = D + 0.5 where the continuation lands
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    const uint32_t tb = tptr;
    if (threadIdx.x == 0 && crank == 0) {
        const uint32_t id1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // B rows k = 0, 1 only (A[m][2] meets zeros in B)
        umma2_f16(tb, umma_desc_sw128(smem_u32(A)), umma_desc_sw128(smem_u32(B)), id1, 0u);
        if (N2 > 0) {
            const uint32_t id2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            umma2_f16(tb + (uint32_t)col2, umma_desc_sw128(smem_u32(A)), umma_desc_sw128(smem_u32(B2)), id2, 1u);
        }
        umma2_commit_mc(smem_u32(&bar), 3);
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int c0 = 0; c0 < 128; c0 += 32) {
            float v[32];
            tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            for (int j = 0; j < 32; ++j) out[((size_t)crank * 128 + warp * 32 + lane) * 128 + c0 + j] = v[j];
        }
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512)); }
}
static void run_layout(int N, int N2, int col2) {
    float* d; cudaMalloc(&d, 2 * 128 * 128 * 4);
    cudaMemset(d, 0, 2 * 128 * 128 * 4);
    const int smem = 8192 + 32768 + 2048;
    cudaFuncSetAttribute(layout_k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    layout_k<<<2, 128, smem>>>(d, N, N2, col2);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(2 * 128 * 128);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    printf("B layout M=128 cg2 N=%d (+ continuation N2=%d at column %d): %s\n", N, N2, col2, cudaGetErrorString(e));
    // hypothesis: CTA c, lane l, column q holds m = 64 c + (l % 64), n = (l / 64) (N/2) + q, B row n supplied by CTA n / (N/2), local row n % (N/2)
    int bad = 0, bad2 = 0;
    for (int c = 0; c < 2; ++c)
        for (int l = 0; l < 128; ++l)
            for (int q = 0; q < N / 2; ++q) {
                const int m = 64 * c + (l % 64), n = (l / 64) * (N / 2) + q, bc = n / (N / 2), bj = n % (N / 2);
                float expect = (float)(m + 1) + 256.f * (float)(100 * bc + bj + 1);
                if (N2 > 0 && q >= col2 && q < col2 + N2 / 2) {
                    expect += 0.5f;
                }
                const float got = h[((size_t)c * 128 + l) * 128 + q];
                if (got != expect) { if (bad < 12) printf("   mismatch cta %d lane %3d col %3d: got %.1f expect %.1f\n", c, l, q, got, expect); ++bad; }
            }
    printf("   hypothesis [lane = m %% 64 + 64 (n / (N/2)), col = n %% (N/2); B rows [0, N/2) from CTA 0]: %d mismatches\n", bad);
    (void)bad2;
    for (int c = 0; c < 2; ++c)
        for (int l : {0, 1, 31, 32, 63, 64, 65, 127}) {
            printf("   cta %d lane %3d:", c, l);
            for (int q : {0, 1, 7, 8, N / 2 - 1, N / 2, N - 1}) {
                const float v = h[((size_t)c * 128 + l) * 128 + q];
                const int iv = (int)v, m = (iv - 1) % 256, tag = iv / 256 - 1;
                printf("  c%-3d (m %3d, B cta %d row %3d)%s", q, m, tag / 100, tag % 100, (v != (float)iv) ? "+h" : "  ");
            }
            printf("\n");
        }
    cudaFree(d);
}

// ---------------------------------------------------------------- C: epilogue
template <int W>   // features per tcgen05.ld
__global__ void __launch_bounds__(512, 1) epi_k(long long* out, int reps, float* sink) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tptr;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, qd = warp & 3, sub = warp >> 2;
    const int row = (qd & 1) * 32 + lane;                            // column of the tile = row of the activation K-block
    const uint32_t act = smem_u32(smem);
    // zero my TMEM lanes so that the math sees finite values
    {
        for (int c = 0; c < 512; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tb + ((uint32_t)(qd * 32) << 16) + (uint32_t)c), "r"(0x3f000000u + (uint32_t)c) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    float acc = 0.f;
    for (int r = 0; r < reps; ++r) {
        // unit = W features starting at f0 (walks over K-blocks); thread stores W * 2 bytes in 16-byte pieces
        const int f0 = ((r * 4 + sub) * W) % 1280;
        const uint32_t taddr = tb + ((uint32_t)(qd * 32) << 16) + (uint32_t)(((r * 4 + sub) * W) % 448);
        float v[W];
        if (W == 8) { float t8[8]; tmem_ld8(taddr, t8); for (int j = 0; j < 8; ++j) v[j] = t8[j]; }
        else if (W == 16) { float t16[16]; tmem_ld16(taddr, t16); for (int j = 0; j < 16; ++j) v[j] = t16[j]; }
        else { float t32[32]; tmem_ld32(taddr, t32); for (int j = 0; j < W; ++j) v[j] = t32[j]; }
        uint32_t p[W / 2];
#pragma unroll
        for (int j = 0; j < W; j += 2) {
            const float a = gelu2_tc(v[j]), b = gelu2_tc(v[j + 1]);
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p[j / 2]) : "f"(b), "f"(a));
        }
#pragma unroll
        for (int j = 0; j < W / 8; ++j) {
            const int f = f0 + 8 * j, kb = f >> 6, ch = (f >> 3) & 7;
            const uint32_t dst = act + (uint32_t)kb * 8192u + (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(p[4 * j]), "r"(p[4 * j + 1]), "r"(p[4 * j + 2]), "r"(p[4 * j + 3]) : "memory");
        }
        acc += v[0];
    }
    fence_proxy_async();
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[threadIdx.x] = acc;
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512)); }
}
template <int W> static void run_epi() {
    long long* d; cudaMalloc(&d, 16);
    float* sink; cudaMalloc(&sink, 4096);
    const int reps = 400, smem = 168 * 1024 + 2048;
    cudaFuncSetAttribute(epi_k<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    epi_k<W><<<148, 512, smem>>>(d, reps, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("C epilogue, 16 warps, %2d features per tcgen05.ld: %.1f cycles per unit per warp-slot, %.2f elements / cycle / SM  (%s)\n", W, (double)h / reps,
           512.0 * W * reps / (double)h, cudaGetErrorString(e));
    cudaFree(d); cudaFree(sink);
}


// ---------------------------------------------------------------- D: cost of tcgen05.commit forms inside an MMA stream
// per iteration: 4 MMAs (M = 128, N = 144) + commit variant; barriers are initialised with a huge count so no phase completes
template <int VAR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
commit_k(long long* out, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < (131072 + 65536) / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar[i]), i == 7 ? 1 : (1 << 20)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    const uint32_t tb = tptr;
    if (threadIdx.x < 32 && cluster_ctarank() == 0) {
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(144 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 131072);
        const uint32_t desc_lo0 = 1u << 16, desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t a_lo = desc_lo0 + ((a0 + (uint32_t)(i & 15) * 8192u) >> 4), b_lo = desc_lo0 + ((b0 + (uint32_t)(i & 3) * 16384u) >> 4);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma2_f16_lo(tb, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, idesc, 1u);
                const uint32_t b = smem_u32(&bar[i & 3]);
                if (VAR == 1) umma2_commit_mc(b, 3);
                if (VAR == 2) umma2_commit_mc(b, 1);
                if (VAR == 3) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
                if (VAR == 4) { umma2_commit_mc(b, 3); umma2_commit_mc(smem_u32(&bar[4 + (i & 1)]), 3); }
                if (VAR == 5) umma2_commit_mc(b, 2);
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (leader) umma2_commit_mc(smem_u32(&bar[7]), 1);
        __syncwarp();
        mbar_wait(smem_u32(&bar[7]), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512)); }
}
template <int VAR> static void run_commit(const char* what) {
    long long* d; cudaMalloc(&d, 16);
    const int iters = 1000, smem = 131072 + 65536 + 2048;
    cudaFuncSetAttribute(commit_k<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    commit_k<VAR><<<148, 256, smem>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("D 4 MMAs (144 cycles) + %-44s: issue %.0f, complete %.0f cycles per iteration  (%s)\n", what, (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(e));
    cudaFree(d);
}


// ---------------------------------------------------------------- E: the weight-ring handshake in isolation
// issuer (warp 0): spin full[s] -> 4 MMAs -> commit empty[s] (multicast);  producers (warps 1..NP, lane 0, leader CTA only):
// spin empty[s] -> arrive full[s];  NW extra warps park in try_wait on a barrier that never completes (like idle epilogue warps)
template <int MMAS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(768, 1)
ring_k(long long* out, int iters, int n_slots, int n_prod, int n_wait, int spin_try, int arr) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[24];
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < (131072 + 65536) / 16; i += 768) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x == 0) { for (int i = 0; i < 24; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    const uint32_t tb = tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full = smem_u32(&bar[0]), empty = smem_u32(&bar[8]), never = smem_u32(&bar[16]), fin = smem_u32(&bar[17]);
    if (cluster_ctarank() == 0) {
        if (warp == 0) {
            const bool leader = elect_one();
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(144 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 131072);
            const uint32_t desc_lo0 = 1u << 16, desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            uint32_t slot = 0, use = 0;
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                mbar_spin(full + 8 * slot, use & 1u);
                tc_fence_after();
                const uint32_t a_lo = desc_lo0 + ((a0 + (uint32_t)(i & 15) * 8192u) >> 4), b_lo = desc_lo0 + ((b0 + (slot & 3u) * 16384u) >> 4);
                if (leader) {
                    if (MMAS) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) umma2_f16_lo(tb, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, idesc, 1u);
                    }
                    if (arr == 0) umma2_commit_mc(empty + 8 * slot, 3);
                    else if (arr == 1) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty + 8 * slot) : "memory");
                    else if (arr == 3) umma2_commit_mc(empty + 8 * slot, 1);
                    else mbar_arrive(empty + 8 * slot);
                }
                __syncwarp();
                if (++slot == (uint32_t)n_slots) { slot = 0; ++use; }
            }
            const long long t1 = clock64();
            if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
            if (leader) umma2_commit_mc(fin, 1);
            __syncwarp();
            mbar_wait(fin, 0);
        } else if (warp <= n_prod) {
            if (lane == 0) {
                uint32_t slot = 0, use = 0;
                for (int i = 0; i < iters; ++i) {
                    if (i % n_prod == warp - 1) {
                        if (use > 0) { if (spin_try) mbar_wait(empty + 8 * slot, (use - 1) & 1u); else mbar_spin(empty + 8 * slot, (use - 1) & 1u); }
                        mbar_arrive(full + 8 * slot);
                    }
                    if (++slot == (uint32_t)n_slots) { slot = 0; ++use; }
                }
            }
        }
    }
    if (warp >= 8 && warp < 8 + n_wait) {
        // parked waiters: released at the end by thread 0 of their CTA
        mbar_wait(never, 0);
    }
    if (warp < 8) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 0) mbar_arrive(never);
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512)); }
}
template <int MMAS> static void run_ring(int n_slots, int n_prod, int n_wait, int spin_try, int arr = 0) {
    long long* d; cudaMalloc(&d, 16);
    const int iters = 2000, smem = 131072 + 65536 + 2048;
    cudaFuncSetAttribute(ring_k<MMAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    ring_k<MMAS><<<148, 768, smem>>>(d, iters, n_slots, n_prod, n_wait, spin_try, arr);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const char* an[4] = {"commit multicast both", "commit local", "plain arrive", "commit multicast leader"};
    printf("E ring: %d slots, %d producers (%s), %2d parked warps, %-24s %s: %.0f cycles per step  (%s)\n", n_slots, n_prod, spin_try ? "try_wait" : "test_wait spin", n_wait, an[arr],
           MMAS ? "4 MMAs (144 cyc) per step" : "no MMAs", (double)h / iters, cudaGetErrorString(e));
    cudaFree(d);
}


// ---------------------------------------------------------------- F: the tc4 issuer loop in isolation
// issuer (warp IW of the leader): ballot-poll of all ring slots, then 4 MMAs (N) + commit per ready step, exactly like
// mlp_tc4.cuh; producers (warps 1..NP, lane 0): spin empty[s] -> arrive full[s].  Variants (var bits): 1 = commit only after
// every second step, 2 = no __syncwarp per step, 4 = producers use try_wait, 8 = single elected lane runs the whole loop
// with per-step test_wait
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(832, 1)
iss_k(long long* out, int iters, int n_slots, int n_prod, int N, int var, int iw) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[24];
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < (131072 + 65536) / 16; i += 832) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x == 0) { for (int i = 0; i < 24; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    const uint32_t tb = tptr;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t full = smem_u32(&bar[0]), empty = smem_u32(&bar[8]), fin = smem_u32(&bar[17]);
    if (cluster_ctarank() == 0) {
        if (warp == iw) {
            const bool leader = elect_one();
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 131072);
            const uint32_t desc_lo0 = 1u << 16, desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t act_lo = desc_lo0 + (a0 >> 4), ring_lo = desc_lo0 + (b0 >> 4), slot_lo = 8192u >> 4;
            uint32_t slot = 0, use = 0, ready = 0;
            const long long t0 = clock64();
            if (var & (8 | 128)) {
                if (leader) {
                    for (int i = 0; i < iters; ++i) {
                        if (!(var & 128)) mbar_spin(full + 8 * slot, use & 1u);
                        tc_fence_after();
                        const uint32_t a_lo = act_lo + (uint32_t)(i & 15) * 512u, b_lo = ring_lo + slot * slot_lo;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) umma2_f16_lo(tb, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, idesc, 1u);
                        umma2_commit_mc(empty + 8 * slot, 3);
                        if (++slot == (uint32_t)n_slots) { slot = 0; ++use; }
                    }
                }
                __syncwarp();
            } else {
                for (int i = 0; i < iters; ++i) {
                    if (ready == 0 && !(var & 48)) {
                        uint32_t s2 = slot + (uint32_t)lane, u2 = use;
                        if (s2 >= (uint32_t)n_slots) { s2 -= (uint32_t)n_slots; ++u2; }
                        uint32_t mask;
                        do {
                            const bool ok = lane < n_slots && mbar_test(full + 8 * s2, u2 & 1u);
                            mask = __ballot_sync(0xffffffffu, ok);
                        } while (!(mask & 1u));
                        ready = (uint32_t)__ffs((int)~mask) - 1u;
                        tc_fence_after();
                    }
                    --ready;
                    const uint32_t a_lo = act_lo + (uint32_t)(i & 15) * 512u, b_lo = ring_lo + slot * slot_lo;
                    if (var & 64) {
                        if (leader) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma2_f16_lo(tb, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, idesc, 1u);
                            umma2_commit_mc(empty + 8 * slot, 3);
                        }
                    } else {
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) umma2_f16_lo(tb, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, idesc, 1u);
                    }
                    if (leader && (!(var & 1) || (slot & 1u) || slot == (uint32_t)n_slots - 1)) {
                        umma2_commit_mc(empty + 8 * slot, 3);
                        if ((var & 1) && (slot & 1u)) umma2_commit_mc(empty + 8 * (slot - 1), 3);
                    }
                    }
                    if (!(var & 2)) __syncwarp();
                    if (++slot == (uint32_t)n_slots) { slot = 0; ++use; }
                }
            }
            const long long t1 = clock64();
            if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
            if (leader) umma2_commit_mc(fin, 1);
            __syncwarp();
            mbar_wait(fin, 0);
        } else if (warp >= 1 && warp <= n_prod && !(var & 16)) {
            if (lane == 0) {
                uint32_t slot = 0, use = 0;
                for (int i = 0; i < iters; ++i) {
                    if (i % n_prod == warp - 1) {
                        if (use > 0) { if (var & 4) mbar_wait(empty + 8 * slot, (use - 1) & 1u); else mbar_spin(empty + 8 * slot, (use - 1) & 1u); }
                        mbar_arrive(full + 8 * slot);
                    }
                    if (++slot == (uint32_t)n_slots) { slot = 0; ++use; }
                }
            }
        }
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512)); }
}
static void run_iss(int n_slots, int n_prod, int N, int var, int iw) {
    long long* d; cudaMalloc(&d, 16);
    const int iters = 2000, smem = 131072 + 65536 + 2048;
    cudaFuncSetAttribute(iss_k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    iss_k<<<148, 832, smem>>>(d, iters, n_slots, n_prod, N, var, iw);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("F issuer loop: %d slots, %d producers, N = %3d (%3d MMA cycles per step), variant %2d, issuer warp %2d: %.0f cycles per step  (%s)\n", n_slots, n_prod, N, N,
           var, iw, (double)h / iters, cudaGetErrorString(e));
    cudaFree(d);
}

int main(int argc, char** argv) {
    const int what = argc > 1 ? atoi(argv[1]) : 7;
    setvbuf(stdout, nullptr, _IONBF, 0);
    if (what & 2) {
        run_layout(32, 0, 0);
        run_layout(64, 16, 24);      // continuation of "rows" [24, 32) and [56, 64) of a 64-wide accumulator by an N = 16 MMA at column 24
        run_layout(160, 16, 72);
    }
    if (what & 1) {
        for (int N : {16, 32, 64, 96, 128, 144, 160, 192, 224, 256}) run_rate<1, 1>(128, N);
        for (int N : {16, 32, 64, 96, 128, 144, 160, 192}) run_rate<2, 1>(128, N);
        for (int N : {16, 64, 96, 128, 144}) run_rate<2, 4>(128, N);
        for (int N : {16, 64, 96, 128, 144}) run_rate<4, 1>(128, N);
        for (int N : {64, 128}) run_rate<2, 1>(256, N);
        run_rate<1, 1>(256, 128);
    }
    if (what & 4) { run_epi<8>(); run_epi<16>(); run_epi<32>(); }
    if (what & 16) {
        for (int arr = 0; arr < 4; ++arr) { run_ring<0>(1, 1, 0, 0, arr); run_ring<0>(5, 4, 0, 0, arr); }
        run_ring<0>(1, 1, 0, 1, 0); run_ring<0>(1, 1, 0, 1, 2); run_ring<0>(8, 4, 0, 0, 0); run_ring<0>(8, 4, 16, 0, 0);
        run_ring<1>(1, 1, 0, 0, 0); run_ring<1>(5, 4, 0, 0, 0); run_ring<1>(8, 4, 0, 0, 0); run_ring<1>(8, 4, 16, 0, 0); run_ring<1>(8, 4, 16, 1, 0); run_ring<1>(8, 4, 0, 0, 1);
    }
    if (what & 32) {
        for (int N : {144, 16}) { for (int var : {16, 16 + 64, 16 + 2, 16 + 64 + 2, 16 + 128}) run_iss(7, 4, N, var, 24); run_iss(7, 4, N, 16, 0); run_iss(7, 4, N, 16 + 64, 0); }
    }
    if (what & 8) {
        run_commit<0>("no commit");
        run_commit<1>("commit multicast to both CTAs");
        run_commit<2>("commit multicast, leader only");
        run_commit<3>("commit without multicast (local barrier)");
        run_commit<4>("two multicast commits to both CTAs");
        run_commit<5>("commit multicast, peer only");
    }
    return 0;
}
