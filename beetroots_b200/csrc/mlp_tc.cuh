// Forward map kernel, tcgen05 / TMEM variant: bf16 operands, fp32 accumulation in tensor memory.
//
// Same function as mlp_simt_kernel (NeuralNetworkApprox.compute_all / evaluate_log,
// modelling/forward_maps/neural_network_approx.py:107-122,152-305 + nnbma forward / vmap(jacfwd)).
//
// Persistent, warp-specialised, one CTA per SM.  A tile is 64 "columns" (a column = a row of
// theta, or one of its S-1 forward-mode tangents).  GEMM orientation per dense block l:
//     D[out feature (M=128), column (N=64)] = W_l[out, k] (A, K-major) x H[column, k] (B, K-major)
//   * B = the dense-concat activations of the 64 columns, resident in shared memory for the whole
//     tile: K-blocks of 64 features, each a [64 x 64] bf16 tile in the canonical K-major
//     SWIZZLE_128B layout (8 KB) -> 21 blocks = 168 KB for 1296 features; never touches HBM;
//   * A = weight tiles [128 x 64] bf16, pre-swizzled on the host into the exact shared-memory
//     image and streamed from L2 by the bulk-copy engine (cp.async.bulk, 16 KB per tile) through
//     a 3-stage mbarrier ring;
//   * D = 4 accumulator slots of 64 TMEM columns; the epilogue warps drain slot s (tcgen05.ld ->
//     bias + GELU(erf) / GELU' x tangent -> bf16 -> swizzled st.shared into B) while the MMA warp
//     fills slot s+1.  Because the network is a dense concat, block l+1 starts on the K-blocks
//     holding features of blocks < l before the epilogue of block l has finished.
// Warp roles: warp 0 = weight producer, warp 1 = TMEM allocator + MMA issuer (one elected thread),
// warps 4..11 = epilogue (TMEM lane quarter = warp % 4, column half = (warp - 4) / 4).
#pragma once
#include <cuda_bf16.h>
#include <vector>
#include "common.cuh"
#include "mlp_simt.cuh"

#define TC_BN 64            // columns per tile (UMMA N)
#define TC_BM 128           // output features per weight tile (UMMA M)
#define TC_BK 64            // features per K-block (one 128-byte swizzle atom of bf16)
#define TC_STAGES 3
#define TC_SLOTS 8            // 4 units in flight x 2 split-K halves (one per MMA-issuing thread)
#define TC_THREADS 384
#define TC_WTILE_BYTES (TC_BM * TC_BK * 2)
#define TC_KB_BYTES (TC_BN * TC_BK * 2)
#define TC_MAX_LAYERS (BT_MAX_BLOCKS + 1)

struct TcLayer {
    int in, nnew, n_mt, n_kb, first_new_kb, bias_off, is_out;
    int w_off;      // offset (in 1 KB units) of this layer's first weight tile in the stream
};
// bytes of weight tile (layer L, M-tile mt): only the 8-row groups holding live output features are streamed;
// the rest of the 128-row stage keeps (finite) data of earlier tiles and feeds accumulator lanes nobody reads.
__host__ __device__ __forceinline__ int tc_tile_kb(const TcLayer& L, int mt) {
    int rows = L.nnew - mt * TC_BM;
    rows = rows > TC_BM ? TC_BM : rows;
    return (rows + 7) / 8;                      // 1 KB per 8-row group (8 rows x 64 k x 2 B)
}
struct TcSched {
    int n_layers, n_wtiles, act_kb, n_out;
    TcLayer L[TC_MAX_LAYERS];
    const uint16_t* wstream;
    const float* bias;
    long long* dbg;       // optional timeline buffer (BT_TC_DEBUG=1): [0..63] MMA thread, [64..127] epilogue warp 4
};
struct TcNet {
    bool ready = false;
    TcSched sched{};
    void* d_w = nullptr;
    void* d_b = nullptr;
    size_t smem_bytes = 0;
};

static inline void tc_free(TcNet& t) {
    if (t.d_w) cudaFree(t.d_w);
    if (t.d_b) cudaFree(t.d_b);
    t = TcNet();
}

static inline uint16_t f32_to_bf16_rne(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40);
    const uint32_t lsb = (x >> 16) & 1u;
    x += 0x7fffu + lsb;
    return (uint16_t)(x >> 16);
}

// byte offset of element (row r, k in [0,64)) inside a K-major SWIZZLE_128B tile (rows of 128 bytes,
// 8-row groups of 1024 bytes, 16-byte chunk index XOR (r % 8))
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t k) {
    return (r >> 3) * 1024u + (r & 7u) * 128u + ((((k >> 3) ^ (r & 7u)) & 7u) << 4) + ((k & 7u) << 1);
}

// Build the bf16 weight stream + schedule.  Returns 0; t.ready tells whether the shapes fit.
static inline int tc_build(TcNet& t, int n_feat0, int n_blocks, const int32_t* block_new, const double* const* weights,
                           const double* const* biases, const double* w_out, int n_out, double out_scale) {
    t.ready = false;
    if (n_blocks + 1 > TC_MAX_LAYERS || n_out > 32) return 0;
    TcSched& s = t.sched;
    memset(&s, 0, sizeof s);
    s.n_layers = n_blocks + 1;
    s.n_out = n_out;
    int w = n_feat0, prev_in = 0, bias_off = 0, n_wtiles = 0, stream_kb = 0;
    for (int l = 0; l <= n_blocks; ++l) {
        TcLayer& L = s.L[l];
        L.in = w;
        L.nnew = l < n_blocks ? block_new[l] : n_out;
        L.is_out = l == n_blocks;
        L.n_mt = (L.nnew + TC_BM - 1) / TC_BM;
        L.n_kb = (L.in + TC_BK - 1) / TC_BK;
        L.first_new_kb = l == 0 ? 0 : prev_in / TC_BK;
        L.bias_off = bias_off;
        bias_off += L.n_mt * TC_BM;
        n_wtiles += L.n_mt * L.n_kb;
        L.w_off = stream_kb;
        for (int mt = 0; mt < L.n_mt; ++mt) stream_kb += tc_tile_kb(L, mt) * L.n_kb;
        prev_in = w;
        if (l < n_blocks) w += block_new[l];
    }
    s.n_wtiles = n_wtiles;
    s.act_kb = s.L[n_blocks].n_kb;
    const size_t smem = (size_t)s.act_kb * TC_KB_BYTES + (size_t)TC_STAGES * TC_WTILE_BYTES + 1024 + 4096 + 1024;
    if (n_feat0 > 64) return 0;
    if (smem > 227 * 1024) return 0;
    t.smem_bytes = smem;
    std::vector<uint16_t> ws((size_t)stream_kb * 512, 0);
    std::vector<float> bs(bias_off, 0.f);
    size_t pos_kb = 0;
    for (int l = 0; l <= n_blocks; ++l) {
        const TcLayer& L = s.L[l];
        const double* W = L.is_out ? w_out : weights[l];
        const double sc = L.is_out ? out_scale : 1.0;
        for (int o = 0; o < L.nnew && !L.is_out; ++o) bs[L.bias_off + o] = (float)biases[l][o];
        for (int mt = 0; mt < L.n_mt; ++mt)
            for (int kb = 0; kb < L.n_kb; ++kb, pos_kb += tc_tile_kb(L, mt)) {
                uint16_t* dst = ws.data() + pos_kb * 512;
                for (int r = 0; r < TC_BM; ++r) {
                    const int o = mt * TC_BM + r;
                    if (o >= L.nnew) continue;
                    for (int kk = 0; kk < TC_BK; ++kk) {
                        const int k = kb * TC_BK + kk;
                        if (k >= L.in) continue;
                        dst[sw128_offset(r, kk) / 2] = f32_to_bf16_rne((float)(sc * W[(size_t)o * L.in + k]));
                    }
                }
            }
    }
    if (cudaMalloc(&t.d_w, ws.size() * 2) != cudaSuccess) return -1;
    if (cudaMalloc(&t.d_b, bs.size() * 4 + 16) != cudaSuccess) return -1;
    if (cudaMemcpy(t.d_w, ws.data(), ws.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (cudaMemcpy(t.d_b, bs.data(), bs.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    s.wstream = (const uint16_t*)t.d_w;
    s.bias = (const float*)t.d_b;
    t.ready = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// busy-poll variant for the single-thread pipeline roles (producer, MMA issuers): test_wait returns immediately,
// try_wait parks the thread in hardware and pays a wake-up latency on every K-block
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// multicast variants (thread-block cluster): the copy lands at the same shared-memory offset of every CTA in
// ctaMask and performs complete_tx on the mbarrier at the same offset in each of them
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] x B[smem desc], kind::f16 (bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);     // start address        bits [0,14)
    d |= (uint64_t)1 << 16;                       // leading byte offset  bits [16,30)  (unused for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;            // stride byte offset   bits [32,46)
    d |= (uint64_t)1 << 46;                       // descriptor version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                       // layout type SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// GELU(erf) through one MUFU: x Phi(x) ~= 0.5 x (1 + tanh(x (p0 + p1 x^2 + p2 x^4))), coefficients fitted (minimax
// on [-8, 8]) to the erf form: max abs error 2.5e-5, far below the bf16 rounding of the stored activation
// (the textbook two-term tanh form is 4.7e-4 off).  The tangent uses the exact derivative of this function.
#define TC_G0 0.797507884f
#define TC_G1 0.0370056460f
#define TC_G2 -3.51516792e-4f
__device__ __forceinline__ float tanh_mufu(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gelu_tc(float x) {
    const float xc = fminf(fmaxf(x, -8.f), 8.f), x2 = xc * xc;
    const float t = tanh_mufu(xc * fmaf(x2, fmaf(x2, TC_G2, TC_G1), TC_G0));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}
__device__ __forceinline__ float gelu_prime_tc(float x) {
    const float xc = fminf(fmaxf(x, -8.f), 8.f), x2 = xc * xc;
    const float t = tanh_mufu(xc * fmaf(x2, fmaf(x2, TC_G2, TC_G1), TC_G0));
    const float du = fmaf(x2, fmaf(x2, 5.f * TC_G2, 3.f * TC_G1), TC_G0);
    return 0.5f * (1.f + t) + 0.5f * xc * (1.f - t * t) * du;
}
// 32-bit shared-window address + st.shared (STS) instead of a generic 64-bit store: ~4 fewer instructions per element
__device__ __forceinline__ void sts_bf16(uint32_t saddr, float x) {
    const __nv_bfloat16 b = __float2bfloat16_rn(x);
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(saddr), "h"(*reinterpret_cast<const uint16_t*>(&b)) : "memory");
}
__device__ __forceinline__ void st_bf16(uint8_t* base, uint32_t off, float x) {
    *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16_rn(x);
}

// ---------------------------------------------------------------------------------------------
// S = columns per row of theta: 1 (forward only) or 4 (primal + 3 tangents)
// ---------------------------------------------------------------------------------------------
// ---- prologue helpers: polynomial features of one column, 4 threads per column (fg = 0..3), features fg, fg+4, ...
#define TC_FPT 16                  // features per thread (F0 <= 64)
__device__ __forceinline__ float tc_sel(const float (&xf)[BT_MAX_D + 1], int e) {   // xf[e + 1], or 1 when e < 0
    float r = 1.f;
#pragma unroll
    for (int i = 1; i <= BT_MAX_D; ++i) if (i == e + 1) r = xf[i];
    return r;
}
// full parameter vector of the row of theta behind a column (global loads: issue early, consume late)
__device__ __forceinline__ void tc_load_column(const FmapDev& fm, const float* __restrict__ rows, int M, int row,
                                               float (&xf)[BT_MAX_D + 1]) {
#pragma unroll
    for (int i = 0; i <= BT_MAX_D; ++i) xf[i] = (i < fm.d_full) ? fm.fixed[i] : 0.f;
    if (row < M) {
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d)
            if (d < fm.D) {
                const float v = __ldg(rows + (size_t)row * fm.D + d);
#pragma unroll
                for (int i = 0; i <= BT_MAX_D; ++i)
                    if (i == fm.full_idx[d]) xf[i] += v;
            }
    }
}
// standardised monomials (s == 0) or their directional derivative along NN input tin (s > 0); order <= 3
__device__ __forceinline__ void tc_poly_feats(const int* pexp_s, const float* pmean_s, const float* pistd_s, int F0, int fg,
                                              int s, int tin, bool valid, const float (&xf)[BT_MAX_D + 1], float (&feat)[TC_FPT]) {
#pragma unroll
    for (int qf = 0; qf < TC_FPT; ++qf) {
        const int f = fg + 4 * qf;
        float val = 0.f;
        if (f < F0 && valid) {
            const int e0 = pexp_s[f * 4], e1 = pexp_s[f * 4 + 1], e2 = pexp_s[f * 4 + 2];
            const float a0 = tc_sel(xf, e0), a1 = tc_sel(xf, e1), a2 = tc_sel(xf, e2);
            if (s == 0) val = (a0 * a1 * a2 - pmean_s[f]) * pistd_s[f];
            else val = ((e0 == tin ? a1 * a2 : 0.f) + (e1 == tin ? a0 * a2 : 0.f) + (e2 == tin ? a0 * a1 : 0.f)) * pistd_s[f];
        }
        feat[qf] = val;
    }
}

// CS = thread-block cluster size: the CS CTAs of a cluster walk the weight stream in lock-step, each fetches 1/CS of
// every weight tile from L2 and multicasts it to all of them (L2 -> SM traffic / CS).
template <int S, int CS>
__global__ void __launch_bounds__(TC_THREADS, 1)
mlp_tc_kernel(TcSched sch, NetDev net, FmapDev fm, const float* __restrict__ rows, int M, float* __restrict__ lf,
              float* __restrict__ jac) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* act = smem;                                             // [act_kb][64 x 64 bf16, SW128]
    uint8_t* wst = smem + (size_t)sch.act_kb * TC_KB_BYTES;          // [TC_STAGES][128 x 64 bf16, SW128]
    uint8_t* misc = wst + TC_STAGES * TC_WTILE_BYTES;
    float* x0s = reinterpret_cast<float*>(misc);                     // [2][64] theta_full[0] of every column (ping-pong over tiles)
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc + 256);
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(misc + 256 + 64 * 8);
    float* xin_s = reinterpret_cast<float*>(misc + 1024);            // [BT_MAX_D + 1][64] full parameter vector per column
    float* pmean_s = xin_s + (BT_MAX_D + 1) * TC_BN;                 // [F0] (F0 <= 64)
    float* pistd_s = pmean_s + 64;                                   // [F0]
    int* pexp_s = reinterpret_cast<int*>(pistd_s + 64);              // [F0][4] exponent table (-1 padded)
    const uint32_t bar_full = smem_u32(bars + 0), bar_empty = smem_u32(bars + TC_STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * TC_STAGES), bar_tempty = smem_u32(bars + 2 * TC_STAGES + TC_SLOTS);
    const uint32_t bar_lready = smem_u32(bars + 2 * TC_STAGES + 2 * TC_SLOTS);   // [n_layers]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int P = TC_BN / S;                                     // rows of theta per tile
    const int ntiles = (M + P - 1) / P;
    const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t cmask = (uint16_t)((1u << CS) - 1u);
    // lock-step tile loop: every CTA of a cluster runs the same number of iterations (tiles >= ntiles are empty)
    const int n_ctile = (ntiles + CS - 1) / CS, ctile0 = blockIdx.x / CS, ctile_step = gridDim.x / CS;
#define TC_TILE_LOOP(var) for (int ct_ = ctile0, var = ctile0 * CS + (int)crank; ct_ < n_ctile; ct_ += ctile_step, var += ctile_step * CS)

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, CS); }
        for (int i = 0; i < TC_SLOTS; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 8); }
        mbar_init(bar_lready, 8);
        for (int l = 1; l < sch.n_layers; ++l) mbar_init(bar_lready + 8 * l, 8 * sch.L[l - 1].n_mt);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // zero the activation buffer once: padded / not-yet-written features must be finite (they meet zero weights)
        uint4* a4 = reinterpret_cast<uint4*>(act);
        const int n16 = (sch.act_kb * TC_KB_BYTES + TC_STAGES * TC_WTILE_BYTES) / 16;   // activations + weight stages
        for (int i = threadIdx.x; i < n16; i += TC_THREADS) a4[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    for (int i = threadIdx.x; i < net.F0; i += TC_THREADS) {
        pmean_s[i] = net.mean[i];
        pistd_s[i] = net.inv_std[i];
        for (int j = 0; j < 4; ++j) pexp_s[i * 4 + j] = j < net.order ? net.exps[i * net.order + j] : -1;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();                                  // peers' barriers are initialised before any multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp == 0) {
        // ===================== weight producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            TC_TILE_LOOP(tile)
                for (int l = 0; l < sch.n_layers; ++l) {
                    const TcLayer L = sch.L[l];
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(sch.wstream) + (size_t)L.w_off * 1024;
                    for (int mt = 0; mt < L.n_mt; ++mt) {
                        const uint32_t bytes = (uint32_t)tc_tile_kb(L, mt) * 1024u;
                        for (int kb = 0; kb < L.n_kb; ++kb, ++it, src += bytes) {
                            const uint32_t stage = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                            mbar_spin(bar_empty + 8 * stage, ph ^ 1u);
                            mbar_expect_tx(bar_full + 8 * stage, bytes);
                            if (CS == 1) {
                                bulk_g2s(smem_u32(wst + stage * TC_WTILE_BYTES), src, bytes, bar_full + 8 * stage);
                            } else {
                                const uint32_t part = bytes / CS;       // 1 KB groups / CS: multiple of 16 B
                                bulk_g2s_mc(smem_u32(wst + stage * TC_WTILE_BYTES) + crank * part, src + crank * part, part,
                                            bar_full + 8 * stage, cmask);
                            }
                        }
                    }
                }
        }
    } else if (warp == 1 || warp == 2) {
        // ===================== MMA issuers =====================
        // Issuing a tcgen05.mma blocks the thread for the whole operand fetch (~85 cycles at M=128, N<=128) and the
        // per-K-block barrier wait + commit (~260 cycles) cannot overlap with it, so ONE thread keeps the tensor core
        // only ~50 % busy.  Two threads split every unit (block l, M-tile mt) along K: warp 1 takes the even K-blocks,
        // warp 2 the odd ones, each into its own accumulator slot; the epilogue adds the two partial sums.
        const int par = warp - 1;
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=64, M=128 (cute UMMA::InstrDescriptor)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            const uint32_t act_addr = smem_u32(act), wst_addr = smem_u32(wst);
            const uint64_t desc_base = umma_desc_sw128(0);
            uint32_t it = 0, q = 0, tcount = 0;
            TC_TILE_LOOP(tile) {
                const bool rec = sch.dbg && blockIdx.x == 0 && tcount == 1 && par == 0;
                if (rec) sch.dbg[0] = clock64();
                for (int l = 0; l < sch.n_layers; ++l) {
                    const TcLayer L = sch.L[l];
                    bool waited = false;
                    if (rec) sch.dbg[1 + 3 * l] = clock64();
                    for (int mt = 0; mt < L.n_mt; ++mt, ++q) {
                        const uint32_t slot = 2u * (q % (TC_SLOTS / 2)) + (uint32_t)par;
                        mbar_wait(bar_tempty + 8 * slot, ((q / (TC_SLOTS / 2)) & 1u) ^ 1u);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + slot * TC_BN;
                        uint32_t acc = 0;                                  // first MMA of this half-unit overwrites
                        for (int kb = 0; kb < L.n_kb; ++kb, ++it) {
                            if ((kb & 1) != par) continue;                 // the other issuer's K-block
                            if (!waited && kb >= L.first_new_kb) {        // features written by the previous block
                                if (rec) sch.dbg[2 + 3 * l] = clock64();
                                mbar_wait(bar_lready + 8 * l, tcount & 1u);
                                if (rec) sch.dbg[3 + 3 * l] = clock64();
                                waited = true;
                            }
                            const uint32_t stage = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                            mbar_spin(bar_full + 8 * stage, ph);
                            const int krem = L.in - kb * TC_BK;
                            const int nks = krem >= TC_BK ? 4 : (krem + 15) / 16;
                            // descriptors differ only in the 14-bit start-address field (16-byte units): +2 per k-step of 32 B
                            const uint64_t a_desc = desc_base + ((wst_addr + stage * TC_WTILE_BYTES) >> 4);
                            const uint64_t b_desc = desc_base + ((act_addr + kb * TC_KB_BYTES) >> 4);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                if (ks < nks) { umma_f16(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc, acc); acc = 1u; }
                            if (CS == 1) umma_commit(bar_empty + 8 * stage);   // frees the weight stage when the MMAs retire
                            else umma_commit_mc(bar_empty + 8 * stage, cmask);  // ... in every CTA of the cluster
                        }
                        umma_commit(bar_tfull + 8 * slot);                // accumulator complete -> epilogue
                    }
                }
                if (rec) sch.dbg[1 + 3 * sch.n_layers] = clock64();
                ++tcount;
            }
        }
    } else if (warp >= 4) {
        // ===================== prologue + epilogue (8 warps) =====================
        const int ew = warp - 4, qd = ew & 3, h = ew >> 2;
        uint32_t q = 0, tcount_e = 0;
        // prologue work distribution: thread (pcol, fg) computes features fg, fg + 4, ... of column pcol
        const int t256 = ew * 32 + lane, pcol = t256 & 63, fg = t256 >> 6;
        const int pp = pcol / S, ps = pcol - pp * S;
        const int ptin = (S > 1 && ps > 0) ? fm.tinput[ps - 1] : -1;
        const uint32_t act_s = smem_u32(act);
        float feat[TC_FPT], x0_next;
        {
            float xf0[BT_MAX_D + 1];
            const int tile_first = ctile0 * CS + (int)crank;
            tc_load_column(fm, rows, M, tile_first * P + pp, xf0);
            tc_poly_feats(pexp_s, pmean_s, pistd_s, net.F0, fg, ps, ptin, tile_first * P + pp < M, xf0, feat);
            x0_next = xf0[0];
        }
        // Hand-over of the activation buffer between tiles: the input features of tile k+1 are written by each warp
        // right after IT has seen the output-layer accumulators of tile k complete (all MMAs of tile k have then read
        // the buffer), before it drains them -- so block 0 of the next tile overlaps with the output epilogue of this one.
        auto publish_inputs = [&](uint32_t parity) {
#pragma unroll
            for (int qf = 0; qf < TC_FPT; ++qf) {
                const int f = fg + 4 * qf;
                if (f < net.F0) sts_bf16(act_s + (uint32_t)(f / TC_BK) * TC_KB_BYTES + sw128_offset(pcol, f % TC_BK), feat[qf]);
            }
            if (fg == 0) x0s[parity * 64 + pcol] = x0_next;          // read by the output-layer epilogue of that tile
            fence_proxy_async();                                      // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_lready);                   // inputs of block 0 ready
        };
        if (ctile0 < n_ctile) publish_inputs(0);
        TC_TILE_LOOP(tile) {
            const bool rec = sch.dbg && blockIdx.x == 0 && tcount_e == 1 && ew == 0 && lane == 0;
            if (rec) sch.dbg[64] = sch.dbg[65] = sch.dbg[66] = clock64();
            const bool has_next = ct_ + ctile_step < n_ctile;
            // ---- prefetch the rows of theta of my NEXT tile now (global latency hidden behind this tile's blocks) ...
            float xf_n[BT_MAX_D + 1];
            const int tile_next = tile + ctile_step * CS;
            const bool valid_n = tile_next * P + pp < M;
            tc_load_column(fm, rows, M, tile_next * P + pp, xf_n);

            for (int l = 0; l < sch.n_layers; ++l) {
                const TcLayer L = sch.L[l];
                for (int mt = 0; mt < L.n_mt; ++mt, ++q) {
                    const uint32_t slot = 2u * (q % (TC_SLOTS / 2));     // even-K partial sums; slot + 1 = odd-K partial sums
                    if (L.is_out && mt == 0) {                           // ... and turn them into features while the last block runs
                        tc_poly_feats(pexp_s, pmean_s, pistd_s, net.F0, fg, ps, ptin, valid_n, xf_n, feat);
                        x0_next = xf_n[0];
                    }
                    if (rec && mt == 0) sch.dbg[67 + 3 * l] = clock64();
                    mbar_wait(bar_tfull + 8 * slot, (q / (TC_SLOTS / 2)) & 1u);
                    mbar_wait(bar_tfull + 8 * (slot + 1), (q / (TC_SLOTS / 2)) & 1u);
                    if (rec && mt == 0) sch.dbg[68 + 3 * l] = clock64();
                    if (L.is_out && mt == L.n_mt - 1 && has_next) publish_inputs((tcount_e + 1u) & 1u);
                    tc_fence_after();
                    const int f = mt * TC_BM + qd * 32 + lane;           // output feature of this thread
                    if (mt * TC_BM + qd * 32 >= L.nnew) {                // no live feature in this warp's 32 lanes
                        if (lane == 0) {
                            mbar_arrive(bar_tempty + 8 * slot);
                            mbar_arrive(bar_tempty + 8 * (slot + 1));
                            if (!L.is_out) mbar_arrive(bar_lready + 8 * (l + 1));
                        }
                        continue;
                    }
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + slot * TC_BN + h * 32, v);
                    if (L.n_kb > 1) {                                    // odd K-blocks were accumulated by the second issuer
                        float w[32];
                        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (slot + 1) * TC_BN + h * 32, w);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += w[j];
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(bar_tempty + 8 * slot); mbar_arrive(bar_tempty + 8 * (slot + 1)); }   // slots free again
                    if (!L.is_out) {
                        const bool live = f < L.nnew;
                        const float b = live ? __ldg(sch.bias + L.bias_off + f) : 0.f;
                        const uint32_t k = (uint32_t)(L.in + f);
                        const uint32_t kk = k % TC_BK, chunk = kk >> 3;
                        const uint32_t dst = smem_u32(act) + (k / TC_BK) * TC_KB_BYTES + ((kk & 7u) << 1) + h * 4096;
                        // row n = h*32 + j of the K-block: (n>>3)*1024 + (n&7)*128 + ((chunk ^ (n&7)) << 4); the eight
                        // swizzled chunk offsets are hoisted out of the element loop
                        uint32_t swz[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) swz[r] = dst + (uint32_t)(r * 128) + ((chunk ^ (uint32_t)r) << 4);
                        if (live) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float val;
                                if (S == 1) {
                                    val = gelu_tc(v[j] + b);
                                } else {
                                    const int jp = j - (j % S);             // primal column of this row of theta
                                    const float pre = v[jp] + b;
                                    val = (j % S == 0) ? gelu_tc(pre) : gelu_prime_tc(pre) * v[j];
                                }
                                sts_bf16(swz[j & 7] + (uint32_t)((j >> 3) * 1024), val);
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_lready + 8 * (l + 1));
                        if (rec && mt == L.n_mt - 1) sch.dbg[69 + 3 * l] = clock64();
                    } else if (f < sch.n_out) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = h * 32 + j, p = col / S, s = col - p * S;
                            const int row = tile * P + p;
                            if (row < M) {
                                if (s == 0) lf[(size_t)row * sch.n_out + f] = v[j] + x0s[(tcount_e & 1u) * 64 + col];
                                else jac[((size_t)row * (S - 1) + (s - 1)) * sch.n_out + f] = v[j];
                            }
                        }
                    }
                }
            }
            ++tcount_e;
        }
    }
#undef TC_TILE_LOOP
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();          // no CTA leaves while a peer may still multicast into it / arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

static inline int tc_launch(TcNet& t, const FmapDev& fm, const NetDev& net, const float* rows, int M, float* lf, float* jac,
                            int sm_count, cudaStream_t stream) {
    if (!t.ready) return -1;
    const int S = jac ? 1 + fm.T : 1;
    if (S != 1 && S != 4) return -2;
    if (net.order > 3) return -4;
    const int P = TC_BN / S;
    const int ntiles = (M + P - 1) / P;
    long long* dbg = nullptr;
    static const bool want_dbg = getenv("BT_TC_DEBUG") != nullptr;
    if (want_dbg && ntiles > 2 * sm_count) {
        cudaMalloc(&dbg, 128 * sizeof(long long));
        cudaMemset(dbg, 0, 128 * sizeof(long long));
    }
    t.sched.dbg = dbg;
    static const int cs_env = getenv("BT_TC_CLUSTER") ? atoi(getenv("BT_TC_CLUSTER")) : 1;   // weight multicast over a cluster: no gain measured
    const int CS = (cs_env == 1 || cs_env == 2 || cs_env == 4) ? cs_env : 2;
    int g = ((ntiles + CS - 1) / CS) * CS;
    const int gmax = (sm_count / CS) * CS;
    if (g > gmax) g = gmax;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = t.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t err = cudaSuccess;
#define TC_LAUNCH(S_, CS_)                                                                                              \
    do {                                                                                                                \
        err = cudaFuncSetAttribute(mlp_tc_kernel<S_, CS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_bytes); \
        if (err == cudaSuccess) err = cudaLaunchKernelEx(&cfg, mlp_tc_kernel<S_, CS_>, t.sched, net, fm, rows, M, lf, jac); \
    } while (0)
    if (S == 1) { if (CS == 1) TC_LAUNCH(1, 1); else if (CS == 2) TC_LAUNCH(1, 2); else TC_LAUNCH(1, 4); }
    else        { if (CS == 1) TC_LAUNCH(4, 1); else if (CS == 2) TC_LAUNCH(4, 2); else TC_LAUNCH(4, 4); }
#undef TC_LAUNCH
    if (err != cudaSuccess) return -5;
    if (dbg) {
        long long h[128];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        t.sched.dbg = nullptr;
        const long long t0 = h[0];
        fprintf(stderr, "[tc timeline S=%d M=%d] MMA thread (cycles since tile start): ", S, M);
        for (int l = 0; l < t.sched.n_layers; ++l)
            fprintf(stderr, "L%d start %lld wait %lld..%lld | ", l, h[1 + 3 * l] - t0, h[2 + 3 * l] - t0, h[3 + 3 * l] - t0);
        fprintf(stderr, "end %lld\n[tc timeline] epilogue warp4: enter %lld prologue-start %lld prologue-done %lld | ",
                h[1 + 3 * t.sched.n_layers] - t0, h[64] - t0, h[65] - t0, h[66] - t0);
        for (int l = 0; l < t.sched.n_layers; ++l)
            fprintf(stderr, "L%d wait %lld..%lld done %lld | ", l, h[67 + 3 * l] - t0, h[68 + 3 * l] - t0, h[69 + 3 * l] - t0);
        fprintf(stderr, "\n");
    }
    return 0;
}
