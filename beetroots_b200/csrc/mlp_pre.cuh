// Pre-pass of the tensor-core forward map: the leading small blocks of the dense-concat network.
//
// The four smallest blocks (17-57 output rows) are 1.6 % of the FLOPs but each of them costs the pair kernel
// (mlp_tc3.cuh) a full step of its per-unit dependency chain (~13 k of ~60 k cycles per tile).  This
// kernel evaluates them for ALL columns up front, with every tile independent of every other one, and writes, per
// pair-tile and CTA, the exact SWIZZLE_128B shared-memory image of K-blocks [0, pre_kb) of the activation buffer
// (monomials + the outputs of the leading blocks, bf16, stored as 2 x GELU like everywhere else); the pair kernel
// copies that image in and starts at block pre_blocks.
//
// Variant A (this file, first version): CUDA-core float32 arithmetic, one CTA per 32 columns -- the reference for the
// tensor-core version of the pre-pass and the fallback for network shapes it does not take.
#pragma once
#include "mlp_tc3.cuh"

#define PRE_COLS 32
#define PRE_THREADS 256

// round-to-nearest-even float -> bf16 (what cvt.rn.bf16.f32 and the host-side f32_to_bf16_rne do) and back
__device__ __forceinline__ uint16_t pre_bf16(float x) {
    const uint32_t u = __float_as_uint(x);
    return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
__device__ __forceinline__ float pre_rn(float x) { return __uint_as_float((uint32_t)pre_bf16(x) << 16); }

template <bool WITH_JAC>
__global__ void __launch_bounds__(PRE_THREADS)
mlp_pre_simt_kernel(NetDev net, FmapDev fm, int n_pre, int pre_feats, int pre_kb, const float* __restrict__ rows, int M,
                    int ntiles, uint8_t* __restrict__ img) {
    extern __shared__ float pre_smem[];
    float* H = pre_smem;                                   // [pre_kb * 64][32], features beyond pre_feats stay 0
    float* xin = pre_smem + (size_t)pre_kb * TC_BK * PRE_COLS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = PRE_THREADS / 32;
    const int S = WITH_JAC ? 1 + fm.T : 1;
    const int P = PRE_COLS / S;
    const int p = lane / S, s = lane - p * S;
    const int FT = pre_kb * TC_BK;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row = tile * P + p;
        const bool valid = row < M;
        __syncthreads();
        if (warp == 0) {
            float xf[BT_MAX_D + 1];
            tc_load_column(fm, rows, M, row, xf);
#pragma unroll
            for (int i = 0; i <= BT_MAX_D; ++i)
                if (i < fm.d_full) xin[i * PRE_COLS + lane] = xf[i];
        }
        for (int f = pre_feats + warp; f < FT; f += nwarps) H[f * PRE_COLS + lane] = 0.f;
        __syncthreads();
        const int tin = (WITH_JAC && s > 0) ? fm.tinput[s - 1] : -1;
        for (int f = warp; f < net.F0; f += nwarps) {
            float val;
            if (s == 0) {
                val = 1.f;
                for (int j = 0; j < net.order; ++j) {
                    const int e = net.exps[f * net.order + j];
                    if (e >= 0) val *= xin[(e + 1) * PRE_COLS + lane];
                }
                val = (val - net.mean[f]) * net.inv_std[f];
            } else {
                val = 0.f;
                for (int j = 0; j < net.order; ++j) {
                    if (net.exps[f * net.order + j] != tin) continue;
                    float term = 1.f;
                    for (int j2 = 0; j2 < net.order; ++j2) {
                        const int e2 = net.exps[f * net.order + j2];
                        if (j2 != j && e2 >= 0) term *= xin[(e2 + 1) * PRE_COLS + lane];
                    }
                    val += term;
                }
                val *= net.inv_std[f];
            }
            // the pair kernel rounds its inputs to bf16 when it stores them: do the same so that both paths see the same features
            H[f * PRE_COLS + lane] = valid ? pre_rn(val) : 0.f;
        }
        __syncthreads();
        for (int l = 0; l < n_pre; ++l) {
            const int K = net.in_[l], NEW = net.new_[l], LD = net.ld[l];
            const float* __restrict__ Wt = net.Wt[l];
            for (int o = warp; o < NEW; o += nwarps) {
                float acc = (s == 0) ? __ldg(net.bias[l] + o) : 0.f;
                for (int k = 0; k < K; ++k) {
                    // activations are stored as 2 x GELU (the factor 1/2 lives in the weights)
                    const float w = __ldg(Wt + (size_t)k * LD + o) * (k >= net.F0 ? 0.5f : 1.f);
                    acc = fmaf(pre_rn(w), H[k * PRE_COLS + lane], acc);
                }
                const float prim = __shfl_sync(0xffffffffu, acc, p * S);
                const float val = (s == 0) ? gelu2_tc(acc) : gelu2_prime_tc(prim) * acc;
                H[(K + o) * PRE_COLS + lane] = valid ? pre_rn(val) : 0.f;
            }
            __syncthreads();
        }
        // image: column gc of the whole launch -> pair-tile gc / 128, CTA (gc % 128) / 64, row gc % 64
        const long long gc = (long long)tile * PRE_COLS + lane;
        const long long pt = gc / TC2_BN;
        const int c = (int)(gc - pt * TC2_BN), cta = c / TC_BN, r = c % TC_BN;
        uint8_t* dst = img + ((size_t)pt * 2 + cta) * (size_t)pre_kb * TC_KB_BYTES;
        for (int f = warp; f < FT; f += nwarps) {
            const uint32_t off = (uint32_t)(f / TC_BK) * TC_KB_BYTES + sw128_offset((uint32_t)r, (uint32_t)(f % TC_BK));
            *reinterpret_cast<uint16_t*>(dst + off) = pre_bf16(H[f * PRE_COLS + lane]);
        }
    }
}

// img must hold ceil(M * S / 128) pair-tiles x 2 CTAs x pre_kb x 8 KB
static inline size_t pre_image_bytes(int M, int S, int pre_kb) {
    const size_t n_ptiles = ((size_t)M * S + TC2_BN - 1) / TC2_BN;
    return n_ptiles * 2 * (size_t)pre_kb * TC_KB_BYTES;
}

static inline int pre_launch(const Tc3Net& t, const FmapDev& fm, const NetDev& net, const float* rows, int M, bool jac,
                             uint8_t* img, int sm_count, cudaStream_t stream) {
    const int S = jac ? 1 + fm.T : 1;
    if (PRE_COLS % S) return -1;
    const int P = PRE_COLS / S;
    const int ntiles = (M + P - 1) / P;
    // the image of a partially filled last pair-tile must still be defined: columns beyond M*S are written as zeros by
    // the tiles that cover them (valid == false), so round the tile count up to whole pair-tiles
    const int per_ptile = TC2_BN / PRE_COLS;
    const int ntiles_full = (ntiles + per_ptile - 1) / per_ptile * per_ptile;
    const size_t smem = ((size_t)t.sched.pre_kb * TC_BK + BT_MAX_D + 1) * PRE_COLS * sizeof(float);
    int grid = ntiles_full < sm_count * 4 ? ntiles_full : sm_count * 4;
    if (jac) {
        cudaFuncSetAttribute(mlp_pre_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mlp_pre_simt_kernel<true><<<grid, PRE_THREADS, smem, stream>>>(net, fm, t.sched.pre_blocks, t.sched.pre_feats, t.sched.pre_kb, rows, M, ntiles_full, img);
    } else {
        cudaFuncSetAttribute(mlp_pre_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mlp_pre_simt_kernel<false><<<grid, PRE_THREADS, smem, stream>>>(net, fm, t.sched.pre_blocks, t.sched.pre_feats, t.sched.pre_kb, rows, M, ntiles_full, img);
    }
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -2;
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant B: warp-level tensor-core pre-pass.  One warp owns 32 columns at a time and runs the leading blocks on them
// with mma.sync.m16n8k16 (bf16 x bf16 -> f32), activations of its 32 columns in a private shared-memory tile
// H[32][192] (row = column, K-major, pitch 400 B: ldmatrix conflict-free), the weights of the leading blocks (34 KB,
// zero padded to 16-column / 16-row multiples) resident in shared memory for the whole kernel.  Orientation
// D[columns, features] = H[columns, k] x W^T[k, features]: a thread ends up with two adjacent features of one column.
// No inter-warp synchronisation after the weights are loaded: every warp is an independent pipeline, 8 per SM.
#define PREB_WARPS 14
#define PREB_PITCH 400                 // bytes per column row of H: 192 bf16 + 16 B
#define PREB_MAX 4
struct PreNetDev {
    int n_pre, F0, pre_kb;
    int in_[PREB_MAX], new_[PREB_MAX], npad[PREB_MAX], kpad[PREB_MAX], woff[PREB_MAX], wpitch[PREB_MAX], boff[PREB_MAX];
    int wbytes, nbias;
    int one_pos;                       // >= 0: features one_pos, one_pos + 1 of a primal column are the constant 1 (bias inputs of mlp_tc4.cuh)
    const uint8_t* wimg;               // per block [npad][wpitch] bf16, K-major, rows >= new and k >= in zero
    const float* bias;                 // per block [npad], zero padded
};
struct PreNet {
    bool ready = false;
    PreNetDev dev{};
    void *d_w = nullptr, *d_b = nullptr;
};
static inline void preb_free(PreNet& p) {
    if (p.d_w) cudaFree(p.d_w);
    if (p.d_b) cudaFree(p.d_b);
    p = PreNet();
}
static inline int preb_build(PreNet& p, int n_feat0, int n_pre, const int32_t* block_new, const double* const* weights,
                             const double* const* biases, int one_pos = -1) {
    p.ready = false;
    if (n_pre < 1 || n_pre > PREB_MAX || n_feat0 > 64) return 0;
    PreNetDev& d = p.dev;
    memset(&d, 0, sizeof d);
    d.n_pre = n_pre; d.F0 = n_feat0; d.one_pos = one_pos;
    int in = n_feat0, wb = 0, nb = 0;
    for (int l = 0; l < n_pre; ++l) {
        d.in_[l] = in;
        d.new_[l] = block_new[l];
        d.npad[l] = (block_new[l] + 15) / 16 * 16;
        d.kpad[l] = (in + 15) / 16 * 16;
        d.wpitch[l] = d.kpad[l] * 2 + 16;
        d.woff[l] = wb; d.boff[l] = nb;
        if (in + d.npad[l] > 192 || d.kpad[l] > 192) return 0;       // padded outputs must stay inside the 3 K-blocks
        wb += d.npad[l] * d.wpitch[l];
        nb += d.npad[l];
        in += block_new[l];
    }
    if (in > 192 || (one_pos >= 0 && (one_pos < in || one_pos + 2 > 192))) return 0;
    d.pre_kb = ((one_pos >= 0 ? one_pos + 2 : in) + TC_BK - 1) / TC_BK;
    d.wbytes = wb; d.nbias = nb;
    std::vector<uint8_t> w(wb, 0);
    std::vector<float> b(nb, 0.f);
    for (int l = 0; l < n_pre; ++l)
        for (int o = 0; o < block_new[l]; ++o) {
            b[d.boff[l] + o] = (float)biases[l][o];
            uint16_t* row = reinterpret_cast<uint16_t*>(w.data() + d.woff[l] + (size_t)o * d.wpitch[l]);
            for (int k = 0; k < d.in_[l]; ++k)       // activations at k >= n_feat0 are stored as 2 x GELU: halve those weights (exact)
                row[k] = f32_to_bf16_rne((float)((k >= n_feat0 ? 0.5 : 1.0) * weights[l][(size_t)o * d.in_[l] + k]));
        }
    if (cudaMalloc(&p.d_w, wb) != cudaSuccess || cudaMalloc(&p.d_b, nb * sizeof(float)) != cudaSuccess) return -1;
    if (cudaMemcpy(p.d_w, w.data(), wb, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (cudaMemcpy(p.d_b, b.data(), nb * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    d.wimg = (const uint8_t*)p.d_w; d.bias = (const float*)p.d_b;
    p.ready = true;
    return 0;
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int S>
__global__ void __launch_bounds__(PREB_WARPS * 32)
mlp_pre_mma_kernel(PreNetDev pn, NetDev net, FmapDev fm, const float* __restrict__ rows, int M, int n_groups,
                   uint8_t* __restrict__ img) {
    extern __shared__ __align__(16) uint8_t preb_smem[];
    uint8_t* wsm = preb_smem;                                                    // weights
    float* bsm = reinterpret_cast<float*>(preb_smem + ((pn.wbytes + 15) & ~15));  // biases
    float* pmean = bsm + ((pn.nbias + 3) & ~3);
    float* pistd = pmean + 64;
    int* pexp = reinterpret_cast<int*>(pistd + 64);                              // [64][4]
    float* xtab = reinterpret_cast<float*>(pexp + 256);                          // [warp][6][32]: 1, x_0..x_4 of every column (lane)
    uint8_t* Hall = reinterpret_cast<uint8_t*>(xtab + PREB_WARPS * 6 * 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* Hw = Hall + (size_t)warp * 32 * PREB_PITCH;
    for (int i = threadIdx.x; i < pn.wbytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(pn.wimg) + i);
    for (int i = threadIdx.x; i < pn.nbias; i += blockDim.x) bsm[i] = __ldg(pn.bias + i);
    for (int i = threadIdx.x; i < net.F0; i += blockDim.x) {
        pmean[i] = net.mean[i];
        pistd[i] = net.inv_std[i];
        for (int j = 0; j < 4; ++j) pexp[i * 4 + j] = j < net.order ? net.exps[i * net.order + j] : -1;
    }
    for (int i = threadIdx.x; i < PREB_WARPS * 32 * PREB_PITCH / 16; i += blockDim.x) reinterpret_cast<uint4*>(Hall)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint32_t Hw_s = smem_u32(Hw), w_s = smem_u32(wsm);
    const int g = lane >> 2, t = lane & 3;
    for (int grp = blockIdx.x * PREB_WARPS + warp; grp < n_groups; grp += gridDim.x * PREB_WARPS) {
        // ---- inputs: lane = column
        {
            const long long gc = (long long)grp * 32 + lane;
            const int row = (int)(gc / S), s = (int)(gc - (long long)row * S);
            const int tin = (S > 1 && s > 0) ? fm.tinput[s - 1] : -1;
            float xf[BT_MAX_D + 1];
            tc_load_column(fm, rows, M, row, xf);
            const bool valid = row < M;
            // x_e of my column at xt[(e + 1) * 32], the constant 1 (exponent slot unused, e = -1) at xt[0]
            float* xt = xtab + (size_t)warp * 6 * 32 + lane;
            xt[0] = 1.f;
#pragma unroll
            for (int i = 0; i < 5; ++i) xt[(i + 1) * 32] = xf[i + 1];
            uint16_t* hrow = reinterpret_cast<uint16_t*>(Hw + (size_t)lane * PREB_PITCH);
            for (int f = 0; f < pn.F0; ++f) {
                const int4 e = *reinterpret_cast<const int4*>(pexp + f * 4);       // warp-uniform
                const float a0 = xt[(e.x + 1) * 32], a1 = xt[(e.y + 1) * 32], a2 = xt[(e.z + 1) * 32];
                float val;
                if (S == 1 || s == 0) val = (a0 * a1 * a2 - pmean[f]) * pistd[f];
                else val = ((e.x == tin ? a1 * a2 : 0.f) + (e.y == tin ? a0 * a2 : 0.f) + (e.z == tin ? a0 * a1 : 0.f)) * pistd[f];
                hrow[f] = pre_bf16(valid ? val : 0.f);
            }
        }
        __syncwarp();
        // ---- leading blocks
        for (int l = 0; l < pn.n_pre; ++l) {
            const int in = pn.in_[l], ksteps = pn.kpad[l] >> 4, npairs = pn.npad[l] >> 4;
            const uint32_t wl = w_s + (uint32_t)pn.woff[l], wp = (uint32_t)pn.wpitch[l];
            const float* bl = bsm + pn.boff[l];
            const int n_new = pn.new_[l];
            for (int np = 0; np < npairs; ++np) {
                const bool nt1_live = np * 16 + 8 < n_new;           // second 8-feature tile of the pair holds live rows
                float acc[2][2][4];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[a][b][q] = 0.f;
                for (int ks = 0; ks < ksteps; ++ks) {
                    uint32_t a0[4], a1[4], bq[4];
                    const uint32_t aoff = (uint32_t)(lane & 15) * PREB_PITCH + (uint32_t)(ks * 16 + (lane >> 4) * 8) * 2u;
                    ldsm_x4(Hw_s + aoff, a0[0], a0[1], a0[2], a0[3]);
                    ldsm_x4(Hw_s + 16u * PREB_PITCH + aoff, a1[0], a1[1], a1[2], a1[3]);
                    ldsm_x4(wl + (uint32_t)(np * 16 + (lane & 7) + ((lane >> 4) << 3)) * wp + (uint32_t)(ks * 16 + ((lane >> 3) & 1) * 8) * 2u,
                            bq[0], bq[1], bq[2], bq[3]);
                    mma_bf16_16816(acc[0][0], a0, bq[0], bq[1]);
                    mma_bf16_16816(acc[1][0], a1, bq[0], bq[1]);
                    if (nt1_live) {
                        mma_bf16_16816(acc[0][1], a0, bq[2], bq[3]);
                        mma_bf16_16816(acc[1][1], a1, bq[2], bq[3]);
                    }
                }
                // ---- bias, activation, store: acc[mt][nt] = {(col g, f, f+1), (col g + 8, f, f+1)}, f = np*16 + nt*8 + 2t
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        if (nt == 1 && !nt1_live) continue;
                        const int fo = np * 16 + nt * 8 + 2 * t;
                        const float b0 = bl[fo], b1 = bl[fo + 1];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {                   // h = 0: column g, h = 1: column g + 8
                            const int col = mt * 16 + g + 8 * h;
                            float v0 = acc[mt][nt][2 * h], v1 = acc[mt][nt][2 * h + 1];
                            float o0, o1;
                            if (S == 1) {
                                o0 = gelu2_tc(v0 + b0);
                                o1 = gelu2_tc(v1 + b1);
                            } else {
                                // tangent columns: GELU'(primal pre-activation) x tangent pre-activation; the primal of column
                                // c is column c - c % S: same (mt, h), lane with g' = g - g % S
                                const int sc = col % S;
                                const int src = lane - 4 * (g % S);
                                const float p0 = __shfl_sync(0xffffffffu, v0 + b0, src), p1 = __shfl_sync(0xffffffffu, v1 + b1, src);
                                o0 = sc == 0 ? gelu2_tc(v0 + b0) : gelu2_prime_tc(p0) * v0;
                                o1 = sc == 0 ? gelu2_tc(v1 + b1) : gelu2_prime_tc(p1) * v1;
                            }
                            uint16_t* dst = reinterpret_cast<uint16_t*>(Hw + (size_t)col * PREB_PITCH) + in + fo;
                            uint32_t pk;                                  // {o1, o0} as bf16x2, round to nearest even
                            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(o1), "f"(o0));
                            dst[0] = (uint16_t)pk;                      // (in + fo may be odd: two 2-byte stores)
                            dst[1] = (uint16_t)(pk >> 16);
                        }
                    }
            }
            __syncwarp();
        }
        if (pn.one_pos >= 0) {     // constant-one features of the primal columns (after the blocks: their padded rows wrote zeros here)
            const long long gc = (long long)grp * 32 + lane;
            const int row = (int)(gc / S), s = (int)(gc - (long long)row * S);
            uint16_t* hrow = reinterpret_cast<uint16_t*>(Hw + (size_t)lane * PREB_PITCH);
            hrow[pn.one_pos] = hrow[pn.one_pos + 1] = (s == 0 && row < M) ? (uint16_t)0x3F80 : (uint16_t)0;
            __syncwarp();
        }
        // ---- image of my 32 columns: pair-tile / CTA half / rows of the SWIZZLE_128B K-block tiles
        {
            const long long gc0 = (long long)grp * 32;
            const long long pt = gc0 / TC2_BN;
            const int c0 = (int)(gc0 - pt * TC2_BN), cta = c0 / TC_BN, r0 = c0 % TC_BN;
            uint8_t* dst = img + ((size_t)pt * 2 + cta) * (size_t)pn.pre_kb * TC_KB_BYTES;
            for (int kb = 0; kb < pn.pre_kb; ++kb)
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int idx = it * 32 + lane, rl = idx >> 3, ch = idx & 7;
                    const uint4 v = *reinterpret_cast<const uint4*>(Hw + (size_t)rl * PREB_PITCH + kb * 128 + ch * 16);
                    const int r = r0 + rl;
                    *reinterpret_cast<uint4*>(dst + (size_t)kb * TC_KB_BYTES + (size_t)r * 128 + ((ch ^ (r & 7)) << 4)) = v;
                }
        }
        __syncwarp();
    }
}

static inline size_t preb_smem_bytes(const PreNetDev& d) {
    return ((size_t)(d.wbytes + 15) & ~(size_t)15) + (size_t)((d.nbias + 3) & ~3) * 4 + 128 * 4 + 256 * 4 + (size_t)PREB_WARPS * 6 * 32 * 4 + (size_t)PREB_WARPS * 32 * PREB_PITCH;
}
static inline int preb_launch(const PreNet& p, const FmapDev& fm, const NetDev& net, const float* rows, int M, bool jac,
                              uint8_t* img, int sm_count, cudaStream_t stream) {
    if (!p.ready) return -1;
    const int S = jac ? 1 + fm.T : 1;
    if (S != 1 && S != 4) return -2;
    if (net.order > 3) return -3;
    const long long cols = (long long)M * S;
    const long long n_ptiles = (cols + TC2_BN - 1) / TC2_BN;
    const int n_groups = (int)(n_ptiles * (TC2_BN / 32));
    const size_t smem = preb_smem_bytes(p.dev);
    int grid = (n_groups + PREB_WARPS - 1) / PREB_WARPS;
    if (grid > sm_count) grid = sm_count;
    cudaError_t e;
    if (S == 1) {
        e = cudaFuncSetAttribute(mlp_pre_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) mlp_pre_mma_kernel<1><<<grid, PREB_WARPS * 32, smem, stream>>>(p.dev, net, fm, rows, M, n_groups, img);
    } else {
        e = cudaFuncSetAttribute(mlp_pre_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) mlp_pre_mma_kernel<4><<<grid, PREB_WARPS * 32, smem, stream>>>(p.dev, net, fm, rows, M, n_groups, img);
    }
    if (e != cudaSuccess) return -4;
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -5;
}
