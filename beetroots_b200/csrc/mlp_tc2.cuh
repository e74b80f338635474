// Forward map kernel, tcgen05 cta_group::2 variant: a pair of SMs (thread-block cluster of 2) works on
// one tile of 128 columns.
//
// Why: with the activations resident in shared memory only 64 columns fit per SM, and a
// cta_group::1 MMA of N=64 is operand-fetch bound (85 cycles for 131 kMAC = 37 % of the tensor peak)
// while every weight byte streamed from L2 serves only 64 columns.  Pairing two SMs gives
//     D[256 out features, 128 columns] += W[256, k] x H[128, k]        (UMMA M=256, N=128, K=16)
// where CTA c supplies rows [128c, 128c+128) of W (its own 16 KB weight stage) and columns
// [64c, 64c+64) of H (its own resident activation buffer), and receives rows [128c, 128c+128) of D in
// its TMEM.  Per SM and per column: 1/2 of the MMA issue slots and 1/2 of the L2 -> SM weight bytes of the
// single-CTA kernel.  The price: CTA c's epilogue produces features for all 128 columns, so half of its
// bf16 stores go to the peer's shared memory (DSMEM), and the barriers that gate the leader's MMA
// (weights landed in the peer, accumulator drained, features of the previous block written) collect
// remote arrivals.
//
// Roles (per CTA): warp 0 = weight producer; warps 1, 2 = MMA issuers (leader CTA only, split-K as in
// mlp_tc.cuh); warp 3 = relay (peer CTA only: forwards "weight stage full" to the leader); warps 4..11 =
// prologue + epilogue.
#pragma once
#include "mlp_tc.cuh"

#define TC2_BN 128                 // columns per pair-tile (UMMA N)
#define TC2_BM 256                 // output features per unit (UMMA M), 128 per CTA
#define TC2_SLOTS 4                // accumulator slots of 128 TMEM columns: 2 units in flight x 2 split-K halves

struct Tc2Layer {
    int in, nnew, n_mt, n_kb, first_new_kb, bias_off, is_out;
    int w_off[2];                  // per CTA: offset (1 KB units) of this layer's first weight tile in its stream
};
struct Tc2Sched {
    int n_layers, act_kb, n_out;
    Tc2Layer L[TC_MAX_LAYERS];
    const uint8_t* wstream[2];
    const float* bias;
    long long* dbg;
};
struct Tc2Net {
    bool ready = false;
    Tc2Sched sched{};
    void* d_w[2] = {nullptr, nullptr};
    void* d_b = nullptr;
    size_t smem_bytes = 0;
};
static inline void tc2_free(Tc2Net& t) {
    for (int c = 0; c < 2; ++c) if (t.d_w[c]) cudaFree(t.d_w[c]);
    if (t.d_b) cudaFree(t.d_b);
    t = Tc2Net();
}
// live 8-row groups of CTA c's half of unit (L, mt)
__host__ __device__ __forceinline__ int tc2_tile_kb(const Tc2Layer& L, int mt, int c) {
    int rows = L.nnew - mt * TC2_BM - c * 128;
    rows = rows > 128 ? 128 : (rows < 0 ? 0 : rows);
    return (rows + 7) / 8;
}

static inline int tc2_build(Tc2Net& t, int n_feat0, int n_blocks, const int32_t* block_new, const double* const* weights,
                            const double* const* biases, const double* w_out, int n_out, double out_scale) {
    t.ready = false;
    if (n_blocks + 1 > TC_MAX_LAYERS || n_out > 32 || n_feat0 > 64) return 0;
    Tc2Sched& s = t.sched;
    memset(&s, 0, sizeof s);
    s.n_layers = n_blocks + 1;
    s.n_out = n_out;
    int w = n_feat0, prev_in = 0, bias_off = 0, stream_kb[2] = {0, 0};
    for (int l = 0; l <= n_blocks; ++l) {
        Tc2Layer& L = s.L[l];
        L.in = w;
        L.nnew = l < n_blocks ? block_new[l] : n_out;
        L.is_out = l == n_blocks;
        L.n_mt = (L.nnew + TC2_BM - 1) / TC2_BM;
        L.n_kb = (L.in + TC_BK - 1) / TC_BK;
        L.first_new_kb = l == 0 ? 0 : prev_in / TC_BK;
        L.bias_off = bias_off;
        bias_off += L.n_mt * TC2_BM;
        for (int c = 0; c < 2; ++c) {
            L.w_off[c] = stream_kb[c];
            for (int mt = 0; mt < L.n_mt; ++mt) stream_kb[c] += tc2_tile_kb(L, mt, c) * L.n_kb;
        }
        prev_in = w;
        if (l < n_blocks) w += block_new[l];
    }
    s.act_kb = s.L[n_blocks].n_kb;
    const size_t smem = (size_t)s.act_kb * TC_KB_BYTES + (size_t)TC_STAGES * TC_WTILE_BYTES + 1024 + 4096 + 1024;
    if (smem > 227 * 1024) return 0;
    t.smem_bytes = smem;
    std::vector<float> bs(bias_off, 0.f);
    for (int c = 0; c < 2; ++c) {
        std::vector<uint16_t> ws((size_t)stream_kb[c] * 512 + 8, 0);
        size_t pos_kb = 0;
        for (int l = 0; l <= n_blocks; ++l) {
            const Tc2Layer& L = s.L[l];
            const double* W = L.is_out ? w_out : weights[l];
            const double sc = L.is_out ? out_scale : 1.0;
            if (c == 0) for (int o = 0; o < L.nnew && !L.is_out; ++o) bs[L.bias_off + o] = (float)biases[l][o];
            for (int mt = 0; mt < L.n_mt; ++mt) {
                const int tkb = tc2_tile_kb(L, mt, c);
                for (int kb = 0; kb < L.n_kb; ++kb, pos_kb += tkb) {
                    uint16_t* dst = ws.data() + pos_kb * 512;
                    for (int r = 0; r < tkb * 8; ++r) {
                        const int o = mt * TC2_BM + c * 128 + r;
                        if (o >= L.nnew) continue;
                        for (int kk = 0; kk < TC_BK; ++kk) {
                            const int k = kb * TC_BK + kk;
                            if (k >= L.in) continue;
                            dst[sw128_offset(r, kk) / 2] = f32_to_bf16_rne((float)(sc * W[(size_t)o * L.in + k]));
                        }
                    }
                }
            }
        }
        if (cudaMalloc(&t.d_w[c], ws.size() * 2) != cudaSuccess) return -1;
        if (cudaMemcpy(t.d_w[c], ws.data(), ws.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
        s.wstream[c] = (const uint8_t*)t.d_w[c];
    }
    if (cudaMalloc(&t.d_b, bs.size() * 4 + 16) != cudaSuccess) return -1;
    if (cudaMemcpy(t.d_b, bs.data(), bs.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    s.bias = (const float*)t.d_b;
    t.ready = true;
    return 0;
}

// ---- cluster-scope PTX helpers -----------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
// remote arrive WITHOUT cluster-scope release: ptxas lowers .release.cluster to MEMBAR + CCTL.IVALL (L1 invalidate,
// measured ~3 k cycles per arrive here).  Ordering of the data is provided by fence.proxy.async issued by every
// writer before the warp's elected lane arrives (same pattern as cutlass ClusterBarrier::arrive(cta_id)).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_bf16(uint32_t cluster_addr, float x) {
    const __nv_bfloat16 b = __float2bfloat16_rn(x);
    asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(cluster_addr), "h"(*reinterpret_cast<const uint16_t*>(&b)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

template <int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
mlp_tc2_kernel(Tc2Sched sch, NetDev net, FmapDev fm, const float* __restrict__ rows, int M, float* __restrict__ lf,
               float* __restrict__ jac) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* act = smem;                                             // [act_kb][64 columns x 64 features bf16, SW128]
    uint8_t* wst = smem + (size_t)sch.act_kb * TC_KB_BYTES;          // [TC_STAGES][128 rows x 64 k bf16, SW128]
    uint8_t* misc = wst + TC_STAGES * TC_WTILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc + 256);
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(misc + 256 + 64 * 8);
    float* x0s2 = reinterpret_cast<float*>(misc + 1024);             // [2][128] theta_full[0] of the pair-tile's columns (leader's copy is read)
    float* pmean_s = x0s2 + (BT_MAX_D + 1) * 64;
    float* pistd_s = pmean_s + 64;
    int* pexp_s = reinterpret_cast<int*>(pistd_s + 64);
    // barriers (same offsets in both CTAs): full[3] empty[3] pfull[3] tfull[4] tempty[4] lready[n_layers]
    const uint32_t bar_full = smem_u32(bars + 0), bar_empty = smem_u32(bars + 3), bar_pfull = smem_u32(bars + 6);
    const uint32_t bar_tfull = smem_u32(bars + 9), bar_tempty = smem_u32(bars + 9 + TC2_SLOTS);
    const uint32_t bar_lready = smem_u32(bars + 9 + 2 * TC2_SLOTS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    constexpr int P2 = TC2_BN / S;                                   // rows of theta per pair-tile
    const int n_ptiles = (M + P2 - 1) / P2;
    const int pt0 = blockIdx.x >> 1, pt_step = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); mbar_init(bar_pfull + 8 * i, 1); }
        for (int i = 0; i < TC2_SLOTS; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 16); }
        mbar_init(bar_lready, 16);
        for (int l = 1; l < sch.n_layers; ++l) mbar_init(bar_lready + 8 * l, 16 * sch.L[l - 1].n_mt);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        uint4* a4 = reinterpret_cast<uint4*>(act);
        const int n16 = (sch.act_kb * TC_KB_BYTES + TC_STAGES * TC_WTILE_BYTES) / 16;
        for (int i = threadIdx.x; i < n16; i += TC_THREADS) a4[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    for (int i = threadIdx.x; i < net.F0; i += TC_THREADS) {
        pmean_s[i] = net.mean[i];
        pistd_s[i] = net.inv_std[i];
        for (int j = 0; j < 4; ++j) pexp_s[i * 4 + j] = j < net.order ? net.exps[i * net.order + j] : -1;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp == 0) {
        // ===================== weight producer (both CTAs, own half of every unit) =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int pt = pt0; pt < n_ptiles; pt += pt_step)
                for (int l = 0; l < sch.n_layers; ++l) {
                    const Tc2Layer L = sch.L[l];
                    const uint8_t* src = sch.wstream[crank] + (size_t)L.w_off[crank] * 1024;
                    for (int mt = 0; mt < L.n_mt; ++mt) {
                        const uint32_t bytes = (uint32_t)tc2_tile_kb(L, mt, (int)crank) * 1024u;
                        for (int kb = 0; kb < L.n_kb; ++kb, ++it, src += bytes) {
                            const uint32_t stage = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                            mbar_spin(bar_empty + 8 * stage, ph ^ 1u);
                            if (bytes) {
                                mbar_expect_tx(bar_full + 8 * stage, bytes);
                                bulk_g2s(smem_u32(wst + stage * TC_WTILE_BYTES), src, bytes, bar_full + 8 * stage);
                            } else {
                                mbar_arrive(bar_full + 8 * stage);          // nothing to fetch for this half
                            }
                        }
                    }
                }
        }
    } else if (warp == 3) {
        // ===================== relay (peer CTA): "my weight stage is full" -> leader =====================
        if (lane == 0 && crank == 1) {
            const uint32_t leader_pfull = mapa_u32(bar_pfull, 0);
            uint32_t it = 0;
            for (int pt = pt0; pt < n_ptiles; pt += pt_step)
                for (int l = 0; l < sch.n_layers; ++l) {
                    const int n = sch.L[l].n_mt * sch.L[l].n_kb;
                    for (int i = 0; i < n; ++i, ++it) {
                        const uint32_t stage = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                        mbar_spin(bar_full + 8 * stage, ph);
                        mbar_arrive_cluster(leader_pfull + 8 * stage);
                    }
                }
        }
    } else if (warp == 1 || warp == 2) {
        // ===================== MMA issuers (leader CTA only), split-K over two threads =====================
        const int par = warp - 1;
        if (lane == 0 && crank == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC2_BN >> 3) << 17) | ((uint32_t)(TC2_BM >> 4) << 24);
            const uint32_t act_addr = smem_u32(act), wst_addr = smem_u32(wst);
            const uint64_t desc_base = umma_desc_sw128(0);
            uint32_t it = 0, q = 0, tcount = 0;
            for (int pt = pt0; pt < n_ptiles; pt += pt_step, ++tcount) {
                const bool rec = sch.dbg && blockIdx.x == 0 && tcount == 1 && par == 0;
                if (rec) sch.dbg[0] = clock64();
                for (int l = 0; l < sch.n_layers; ++l) {
                    const Tc2Layer L = sch.L[l];
                    bool waited = false;
                    if (rec) sch.dbg[1 + 3 * l] = clock64();
                    for (int mt = 0; mt < L.n_mt; ++mt, ++q) {
                        const uint32_t slot = 2u * (q % (TC2_SLOTS / 2)) + (uint32_t)par;
                        mbar_wait(bar_tempty + 8 * slot, ((q / (TC2_SLOTS / 2)) & 1u) ^ 1u);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + slot * TC2_BN;
                        uint32_t acc = 0;
                        for (int kb = 0; kb < L.n_kb; ++kb, ++it) {
                            if ((kb & 1) != par) continue;
                            if (!waited && kb >= L.first_new_kb) {
                                if (rec) sch.dbg[2 + 3 * l] = clock64();
                                mbar_wait(bar_lready + 8 * l, tcount & 1u);
                                if (rec) sch.dbg[3 + 3 * l] = clock64();
                                waited = true;
                            }
                            const uint32_t stage = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                            mbar_spin(bar_full + 8 * stage, ph);                // my half of the weights
                            mbar_spin(bar_pfull + 8 * stage, ph);               // the peer's half
                            const int krem = L.in - kb * TC_BK;
                            const int nks = krem >= TC_BK ? 4 : (krem + 15) / 16;
                            const uint64_t a_desc = desc_base + ((wst_addr + stage * TC_WTILE_BYTES) >> 4);
                            const uint64_t b_desc = desc_base + ((act_addr + kb * TC_KB_BYTES) >> 4);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                if (ks < nks) { umma2_f16(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc, acc); acc = 1u; }
                            umma2_commit_mc(bar_empty + 8 * stage, 3);          // frees the stage in both CTAs
                        }
                        umma2_commit_mc(bar_tfull + 8 * slot, 3);               // accumulator half complete, both CTAs
                    }
                }
                if (rec) sch.dbg[1 + 3 * sch.n_layers] = clock64();
            }
        }
    } else if (warp >= 4) {
        // ===================== prologue + epilogue (8 warps per CTA) =====================
        const int ew = warp - 4, qd = ew & 3, h = ew >> 2;            // TMEM lane quarter, column half (= owner CTA of the columns)
        const uint32_t act_dst = mapa_u32(smem_u32(act), (uint32_t)h);  // activation buffer of the CTA owning my 64 columns
        const uint32_t l_tempty = mapa_u32(bar_tempty, 0), l_lready = mapa_u32(bar_lready, 0);
        uint32_t q = 0, tcount_e = 0;
        // prologue work distribution: thread (pcol, fg) computes features fg, fg + 4, ... of my column pcol
        const int t256 = ew * 32 + lane, pcol = t256 & 63, fg = t256 >> 6;
        const int gcol_p = (int)crank * 64 + pcol, pp = gcol_p / S, ps = gcol_p - pp * S;
        const int ptin = (S > 1 && ps > 0) ? fm.tinput[ps - 1] : -1;
        const uint32_t act_s = smem_u32(act);
        float feat[TC_FPT], x0_cur;
        const uint32_t x0_leader = mapa_u32(smem_u32(x0s2), 0);       // the output-layer epilogue runs in the leader CTA
        {
            float xf0[BT_MAX_D + 1];
            tc_load_column(fm, rows, M, pt0 * P2 + pp, xf0);
            tc_poly_feats(pexp_s, pmean_s, pistd_s, net.F0, fg, ps, ptin, pt0 * P2 + pp < M, xf0, feat);
            x0_cur = xf0[0];
        }
        for (int pt = pt0; pt < n_ptiles; pt += pt_step, ++tcount_e) {
            const bool rec = sch.dbg && blockIdx.x == 0 && tcount_e == 1 && ew == 0 && lane == 0;
            if (rec) sch.dbg[64] = clock64();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (rec) sch.dbg[65] = clock64();
            // ---- input features of MY 64 columns (computed one tile ahead, see below) -> my activation buffer
#pragma unroll
            for (int qf = 0; qf < TC_FPT; ++qf) {
                const int f = fg + 4 * qf;
                if (f < net.F0) sts_bf16(act_s + (uint32_t)(f / TC_BK) * TC_KB_BYTES + sw128_offset(pcol, f % TC_BK), feat[qf]);
            }
            if (fg == 0)                                                  // double-buffered: the peer may run one tile ahead
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(x0_leader + (uint32_t)(((tcount_e & 1u) * 128u + (uint32_t)gcol_p) * 4u)), "f"(x0_cur) : "memory");
            fence_proxy_async();                                          // local generic stores -> async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(l_lready);                 // inputs of block 0 ready (16 arrivals)
            if (rec) sch.dbg[66] = clock64();
            // ---- prefetch the rows of theta of the NEXT pair-tile (latency hidden behind this tile's blocks)
            float xf_n[BT_MAX_D + 1];
            const int row_n = (pt + pt_step) * P2 + pp;
            tc_load_column(fm, rows, M, row_n, xf_n);

            for (int l = 0; l < sch.n_layers; ++l) {
                const Tc2Layer L = sch.L[l];
                for (int mt = 0; mt < L.n_mt; ++mt, ++q) {
                    const uint32_t slot = 2u * (q % (TC2_SLOTS / 2));
                    if (L.is_out && mt == 0) {                               // features of the next tile, while the last block runs
                        tc_poly_feats(pexp_s, pmean_s, pistd_s, net.F0, fg, ps, ptin, row_n < M, xf_n, feat);
                        x0_cur = xf_n[0];
                    }
                    if (rec && mt == 0) sch.dbg[67 + 3 * l] = clock64();
                    mbar_wait(bar_tfull + 8 * slot, (q / (TC2_SLOTS / 2)) & 1u);
                    mbar_wait(bar_tfull + 8 * (slot + 1), (q / (TC2_SLOTS / 2)) & 1u);
                    if (rec && mt == 0) sch.dbg[68 + 3 * l] = clock64();
                    tc_fence_after();
                    const int fbase = mt * TC2_BM + (int)crank * 128 + qd * 32;
                    const int f = fbase + lane;                              // output feature of this thread
                    if (fbase >= L.nnew) {
                        if (lane == 0) {
                            mbar_arrive_cluster(l_tempty + 8 * slot);
                            mbar_arrive_cluster(l_tempty + 8 * (slot + 1));
                            if (!L.is_out) mbar_arrive_cluster(l_lready + 8 * (l + 1));
                        }
                        continue;
                    }
                    const bool live = f < L.nnew;
                    const float b = (live && !L.is_out) ? __ldg(sch.bias + L.bias_off + f) : 0.f;
                    const uint32_t k = (uint32_t)(L.in + f);
                    const uint32_t kk = k % TC_BK, chunk = kk >> 3;
                    // destination = activation buffer of the CTA that owns my 64 columns, addressed in the shared::cluster
                    // window (works for my own buffer too): one store path, eight hoisted swizzle offsets
                    const uint32_t dst = act_dst + (k / TC_BK) * TC_KB_BYTES + ((kk & 7u) << 1);
                    uint32_t swz[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) swz[r] = dst + (uint32_t)(r * 128) + ((chunk ^ (uint32_t)r) << 4);
#pragma unroll
                    const int fine = (rec && mt == 0 && (l == 0 || l == 8)) ? 100 + (l ? 8 : 0) : -1;
                    for (int half = 0; half < 2; ++half) {                   // my 64 columns = two 32-column TMEM loads
                        float v[32];
                        const uint32_t tcol = (uint32_t)(h * 64 + half * 32);
                        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + slot * TC2_BN + tcol, v);
                        if (L.n_kb > 1) {
                            float w[32];
                            tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (slot + 1) * TC2_BN + tcol, w);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += w[j];
                        }
                        if (fine >= 0) sch.dbg[fine + 2 * half] = clock64();
                        if (half == 1) {                                     // both halves read: the slots can be refilled
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) { mbar_arrive_cluster(l_tempty + 8 * slot); mbar_arrive_cluster(l_tempty + 8 * (slot + 1)); }
                        }
                        if (!L.is_out) {
                            if (live) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    float val;
                                    if (S == 1) {
                                        val = gelu_tc(v[j] + b);
                                    } else {
                                        const int jp = j - (j % S);
                                        const float pre = v[jp] + b;
                                        val = (j % S == 0) ? gelu_tc(pre) : gelu_prime_tc(pre) * v[j];
                                    }
                                    const int n = half * 32 + j;             // column inside the owner's 64-column buffer
                                    st_cluster_bf16(swz[n & 7] + (uint32_t)((n >> 3) * 1024), val);
                                }
                            }
                        } else if (f < sch.n_out) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int gcol = h * 64 + half * 32 + j, p = gcol / S, s = gcol - p * S;
                                const int row = pt * P2 + p;
                                if (row < M) {
                                    if (s == 0) {
                                        lf[(size_t)row * sch.n_out + f] = v[j] + x0s2[(tcount_e & 1u) * 128u + (uint32_t)gcol];
                                    } else {
                                        jac[((size_t)row * (S - 1) + (s - 1)) * sch.n_out + f] = v[j];
                                    }
                                }
                            }
                        }
                        if (fine >= 0) sch.dbg[fine + 2 * half + 1] = clock64();
                    }
                    if (!L.is_out) {
                        // .shared::cluster / unqualified forms lower to MEMBAR.ALL.GPU, which stalls the whole SM for
                        // thousands of cycles; the CTA-scoped form orders this thread's generic stores (local and
                        // DSMEM) before the async-proxy reads that follow the barrier hand-off
                        fence_proxy_async();
                        if (fine >= 0) sch.dbg[fine + 4] = clock64();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(l_lready + 8 * (l + 1));
                        if (fine >= 0) sch.dbg[fine + 5] = clock64();
                        if (rec && mt == L.n_mt - 1) sch.dbg[69 + 3 * l] = clock64();
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

static inline int tc2_launch(Tc2Net& t, const FmapDev& fm, const NetDev& net, const float* rows, int M, float* lf, float* jac,
                             int sm_count, cudaStream_t stream) {
    if (!t.ready) return -1;
    const int S = jac ? 1 + fm.T : 1;
    if (S != 1 && S != 4) return -2;
    if (net.order > 3) return -4;
    const int P2 = TC2_BN / S;
    const int n_ptiles = (M + P2 - 1) / P2;
    int g = 2 * n_ptiles;
    const int gmax = (sm_count / 2) * 2;
    if (g > gmax) g = gmax;
    long long* dbg = nullptr;
    static const bool want_dbg = getenv("BT_TC_DEBUG") != nullptr;
    if (want_dbg && n_ptiles > sm_count) {
        cudaMalloc(&dbg, 128 * sizeof(long long));
        cudaMemset(dbg, 0, 128 * sizeof(long long));
    }
    t.sched.dbg = dbg;
    cudaError_t err;
    if (S == 1) {
        err = cudaFuncSetAttribute(mlp_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_bytes);
        if (err == cudaSuccess) mlp_tc2_kernel<1><<<g, TC_THREADS, t.smem_bytes, stream>>>(t.sched, net, fm, rows, M, lf, jac);
    } else {
        err = cudaFuncSetAttribute(mlp_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_bytes);
        if (err == cudaSuccess) mlp_tc2_kernel<4><<<g, TC_THREADS, t.smem_bytes, stream>>>(t.sched, net, fm, rows, M, lf, jac);
    }
    if (err != cudaSuccess) return -5;
    if (dbg) {
        long long h[128];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        t.sched.dbg = nullptr;
        const long long t0 = h[0];
        fprintf(stderr, "[tc2 timeline S=%d M=%d] MMA thread: ", S, M);
        for (int l = 0; l < t.sched.n_layers; ++l)
            fprintf(stderr, "L%d start %lld wait %lld..%lld | ", l, h[1 + 3 * l] - t0, h[2 + 3 * l] - t0, h[3 + 3 * l] - t0);
        fprintf(stderr, "end %lld\n[tc2 timeline] epilogue warp4: enter %lld prologue %lld..%lld | ", h[1 + 3 * t.sched.n_layers] - t0,
                h[64] - t0, h[65] - t0, h[66] - t0);
        for (int l = 0; l < t.sched.n_layers; ++l)
            fprintf(stderr, "L%d wait %lld..%lld done %lld | ", l, h[67 + 3 * l] - t0, h[68 + 3 * l] - t0, h[69 + 3 * l] - t0);
        fprintf(stderr, "\n[tc2 fine] ");
        for (int i = 0; i < 2; ++i) {
            const long long w0 = h[68 + 3 * (i ? 8 : 0)];
            fprintf(stderr, "L%d: ld0 %lld st0 %lld ld1 %lld st1 %lld fence %lld arrive %lld | ", i ? 8 : 0, h[100 + 8 * i] - w0,
                    h[101 + 8 * i] - w0, h[102 + 8 * i] - w0, h[103 + 8 * i] - w0, h[104 + 8 * i] - w0, h[105 + 8 * i] - w0);
        }
        fprintf(stderr, "\n");
    }
    return 0;
}
