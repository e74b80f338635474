// Forward map kernel, tcgen05 cta_group::2, third generation ("pair kernel").
//
// Same GEMM orientation as mlp_tc2.cuh -- a pair of SMs works on 128 columns,
//     D[256 out features, 128 columns] += W[256, k] x H[128, k]        (UMMA M=256, N=128, K=16: 64 cycles, the
//                                                                        tensor peak, measured scripts/umma2_bench.cu)
// -- rebuilt around what the instrumented timelines of the first two kernels showed (profiles/r01_summary.md):
//   * the weight stream is latency bound (bytes in flight / L2 round trip), so the ring is BYTE granular: a weight
//     chunk takes only the 8-row groups that hold live rows, the live rows of a unit are split evenly between the
//     two CTAs, and up to 16 chunks are in flight in ~54 KB instead of three 16 KB stages;
//   * the output layer (n_out <= 32 rows, K = all features) is folded into the last unit of the last hidden block:
//     its rows ride along as extra rows of that unit for the K range they share, stay in TMEM, and a short
//     continuation over the remaining K range finishes them (saves 2/3 of the output layer's MMAs and chunks);
//   * 16 epilogue warps (lane quarter x column quarter) instead of 8: the activation epilogue is latency bound per
//     warp (scripts/epi_bench.cu), so the per-unit drain time halves;
//   * the input features of the next tile are published right after the last accumulator of this tile completes.
// The schedule is a flat list of "units" built on the host (tc3_build); the device code has no per-layer logic.
#pragma once
#include "mlp_tc2.cuh"

#define TC3_NPROD 2                 // weight-producer threads (warp 0 and the two warps after the epilogue warps)
#define TC3_THREADS (128 + 32 * TC3_EW)
#define TC3_EW 16                  // epilogue warps per CTA
#define TC3_NT 16                  // weight chunks in flight (tickets)
#define TC3_MAX_UNITS 24
#define TC3_F_CONT 1               // accumulate onto the previous unit's slots (no wait for "empty", first MMA accumulates)
#define TC3_F_HOLD 2               // the epilogue leaves the slots allocated: a CONT unit follows
#define TC3_F_PUB 8                // after this unit's accumulators complete nothing of this tile reads K-block 0 any more: publish the next tile's inputs
#define TC3_F_FINAL 4              // lanes [0, n_out) of CTA 0 hold the finished output layer
#define TC3_MISC_BYTES 4096
#ifndef TC3_GATE
#define TC3_GATE 0                 // 1: one polling warp + bar.sync gate for the accumulator barriers -- measured 4 % slower than 16 try_waits
#endif
#ifndef TC3_UNIFORM_ISSUER
#define TC3_UNIFORM_ISSUER 1
#endif
#ifndef TC3_ISSUER_SPIN
#define TC3_ISSUER_SPIN 0         // MMA issuers busy-poll their barriers (test_wait): try_wait parks the thread and was seen to wake late
#endif

struct Tc3Unit {
    short k16_begin, k16_end;      // K range of this unit's MMAs, in steps of 16 features
    short new_k16;                 // first step holding features guarded by ready barrier `dep`
    short dep, sp, flags, ready_idx, two;
    int ring_lo;                   // chunks of this unit may live in [ring_lo, ring_bytes) of the extended ring
    short cw;                      // K-blocks per weight chunk (handshake granularity): 1, 2 or 4
    short rep;                     // small units: the rows are replicated in rep lane groups, each handles 32/rep of a warp's columns
    short groups[2];               // 8-row groups of weights streamed per K-block, per CTA
    short gelu_lane0[2], gelu_rows[2];
    int feat0[2];                  // position in the activation buffer of the feature produced by lane gelu_lane0[c]
    int w_off[2];                  // 1 KB units into CTA c's weight stream
};
struct Tc3Sched {
    int stage_kb;                  // first of the 3 K-blocks used as output staging tile by the final epilogue
    int n_units, n_ready, act_kb, n_out, ring_bytes, ring_ext;   // ring_bytes includes ring_ext bytes borrowed from the tail of act
    int ready_count[TC_MAX_LAYERS + 1];
    Tc3Unit U[TC3_MAX_UNITS];
    const uint8_t* wstream[2];
    const float* bias;             // by position in the activation buffer
    long long* dbg;
    long long* trace;              // debug: per-chunk event times of the second tile, [6][128]
    int trace_i0;
    int dbg_flags;                 // debug: 1 = every chunk advances the ring by 16 KB, 2 = one MMA issuer; knock-outs: 4 = no activation math, 8 = no MMAs, 256 = no weight copies
    uint8_t* dump;                 // debug: activation buffers of both CTAs after the last tile (BT_TC3_DUMP=file)
    int pre_kb;                    // > 0: K-blocks [0, pre_kb) of every tile come ready-made from the pre-pass kernel (mlp_pre.cuh)
    int pre_feats;                 // number of features they hold (= inputs of the first block this kernel evaluates)
    int pre_blocks;                // leading blocks evaluated by the pre-pass
    const uint8_t* pre_img;        // [pair-tile][CTA][pre_kb][8 KB]: the exact shared-memory image of those K-blocks
};
struct Tc3Net {
    bool ready = false;
    Tc3Sched sched{};
    void* d_w[2] = {nullptr, nullptr};
    void* d_b = nullptr;
    size_t smem_bytes = 0;
};
static inline void tc3_free(Tc3Net& t) {
    for (int c = 0; c < 2; ++c) if (t.d_w[c]) cudaFree(t.d_w[c]);
    if (t.d_b) cudaFree(t.d_b);
    t = Tc3Net();
}

static inline int tc3_build(Tc3Net& t, int n_feat0, int n_blocks, const int32_t* block_new, const double* const* weights,
                            const double* const* biases, const double* w_out, int n_out, double out_scale, int skip = 0) {
    t.ready = false;
    if (n_blocks + 1 > TC_MAX_LAYERS || n_out > 32 || n_feat0 > 64 || n_blocks < 1) return 0;
    if (skip < 0 || skip > n_blocks - 2) return 0;
    Tc3Sched& s = t.sched;
    memset(&s, 0, sizeof s);
    s.n_out = n_out;
    s.n_ready = n_blocks - skip + 1;
    s.pre_blocks = skip;
    s.ready_count[0] = 2;                        // one arrival per CTA (the epilogue warps meet at a named barrier first)
    const int og = (n_out + 7) / 8;
    // what every streamed row is: layer (-1 = output layer), row of that layer, K window [klo, khi)
    struct RowSrc { int layer, row, klo, khi; };
    std::vector<std::vector<RowSrc>> rows_of[2];
    int w = n_feat0, prev_in = 0, nu = 0, n_feat = n_feat0;
    for (int l = 0; l < n_blocks; ++l) n_feat += block_new[l];
    bool folded = false;
    int layer_in_of_last_minus_1 = n_feat0;
    int unit_layer[TC3_MAX_UNITS], unit_mt[TC3_MAX_UNITS];
    for (int u = 0; u < TC3_MAX_UNITS; ++u) { unit_layer[u] = -1; unit_mt[u] = 0; }
    for (int l = 0; l < n_blocks; ++l) {
        const int in = w, nnew = block_new[l];
        if (l == n_blocks - 2) layer_in_of_last_minus_1 = in;
        if (l < skip) { prev_in = in; w += nnew; continue; }       // evaluated by the pre-pass kernel
        if (l == skip && skip > 0) { s.pre_feats = in; s.pre_kb = (in + TC_BK - 1) / TC_BK; }
        const int lrel = l - skip;
        const int n_mt = (nnew + 255) / 256;
        for (int mt = 0; mt < n_mt; ++mt) {
            if (nu >= TC3_MAX_UNITS - 1) return 0;
            const int R = std::min(256, nnew - 256 * mt), G = (R + 7) / 8;
            const bool last = l == n_blocks - 1 && mt == n_mt - 1;
            static const bool nofold = getenv("BT_TC3_NOFOLD") != nullptr;
            const bool fold = !nofold && last && (G + og + 1) / 2 <= 16 && G - ((G + og + 1) / 2 - og) <= 16 && (G + og + 1) / 2 >= og;
            const int o = fold ? og : 0;
            // the very first unit goes to CTA 1 alone: CTA 0's lane-quarter-0 warps are still draining the previous
            // tile's outputs when its accumulator completes
            const bool to_peer = nu == 0 && G <= 16 && !fold;
            const int t0 = to_peer ? 0 : (G + o + 1) / 2, gg0 = t0 - o;
            Tc3Unit& U = s.U[nu];
            unit_layer[nu] = l;
            unit_mt[nu] = mt;
            U.k16_begin = 0;
            U.k16_end = (short)((in + 15) / 16);
            U.new_k16 = (short)(lrel == 0 ? 0 : prev_in / 16);
            U.dep = (short)lrel;
            U.sp = (short)(nu & 1);
            U.flags = fold ? TC3_F_HOLD : 0;
            U.ready_idx = (short)(lrel + 1);
            U.two = (short)(((U.k16_end + 3) / 4) >= 2);
            U.gelu_lane0[0] = (short)(o * 8);
            U.gelu_lane0[1] = 0;
            U.gelu_rows[0] = (short)std::min(R, gg0 * 8);
            U.gelu_rows[1] = (short)(R - U.gelu_rows[0]);
            // The activation epilogue is latency bound per warp and a warp can only read its own TMEM lane quarter, so a
            // unit with few rows would keep one scheduler busy and three idle: replicate its rows in 2 or 4 lane groups
            // (identical MMA cost, M is 256 anyway) and let every group convert a share of the columns.
            static const bool norep = getenv("BT_TC3_NOREP") != nullptr;
            const int rmax = std::max<int>(U.gelu_rows[0], U.gelu_rows[1]);
            // (only for units with a short K loop: replication multiplies the weight bytes, and a unit with many K-blocks is
            // paced by its weight stream, not by its epilogue)
            const int nkb_u = (U.k16_end + 3) / 4;
            U.rep = (short)((fold || norep || nkb_u > 4) ? 1 : (rmax <= 32 ? 4 : (rmax <= 64 ? 2 : 1)));
            const int span = 128 / U.rep;
            U.groups[0] = (short)(U.rep == 1 ? t0 : (U.gelu_rows[0] ? (span * (U.rep - 1) + U.gelu_rows[0] + 7) / 8 : 0));
            U.groups[1] = (short)(U.rep == 1 ? (U.gelu_rows[1] + 7) / 8 : (U.gelu_rows[1] ? (span * (U.rep - 1) + U.gelu_rows[1] + 7) / 8 : 0));
            if (U.groups[0] > 16 || U.groups[1] > 16) return 0;
            U.feat0[0] = in + 256 * mt;
            U.feat0[1] = in + 256 * mt + U.gelu_rows[0];
            s.ready_count[lrel + 1] += (U.gelu_rows[0] > 0) + (U.gelu_rows[1] > 0);   // one arrival per CTA that produces features
            for (int c = 0; c < 2; ++c) {
                std::vector<RowSrc> rs((size_t)U.groups[c] * 8, RowSrc{-2, 0, 0, 0});
                if (c == 0 && fold) for (int r = 0; r < n_out; ++r) rs[r] = RowSrc{-1, r, 0, in};
                for (int q = 0; q < U.rep; ++q)
                    for (int r = 0; r < U.gelu_rows[c]; ++r)
                        rs[q * (128 / U.rep) + U.gelu_lane0[c] + r] = RowSrc{l, 256 * mt + (c ? U.gelu_rows[0] : 0) + r, 0, in};
                rows_of[c].push_back(rs);
            }
            folded = fold;
            ++nu;
        }
        prev_in = in;
        w += nnew;
    }
    const int prev_in_saved = prev_in;
    int first_ext_unit = nu;   // set below: first unit issued at least one ring of chunk bytes after the tile's first chunk
    {   // output layer: continuation of the folded unit, or a unit of its own
        const int in_last = prev_in;                 // features available before the last hidden block's outputs
        Tc3Unit& U = s.U[nu];
        U.k16_begin = (short)(folded ? in_last / 16 : 0);
        U.k16_end = (short)((n_feat + 15) / 16);
        U.new_k16 = (short)(in_last / 16);
        U.dep = (short)(n_blocks - skip);
        U.sp = (short)(folded ? s.U[nu - 1].sp : (nu & 1));
        U.flags = (short)(TC3_F_FINAL | (folded ? TC3_F_CONT : 0));
        U.ready_idx = -1;
        U.two = 1;
        U.rep = 1;
        U.groups[0] = (short)og;
        U.groups[1] = 0;
        for (int c = 0; c < 2; ++c) {
            std::vector<RowSrc> rs((size_t)U.groups[c] * 8, RowSrc{-2, 0, 0, 0});
            if (c == 0) for (int r = 0; r < n_out; ++r) rs[r] = RowSrc{-1, r, folded ? in_last : 0, n_feat};
            rows_of[c].push_back(rs);
        }
        ++nu;
    }
    {
        int last0 = 0;
        const int k16_in = 4 * (s.pre_kb > 0 ? s.pre_kb : 1);
        for (int u = 0; u < nu; ++u) if (s.U[u].k16_begin < k16_in) last0 = u;   // reads the K-block(s) holding the tile's inputs
        s.U[last0].flags |= TC3_F_PUB;
    }
    s.n_units = nu;
    // Handshake granularity: every weight chunk costs the single-thread pipeline roles (producer, relay, helper, issuer) a fixed
    // few hundred cycles whatever its size (the bare synchronisation skeleton is ~70 % of the kernel time), so units with many
    // K-blocks stream two K-blocks per chunk.  Both CTAs then stream the same number of 8-row groups (zero rows pad the
    // smaller half) so that the second K-block sits at the same offset in both.
    {
        static const int cw_env = getenv("BT_TC3_CW") ? atoi(getenv("BT_TC3_CW")) : 4;
        for (int u = 0; u < nu; ++u) {
            Tc3Unit& U = s.U[u];
            const int nkb = (U.k16_end + 3) / 4 - U.k16_begin / 4;
            const int gmax = std::max<int>(U.groups[0], U.groups[1]);
            U.cw = (short)((cw_env >= 2 && nkb >= 4) ? 2 : 1);
            if (cw_env >= 4 && nkb >= 8 && gmax * 4 <= 32) U.cw = 4;       // <= 32 KB per chunk, still >= 2 chunks (both issuers)
            if (!(U.flags & TC3_F_CONT) && !(U.flags & TC3_F_FINAL)) U.two = (short)((nkb + U.cw - 1) / U.cw >= 2);
            if (U.cw >= 2) {
                const short g = std::max(U.groups[0], U.groups[1]);
                for (int c = 0; c < 2; ++c) {
                    U.groups[c] = g;
                    rows_of[c][u].resize((size_t)g * 8, RowSrc{-2, 0, 0, 0});
                }
            }
        }
    }
    s.act_kb = (n_feat + TC_BK - 1) / TC_BK;
    const long ring = 227L * 1024 - 1024 - (long)s.act_kb * TC_KB_BYTES - TC3_MISC_BYTES;
    if (ring < 32 * 1024) return 0;
    const int ring_own = (int)(ring / 1024) * 1024;
    t.smem_bytes = (size_t)ring_own + (size_t)s.act_kb * TC_KB_BYTES + TC3_MISC_BYTES + 1024;
    // Extended ring.  The K-blocks that will hold the outputs of the last hidden block are dead until that block's
    // epilogue writes them, which is exactly while the big weight chunks stream: shared memory is laid out
    // [act | ring] and the ring window of a unit starts inside the tail of act.
    //   units before the last hidden block  : nothing reads or writes K-blocks >= kb_first (first K-block holding only
    //                                         outputs of the last block) until the first unit of that block has completed;
    //   unit j of the last hidden block      : its epilogue (after all its chunks are consumed) writes the K-blocks of its
    //                                         own outputs; earlier units of the block have written theirs -> window starts
    //                                         after the outputs of units < j;
    //   output continuation / own unit, and the small units of the next tile (issued while this tile's last MMAs still
    //   read those K-blocks): own ring only.  Units of the next tile that use the extension are issued at least
    //   ring_bytes of chunks after the next tile's first chunk, i.e. after this tile's last MMA has completed (checked).
    {
        static const bool noext = getenv("BT_TC3_NOEXT") != nullptr;
        const int in_last = prev_in_saved;
        const int kb_first = (in_last + TC_BK - 1) / TC_BK;          // K-blocks [kb_first, act_kb) hold only last-block outputs
        int ext = noext ? 0 : (s.act_kb - kb_first) * TC_KB_BYTES;
        // Output staging tile of the final epilogue (<= 17 KB = 3 K-blocks): it must not be ring space (the producer may
        // already stream next-tile chunks into the extension), and nothing may read it before it is rewritten in the next
        // tile.  K-blocks that hold only outputs of the last two hidden blocks qualify: they are rewritten by those blocks'
        // epilogues before any unit reads them again.  Otherwise: staging in the tail of act and no ring extension.
        {
            const int in_prev = n_blocks >= 2 ? layer_in_of_last_minus_1 : n_feat0;
            const int kb_lo = (in_prev + TC_BK - 1) / TC_BK;
            if (kb_first - 3 >= kb_lo && kb_first - 3 >= 1) s.stage_kb = kb_first - 3;
            else { s.stage_kb = s.act_kb - 3; ext = 0; }
            if (s.stage_kb < 1) return 0;
        }
        s.ring_ext = ext;
        s.ring_bytes = ring_own + ext;
        long bytes_before = 0;
        bool ok = true;
        {
            long acc_b = 0;
            for (int u = 0; u < nu; ++u) {
                if (acc_b >= ring_own + ext) { first_ext_unit = u; break; }
                acc_b += 1024L * std::max<int>(s.U[u].groups[0], s.U[u].groups[1]) * ((s.U[u].k16_end + 3) / 4 - s.U[u].k16_begin / 4);
            }
        }
        for (int u = 0; u < nu; ++u) {
            Tc3Unit& U = s.U[u];
            const int nkb = (U.k16_end + 3) / 4 - U.k16_begin / 4;
            const long adv = 1024L * std::max<int>(U.groups[0], U.groups[1]);
            int lo = ext;                                            // own ring only
            if (unit_layer[u] >= 0 && unit_layer[u] < n_blocks - 1 && u >= first_ext_unit) lo = 0;
            else if (unit_layer[u] == n_blocks - 1) {
                // outputs of earlier units of the last block end at position in_last + 256 * mt
                const int written_to = in_last + 256 * unit_mt[u];   // exclusive
                const int kb_safe = unit_mt[u] == 0 ? kb_first : (written_to + TC_BK - 1) / TC_BK;
                lo = std::min(ext, std::max(0, (kb_safe - kb_first) * TC_KB_BYTES));
            }
            if (lo < ext && unit_layer[u] < n_blocks - 1 && bytes_before < s.ring_bytes) ok = false;
            U.ring_lo = lo;
            bytes_before += adv * nkb;
        }
        if (!ok) { for (int u = 0; u < nu; ++u) s.U[u].ring_lo = ext; }
    }
    // ---- streams
    std::vector<float> bs((size_t)n_feat + 8, 0.f);
    {
        int pos = n_feat0;
        for (int l = 0; l < n_blocks; ++l) {
            for (int o = 0; o < block_new[l]; ++o) bs[pos + o] = (float)biases[l][o];
            pos += block_new[l];
        }
    }
    std::vector<int> layer_in(n_blocks + 1);
    {
        int pos = n_feat0;
        for (int l = 0; l < n_blocks; ++l) { layer_in[l] = pos; pos += block_new[l]; }
        layer_in[n_blocks] = n_feat;
    }
    for (int c = 0; c < 2; ++c) {
        size_t total_kb = 0;
        for (int u = 0; u < nu; ++u) {
            s.U[u].w_off[c] = (int)total_kb;
            total_kb += (size_t)s.U[u].groups[c] * (size_t)((s.U[u].k16_end + 3) / 4 - s.U[u].k16_begin / 4);
        }
        std::vector<uint16_t> ws(total_kb * 512 + 8, 0);
        for (int u = 0; u < nu; ++u) {
            const Tc3Unit& U = s.U[u];
            const std::vector<RowSrc>& rs = rows_of[c][u];
            size_t pos_kb = (size_t)U.w_off[c];
            for (int kb = U.k16_begin / 4; kb < (U.k16_end + 3) / 4; ++kb, pos_kb += U.groups[c]) {
                uint16_t* dst = ws.data() + pos_kb * 512;
                for (size_t r = 0; r < rs.size(); ++r) {
                    if (rs[r].layer == -2) continue;
                    const bool is_out = rs[r].layer == -1;
                    const int ld = is_out ? n_feat : layer_in[rs[r].layer];
                    const double* W = is_out ? w_out : weights[rs[r].layer];
                    const double sc = is_out ? out_scale : 1.0;
                    for (int kk = 0; kk < TC_BK; ++kk) {
                        const int k = kb * TC_BK + kk;
                        if (k < rs[r].klo || k >= rs[r].khi) continue;
                        // activations at k >= n_feat0 are stored as 2 x GELU (saves a multiply per element): exact in bf16
                        const double half = k >= n_feat0 ? 0.5 : 1.0;
                        dst[sw128_offset((uint32_t)r, (uint32_t)kk) / 2] = f32_to_bf16_rne((float)(half * sc * W[(size_t)rs[r].row * ld + k]));
                    }
                }
            }
        }
        if (cudaMalloc(&t.d_w[c], ws.size() * 2) != cudaSuccess) return -1;
        if (cudaMemcpy(t.d_w[c], ws.data(), ws.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
        s.wstream[c] = (const uint8_t*)t.d_w[c];
    }
    if (cudaMalloc(&t.d_b, bs.size() * 4) != cudaSuccess) return -1;
    if (cudaMemcpy(t.d_b, bs.data(), bs.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    s.bias = (const float*)t.d_b;
    t.ready = true;
    return 0;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 2 x GELU and 2 x GELU' (the factor 1/2 lives in the weights); one clamp: p(min(x^2, 64)) > 0 keeps the sign of x
__device__ __forceinline__ float gelu2_tc(float x) {
    const float x2 = fminf(x * x, 64.f);
    const float t = tanh_mufu(x * fmaf(x2, fmaf(x2, TC_G2, TC_G1), TC_G0));
    return fmaf(x, t, x);
}
__device__ __forceinline__ float gelu2_prime_tc(float x) {
    const float xc = fminf(fmaxf(x, -8.f), 8.f), x2 = xc * xc;
    const float t = tanh_mufu(xc * fmaf(x2, fmaf(x2, TC_G2, TC_G1), TC_G0));
    const float du = fmaf(x2, fmaf(x2, 5.f * TC_G2, 3.f * TC_G1), TC_G0);
    return (1.f + t) + xc * (1.f - t * t) * du;
}
// standardised monomials of one column: features fg, fg + 8, ... (8 threads per column)
#define TC3_FPT 8
__device__ __forceinline__ void tc3_poly_feats(const int* pexp_s, const float* pmean_s, const float* pistd_s, int F0, int fg,
                                               int s, int tin, bool valid, const float (&xf)[BT_MAX_D + 1], float (&feat)[TC3_FPT]) {
#pragma unroll
    for (int qf = 0; qf < TC3_FPT; ++qf) {
        const int f = fg + 8 * qf;
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

__device__ __forceinline__ void umma2_f16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait_f(uint32_t bar, uint32_t parity, bool spin) {
    if (spin) mbar_spin(bar, parity); else mbar_wait(bar, parity);
}
// PRE: the first sch.pre_kb K-blocks of every tile are copied from sch.pre_img (written by the pre-pass kernel) instead of
// being computed here; the schedule then starts at block sch.pre_blocks
template <int S, int DBG, bool PRE = false>   // DBG: 0 = production, 1 = runtime debug flags / timelines, > 1 = these knock-out flags compiled in
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC3_THREADS, 1)
mlp_tc3_kernel(Tc3Sched sch, NetDev net, FmapDev fm, const float* __restrict__ rows, int M, float* __restrict__ lf,
               float* __restrict__ jac) {
    // debug knobs / timeline buffers compile away in the production instantiation
    const int dbg_flags = DBG == 1 ? sch.dbg_flags : DBG;
    long long* const dbg_buf = DBG == 1 ? sch.dbg : nullptr;
    long long* const trace_buf = DBG == 1 ? sch.trace : nullptr;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* act = smem;                                             // [act_kb][64 columns x 64 features bf16, SW128]
    uint8_t* ring = act + (size_t)sch.act_kb * TC_KB_BYTES - sch.ring_ext;   // extended ring: tail of act + own ring (see tc3_build)
    uint8_t* misc = ring + sch.ring_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(misc + 1000);
    float* x0s2 = reinterpret_cast<float*>(misc + 1024);             // [2][128] theta_full[0] per column (leader's copy is read)
    float* pmean_s = x0s2 + 256;
    float* pistd_s = pmean_s + 64;
    int* pexp_s = reinterpret_cast<int*>(pistd_s + 64);              // [64][4]
    uint32_t* cst_all = reinterpret_cast<uint32_t*>(pexp_s + 256);   // [NPROD][NT] virtual start of the chunks in flight, one table per producer
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + TC3_NT), bar_pfull = smem_u32(bars + 2 * TC3_NT);
    const uint32_t bar_tfull = smem_u32(bars + 3 * TC3_NT), bar_tempty = smem_u32(bars + 3 * TC3_NT + 4);
    const uint32_t bar_lready = smem_u32(bars + 3 * TC3_NT + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    constexpr int P2 = TC2_BN / S;                                   // rows of theta per pair-tile
    const int n_ptiles = (M + P2 - 1) / P2;
    const int pt0 = blockIdx.x >> 1, pt_step = gridDim.x >> 1;
    const uint32_t ring_bytes = (uint32_t)sch.ring_bytes;

    if (threadIdx.x == 0) {
        // leader: a chunk is "full" when its own rows have landed (producer's expect_tx arrival + bytes) AND the peer's relay has arrived
        for (int i = 0; i < 3 * TC3_NT; ++i) mbar_init(bar_full + 8 * i, (i < TC3_NT && crank == 0) ? 2 : 1);
        for (int i = 0; i < 4; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 2); }
        for (int l = 0; l < sch.n_ready; ++l) mbar_init(bar_lready + 8 * l, sch.ready_count[l]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        uint4* a4 = reinterpret_cast<uint4*>(smem);
        const int n16 = (sch.ring_bytes - sch.ring_ext + sch.act_kb * TC_KB_BYTES) / 16;
        for (int i = threadIdx.x; i < n16; i += TC3_THREADS) a4[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    for (int i = threadIdx.x; i < net.F0; i += TC3_THREADS) {
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

    const bool second_producer = TC3_NPROD == 2 && ((crank == 0 && warp == 3) || (crank == 1 && warp == 1));
    if (warp == 0 || second_producer) {
        // ===================== weight producers (both CTAs, own rows of every unit) =====================
        // Issuing a bulk copy costs the thread ~80 cycles, but the copies of ONE warp complete one after the other (~450
        // cycles each whatever their size, scripts/tma_bench.cu) while copies of different warps overlap: two producer
        // warps take the chunks alternately.  The second one is a warp that has nothing else to do: warp 3 in the leader
        // (the MMA issuers poll the "full" barriers themselves), warp 1 in the peer (MMAs are issued by the leader only).
        // Every producer walks the whole schedule (positions are deterministic) and retires chunks on its own.
        const uint32_t pid = warp == 0 ? 0u : 1u;
        uint32_t* cst = cst_all + pid * TC3_NT;
        if (lane == 0) {
            const uint32_t ring_addr = smem_u32(ring);
            // The ring is a byte FIFO in a virtual address space that never wraps back: chunk i lives at virtual [v, v + adv),
            // physical v mod ring (a chunk that would straddle the end skips to offset 0, the skipped bytes count as used).
            // Everything older than v + adv - ring_bytes is overwritten by chunk i, so those chunks must have been consumed.
            uint32_t i = 0, tail = 0, pos = 0, vpos = 0;
            for (int pt = pt0; pt < n_ptiles; pt += pt_step)
                for (int u = 0; u < sch.n_units; ++u) {
                    const Tc3Unit U = sch.U[u];
                    const uint32_t adv_kb = (dbg_flags & 1) ? 16384u : (uint32_t)(U.groups[0] > U.groups[1] ? U.groups[0] : U.groups[1]) * 1024u;
                    const uint32_t bytes_kb = (uint32_t)U.groups[crank] * 1024u;
                    const uint8_t* src = sch.wstream[crank] + (size_t)U.w_off[crank] * 1024;
                    const int kb1 = (U.k16_end + 3) >> 2;
                    // operand windows: K-block q of a chunk is read as 128 rows (16 KB) from pos + q * adv_kb; the last one may run
                    // up to TC3_MISC_BYTES past the ring (into misc: garbage rows feed accumulator lanes nobody reads)
                    const uint32_t win = (uint32_t)(U.cw - 1) * adv_kb + 16384u - TC3_MISC_BYTES;
                    for (int kb = U.k16_begin >> 2; kb < kb1; kb += U.cw, ++i) {
                        const uint32_t nk = (uint32_t)(kb1 - kb < U.cw ? kb1 - kb : U.cw);
                        const uint32_t adv = adv_kb * nk, bytes = bytes_kb * nk;
                        // the 128-row operand windows (16 KB per K-block) must end inside the ring; chunks of this unit start at >= ring_lo
                        if (pos + win > ring_bytes || pos + adv > ring_bytes) { vpos += ring_bytes - pos; pos = 0; }   // data and operand windows fit
                        if (pos < (uint32_t)U.ring_lo) { vpos += (uint32_t)U.ring_lo - pos; pos = (uint32_t)U.ring_lo; }
                        const uint32_t t = i % TC3_NT;
                        if (i % TC3_NPROD != pid) {                          // somebody else's chunk: just remember where it lives
                            if (tail + TC3_NT <= i) { mbar_spin(bar_empty + 8 * (tail % TC3_NT), (tail / TC3_NT) & 1u); ++tail; }
                            cst[t] = vpos;
                            pos += adv;
                            vpos += adv;
                            src += bytes;
                            continue;
                        }
                        const uint32_t lo = vpos + adv - ring_bytes;         // oldest virtual byte that survives this chunk
                        while (tail < i && (tail + TC3_NT <= i || (int32_t)(lo - cst[tail % TC3_NT]) > 0)) {
                            mbar_spin(bar_empty + 8 * (tail % TC3_NT), (tail / TC3_NT) & 1u);
                            if (trace_buf && blockIdx.x == 0 && tail - (uint32_t)sch.trace_i0 < 128u) trace_buf[640 + tail - sch.trace_i0] = clock64();
                            ++tail;
                        }
                        cst[t] = vpos;
                        if (trace_buf && blockIdx.x == 0 && i - (uint32_t)sch.trace_i0 < 128u) trace_buf[i - sch.trace_i0] = clock64();
                        if (bytes && !(dbg_flags & 256)) {
                            mbar_expect_tx(bar_full + 8 * t, bytes);
                            bulk_g2s(ring_addr + pos, src, bytes, bar_full + 8 * t);
                        } else {
                            mbar_arrive(bar_full + 8 * t);                   // nothing to fetch for this CTA
                        }
                        pos += adv;
                        vpos += adv;
                        src += bytes;
                    }
                }
        }
    } else if (warp == 3) {
        // ===================== relay (peer CTA): "my chunk has landed" -> leader =====================
        if (lane == 0 && crank == 1) {
            const uint32_t leader_pfull = mapa_u32(bar_full, 0);     // second arrival of the leader's "full" barrier
            uint32_t i = 0;
            for (int pt = pt0; pt < n_ptiles; pt += pt_step)
                for (int u = 0; u < sch.n_units; ++u) {
                    const int n = (((sch.U[u].k16_end + 3) >> 2) - (sch.U[u].k16_begin >> 2) + sch.U[u].cw - 1) / sch.U[u].cw;
                    for (int k = 0; k < n; ++k, ++i) {
                        const uint32_t t = i % TC3_NT;
                        mbar_spin(bar_full + 8 * t, (i / TC3_NT) & 1u);
                        mbar_arrive_cluster(leader_pfull + 8 * t);
                    }
                }
        }
    } else if (warp == 1 || warp == 2) {
        // ===================== MMA issuers (leader CTA only), K-blocks split by parity over two threads =====================
        const int par = warp - 1;
        // The whole warp runs the (warp-uniform) issue loop and ONE elected lane executes the tcgen05 instructions: with a
        // single active lane the compiler cannot prove uniformity and wraps every MMA in vector->uniform register moves.
        if ((TC3_UNIFORM_ISSUER || lane == 0) && crank == 0) {
            const bool leader_lane = TC3_UNIFORM_ISSUER ? elect_one() : true;
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC2_BN >> 3) << 17) | ((uint32_t)(TC2_BM >> 4) << 24);
            const uint32_t act_addr = smem_u32(act), ring_addr = smem_u32(ring);
            const uint32_t desc_lo0 = 1u << 16, desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // umma_desc_sw128(0), split in words
            uint32_t i = 0, pos = 0, nacq0 = 0, nacq1 = 0, tcount = 0;
            for (int pt = pt0; pt < n_ptiles; pt += pt_step, ++tcount) {
                const bool rec = dbg_buf && blockIdx.x == 0 && tcount == 1 && par == ((dbg_flags >> 12) & 1) && lane == 0;   // flag 4096: trace the second issuer
                if (rec) dbg_buf[0] = clock64();
                for (int u = 0; u < sch.n_units; ++u) {
                    const Tc3Unit U = sch.U[u];
                    const uint32_t adv = (dbg_flags & 1) ? 16384u : (uint32_t)(U.groups[0] > U.groups[1] ? U.groups[0] : U.groups[1]) * 1024u;   // per K-block
                    const uint32_t slot = 2u * (uint32_t)U.sp + (uint32_t)par;
                    if (rec) dbg_buf[1 + 2 * u] = clock64();
                    uint32_t acc = 1u;
                    if (!(U.flags & TC3_F_CONT)) {
                        const uint32_t k = U.sp ? nacq1++ : nacq0++;
                        mbar_wait_f(bar_tempty + 8 * slot, (k & 1u) ^ 1u, TC3_ISSUER_SPIN || (dbg_flags & 32));
                        tc_fence_after();
                        acc = 0u;
                    }
                    const uint32_t d_tmem = tmem_base + slot * TC2_BN;
                    bool waited = false;
                    const int kb0 = U.k16_begin >> 2, kb1 = (U.k16_end + 3) >> 2, cw = U.cw;
                    const uint32_t ring_lo = (uint32_t)U.ring_lo, win = (uint32_t)(cw - 1) * adv + 16384u - TC3_MISC_BYTES;
                    const int k16b = U.k16_begin, k16e = U.k16_end, new_k16 = U.new_k16;
                    int ci = 0;
                    for (int kbc = kb0; kbc < kb1; kbc += cw, ++i, ++ci) {
                        const int nk = kb1 - kbc < cw ? kb1 - kbc : cw;
                        if (pos + win > ring_bytes || pos + adv * (uint32_t)nk > ring_bytes) pos = 0;
                        if (pos < ring_lo) pos = ring_lo;
                        const uint32_t cur = pos;
                        pos += adv * (uint32_t)nk;
                        if ((dbg_flags & 2) ? par != 0 : (ci & 1) != par) continue;
                        const uint32_t t = i % TC3_NT;
                        bool polled = false;
                        for (int q = 0; q < nk; ++q) {
                            const int kb = kbc + q;
                            const int ks_lo = kb == kb0 ? (k16b & 3) : 0;
                            const int ks_hi = kb == kb1 - 1 ? k16e - 4 * kb : 4;
                            if (!waited && 4 * kb + ks_hi > new_k16) {
                                mbar_wait_f(bar_lready + 8 * U.dep, tcount & 1u, TC3_ISSUER_SPIN || (dbg_flags & 32));
                                if (rec) dbg_buf[2 + 2 * u] = clock64();
                                waited = true;
                            }
                            if (!polled) {
                                mbar_spin(bar_full + 8 * t, (i / TC3_NT) & 1u);     // weights of chunk i landed in both CTAs (2 arrivals)
                                polled = true;
                                const bool tr = trace_buf && lane == 0 && blockIdx.x == 0 && i - (uint32_t)sch.trace_i0 < 128u;
                                if (tr) trace_buf[384 + i - sch.trace_i0] = clock64();
                            }
                            // descriptors differ only in the 14-bit start-address field of the low word
                            const uint32_t a_lo = desc_lo0 + ((ring_addr + cur + (uint32_t)q * adv) >> 4);
                            const uint32_t b_lo = desc_lo0 + ((act_addr + kb * TC_KB_BYTES) >> 4);
                            if (!(dbg_flags & 8) && leader_lane) {
                                if (ks_lo == 0 && ks_hi == 4) {
                                    umma2_f16_lo(d_tmem, a_lo, b_lo, desc_hi, idesc, acc);
                                    umma2_f16_lo(d_tmem, a_lo + 2, b_lo + 2, desc_hi, idesc, 1u);
                                    umma2_f16_lo(d_tmem, a_lo + 4, b_lo + 4, desc_hi, idesc, 1u);
                                    umma2_f16_lo(d_tmem, a_lo + 6, b_lo + 6, desc_hi, idesc, 1u);
                                    acc = 1u;
                                } else {
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks)
                                        if (ks >= ks_lo && ks < ks_hi) { umma2_f16_lo(d_tmem, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, idesc, acc); acc = 1u; }
                                }
                            }
                        }
                        if (leader_lane) umma2_commit_mc(bar_empty + 8 * t, 3);   // frees the chunk in both CTAs
                        if (TC3_UNIFORM_ISSUER) __syncwarp();
                        if (trace_buf && lane == 0 && blockIdx.x == 0 && i - (uint32_t)sch.trace_i0 < 128u) trace_buf[512 + i - sch.trace_i0] = clock64();
                    }
                    if (leader_lane) umma2_commit_mc(bar_tfull + 8 * slot, 3);   // accumulator half complete, both CTAs
                    if (TC3_UNIFORM_ISSUER) __syncwarp();
                }
                if (rec) dbg_buf[63] = clock64();
            }
        }
    } else if (warp >= 4 && warp < 4 + TC3_EW) {
        // ===================== prologue + epilogue (16 warps per CTA) =====================
        const int ew = warp - 4, qd = ew & 3, cq = ew >> 2;          // TMEM lane quarter, column quarter of the pair-tile
        const uint32_t owner = (uint32_t)(cq >> 1);                  // CTA whose activation buffer holds my 32 columns
        const uint32_t act_dst = mapa_u32(smem_u32(act), owner) + (uint32_t)((cq & 1) * 32 * 128);
        const uint32_t l_tempty = mapa_u32(bar_tempty, 0), l_lready = mapa_u32(bar_lready, 0);
        const uint32_t x0_leader = mapa_u32(smem_u32(x0s2), 0);
        const uint32_t act_s = smem_u32(act);
        // prologue work distribution: thread (pcol, fg) computes features fg, fg + 8, ... of my CTA's column pcol
        const int t512 = ew * 32 + lane, pcol = t512 & 63, fg = t512 >> 6;
        const int gcol_p = (int)crank * 64 + pcol, pp = gcol_p / S, ps = gcol_p - pp * S;
        const int ptin = (S > 1 && ps > 0) ? fm.tinput[ps - 1] : -1;
        const int lane_g = qd * 32 + lane;
        uint32_t nfull0 = 0, nfull1 = 0, tcount_e = 0;
        float feat[TC3_FPT], x0_cur;
        uint4 pv[3];                                                  // PRE: my 3 x 16 B of the next tile's input K-blocks
        const int pre_n16 = PRE ? sch.pre_kb * (TC_KB_BYTES / 16) : 0;  // 16-byte words per CTA per tile (<= 1536)
        auto load_pre = [&](int pt_) {
            const uint4* src = reinterpret_cast<const uint4*>(sch.pre_img) + ((size_t)pt_ * 2 + crank) * (size_t)pre_n16;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int idx = t512 + 512 * q;
                pv[q] = idx < pre_n16 ? __ldg(src + idx) : make_uint4(0, 0, 0, 0);
            }
        };
        uint32_t pub_parity = 0;
        auto publish_inputs = [&]() {
            const uint32_t parity = pub_parity;
            if (PRE) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int idx = t512 + 512 * q;
                    if (idx < pre_n16)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(act_s + (uint32_t)idx * 16u), "r"(pv[q].x), "r"(pv[q].y), "r"(pv[q].z), "r"(pv[q].w) : "memory");
                }
            } else {
#pragma unroll
            for (int qf = 0; qf < TC3_FPT; ++qf) {
                const int f = fg + 8 * qf;
                if (f < net.F0) sts_bf16(act_s + sw128_offset((uint32_t)pcol, (uint32_t)f), feat[qf]);
            }
            }
            if (fg == 0)
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(x0_leader + (parity * 128u + (uint32_t)gcol_p) * 4u), "f"(x0_cur) : "memory");
            fence_proxy_async();
        };
        if (pt0 < n_ptiles) {
            float xf0[BT_MAX_D + 1];
            tc_load_column(fm, rows, M, pt0 * P2 + pp, xf0);
            if (PRE) load_pre(pt0);
            else tc3_poly_feats(pexp_s, pmean_s, pistd_s, net.F0, fg, ps, ptin, pt0 * P2 + pp < M, xf0, feat);
            x0_cur = xf0[0];
            publish_inputs();
            asm volatile("bar.sync 2, 512;" ::: "memory");
            if (ew == 0 && lane == 0) mbar_arrive_cluster(l_lready);
        }
        for (int pt = pt0; pt < n_ptiles; pt += pt_step, ++tcount_e) {
            const bool rec = dbg_buf && blockIdx.x == 0 && tcount_e == 1 && ew == ((dbg_flags >> 4) & 15) && lane == 0;
            const bool has_next = pt + pt_step < n_ptiles;
            float xf_n[BT_MAX_D + 1];
            const int row_n = (pt + pt_step) * P2 + pp;
            tc_load_column(fm, rows, M, row_n, xf_n);                // latency hidden behind this tile's units

            for (int u = 0; u < sch.n_units; ++u) {
                const Tc3Unit U = sch.U[u];
                const uint32_t slot = 2u * (uint32_t)U.sp;
                const uint32_t par = (U.sp ? nfull1++ : nfull0++) & 1u;
                const int g0 = U.gelu_lane0[crank], gr = U.gelu_rows[crank];
                const int span_m1 = 128 / U.rep - 1;                 // lanes per replica - 1
                const int lane_r = lane_g & span_m1, q0 = (qd * 32) & span_m1;
                const bool gelu_live = lane_r >= g0 && lane_r < g0 + gr;
                const bool warp_gelu = gr > 0 && q0 < g0 + gr && q0 + 32 > g0;
                const bool is_final = (U.flags & TC3_F_FINAL) != 0;
                const bool warp_out = is_final && crank == 0 && qd == 0;
                const bool release = !(U.flags & TC3_F_HOLD);
                const uint32_t kabs = (uint32_t)(U.feat0[crank] + lane_r - g0);
                const float b = gelu_live ? __ldg(sch.bias + kabs) : 0.f;
                const bool pub = (U.flags & TC3_F_PUB) && has_next;
                if (pub) {                                           // features of the next tile, while this unit's MMAs run
                    if (PRE) load_pre(pt + pt_step);
                    else tc3_poly_feats(pexp_s, pmean_s, pistd_s, net.F0, fg, ps, ptin, row_n < M, xf_n, feat);
                    x0_cur = xf_n[0];
                }
#if TC3_GATE
                // ONE warp busy-polls the accumulator barriers and releases the other 15 through a hardware barrier: a warp parked
                // in mbarrier.try_wait was measured to resume ~550 cycles after the phase completes, once per unit on the critical path
                if (ew == 1) { mbar_spin(bar_tfull + 8 * slot, par); mbar_spin(bar_tfull + 8 * (slot + 1), par); }
                asm volatile("bar.sync 3, 512;" ::: "memory");
#else
                mbar_wait_f(bar_tfull + 8 * slot, par, dbg_flags & 64);
                mbar_wait_f(bar_tfull + 8 * (slot + 1), par, dbg_flags & 64);
#endif
                if (rec) dbg_buf[64 + 2 * u] = clock64();
                if (pub) { pub_parity = (tcount_e + 1u) & 1u; publish_inputs(); }   // every MMA of this tile that reads K-block 0 has completed
                tc_fence_after();
                if (gr == 0 && !(is_final && crank == 0) && !pub) {
                    // this CTA holds no row of the unit: no TMEM reads, no stores, no CTA barrier (its warps may still be
                    // busy with the previous tile's outputs); one thread hands the accumulator slots back
                    if (release && ew == 1 && lane == 0) { mbar_arrive_cluster(l_tempty + 8 * slot); mbar_arrive_cluster(l_tempty + 8 * (slot + 1)); }
                    continue;
                }
                // ---- phase 1: accumulators -> registers
                const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + slot * TC2_BN + (uint32_t)(cq * 32);
                const int ng = 4 / U.rep, gbeg = ((qd * U.rep) >> 2) * ng;   // replicated units: my groups of 8 columns
                float v[32];
                if (warp_gelu && U.rep > 1) {
#pragma unroll
                    for (int gg = 0; gg < 2; ++gg)
                        if (gg < ng) {
                            float t8[8];
                            tmem_ld8(taddr + (uint32_t)((gbeg + gg) * 8), t8);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[gg * 8 + j] = t8[j];
                            if (U.two) {
                                tmem_ld8(taddr + TC2_BN + (uint32_t)((gbeg + gg) * 8), t8);
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[gg * 8 + j] += t8[j];
                            }
                        }
                } else if (warp_gelu || warp_out) {
                    tmem_ld32(taddr, v);
                    if (U.two && !(dbg_flags & 2)) {
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            float w[16];
                            tmem_ld16(taddr + TC2_BN + hh * 16, w);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[hh * 16 + j] += w[j];
                        }
                    }
                }
                tc_fence_before();
                // One remote arrival per CTA instead of one per warp: 32 arrivals per barrier per unit (three barriers) serialise
                // on the leader's mbarriers and made the bare synchronisation skeleton cost ~6 k cycles per unit.
                // ONE CTA barrier per unit (after the stores): accumulator slots, next-tile inputs and this unit's features are
                // all handed over together by one thread.  The slots of unit u are next needed by unit u + 2, long after.
                if (is_final && pub) {                               // (only when the output layer is not folded)
                    asm volatile("bar.sync 2, 512;" ::: "memory");
                    if (ew == 1 && lane == 0) mbar_arrive_cluster(l_lready);
                }
                if (is_final && crank != 0 && ew == 1 && lane == 0) {   // the peer reads nothing of the final accumulators
                    mbar_arrive_cluster(l_tempty + 8 * slot);
                    mbar_arrive_cluster(l_tempty + 8 * (slot + 1));
                }
                // ---- phase 2: activation, stores
                if (warp_gelu) {
                    const uint32_t kk = kabs % TC_BK, chunk = kk >> 3;
                    const uint32_t dst = act_dst + (kabs / TC_BK) * TC_KB_BYTES + ((kk & 7u) << 1);
                    if (U.rep > 1) {
                        if (gelu_live && !(dbg_flags & 4)) {
#pragma unroll
                            for (int gg = 0; gg < 2; ++gg)
                                if (gg < ng) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        float val;
                                        if (S == 1) {
                                            val = gelu2_tc(v[gg * 8 + j] + b);
                                        } else {
                                            const int jp = gg * 8 + j - (j % S);
                                            const float pre = v[jp] + b;
                                            val = (j % S == 0) ? gelu2_tc(pre) : gelu2_prime_tc(pre) * v[gg * 8 + j];
                                        }
                                        st_cluster_bf16(dst + (uint32_t)((gbeg + gg) * 1024 + j * 128) + ((chunk ^ (uint32_t)j) << 4), val);
                                    }
                                }
                        }
                    } else {
                        uint32_t swz[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) swz[r] = dst + (uint32_t)(r * 128) + ((chunk ^ (uint32_t)r) << 4);
                        if (gelu_live && !(dbg_flags & 4)) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float val;
                                if (S == 1) {
                                    val = gelu2_tc(v[j] + b);
                                } else {
                                    const int jp = j - (j % S);
                                    const float pre = v[jp] + b;
                                    val = (j % S == 0) ? gelu2_tc(pre) : gelu2_prime_tc(pre) * v[j];
                                }
                                st_cluster_bf16(swz[j & 7] + (uint32_t)((j >> 3) * 1024), val);
                            }
                        }
                    }
                    if (!(dbg_flags & 1024)) fence_proxy_async();   // CTA-scoped (see mlp_tc2.cuh for why not the cluster form)
                }
                if (warp_out) {
                    // stage [column][output] fp32 in the last two K-blocks of the activation buffer (every MMA of this tile has
                    // completed; the next tile's last hidden block rewrites them before they are read again), then write the tile's
                    // outputs with coalesced stores: lf rows and jac rows of a tile are contiguous in global memory
                    const uint32_t stage_s = act_s + (uint32_t)sch.stage_kb * TC_KB_BYTES;   // [n_out][132] floats <= 17 KB, see tc3_build
                    const int no = sch.n_out;
                    if (lane_g < no) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)                 // row pitch 132 floats: conflict-free
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(stage_s + (uint32_t)((lane_g * 132 + cq * 32 + j) * 4)), "f"(v[j]) : "memory");
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    // the four output warps are the only readers of the final accumulators in this CTA: slots free again
                    if (cq == 0 && lane == 0) { mbar_arrive_cluster(l_tempty + 8 * slot); mbar_arrive_cluster(l_tempty + 8 * (slot + 1)); }
                    // thread <-> column of the pair-tile; the n_out values of a column are contiguous in lf / jac
                    const int col = cq * 32 + lane;
                    const int p = col / S, sc = col - p * S, row = pt * P2 + p;
                    if (row < M) {
                        float* dst = sc == 0 ? lf + (size_t)row * no : jac + ((size_t)row * (S - 1) + (sc - 1)) * no;
                        const float add = sc == 0 ? x0s2[(tcount_e & 1u) * 128u + (uint32_t)col] : 0.f;
                        for (int f0 = 0; f0 < no; f0 += 8) {         // loads first, then stores (no aliasing stalls)
                            float o8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                o8[i] = 0.f;
                                if (f0 + i < no) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o8[i]) : "r"(stage_s + (uint32_t)(((f0 + i) * 132 + col) * 4)));
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) if (f0 + i < no) dst[f0 + i] = o8[i] + add;
                        }
                    }
                    // no second barrier: the staging tile is next written a whole tile later
                }
                if (!is_final) {
                    asm volatile("bar.sync 2, 512;" ::: "memory");  // every warp has drained the accumulators; its stores are done and fenced
                    if (ew == 1 && lane == 0) {
                        if (release) { mbar_arrive_cluster(l_tempty + 8 * slot); mbar_arrive_cluster(l_tempty + 8 * (slot + 1)); }
                        if (pub) mbar_arrive_cluster(l_lready);      // inputs of the next tile
                        if (U.ready_idx >= 0) mbar_arrive_cluster(l_lready + 8 * U.ready_idx);
                    }
                }
                if (rec) dbg_buf[65 + 2 * u] = clock64();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (sch.dump && blockIdx.x < 2) {
        const int n16 = sch.act_kb * TC_KB_BYTES / 16;
        uint4* dst = reinterpret_cast<uint4*>(sch.dump) + (size_t)crank * n16;
        for (int i = threadIdx.x; i < n16; i += TC3_THREADS) dst[i] = reinterpret_cast<const uint4*>(act)[i];
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

static inline int tc3_launch(Tc3Net& t, const FmapDev& fm, const NetDev& net, const float* rows, int M, float* lf, float* jac,
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
        cudaMalloc(&dbg, 384 * sizeof(long long));
        cudaMemset(dbg, 0, 384 * sizeof(long long));
    }
    t.sched.dbg = dbg;
    const char* dump_path = getenv("BT_TC3_DUMP");
    const size_t dump_bytes = 2 * (size_t)t.sched.act_kb * TC_KB_BYTES;
    uint8_t* dump = nullptr;
    if (dump_path && n_ptiles == 1) cudaMalloc(&dump, dump_bytes);
    t.sched.dump = dump;
    long long* trace = nullptr;
    int n_chunks = 0;
    for (int u = 0; u < t.sched.n_units; ++u) n_chunks += ((t.sched.U[u].k16_end + 3) / 4 - t.sched.U[u].k16_begin / 4 + t.sched.U[u].cw - 1) / t.sched.U[u].cw;
    if (dbg && getenv("BT_TC3_TRACE")) { cudaMalloc(&trace, 768 * sizeof(long long)); cudaMemset(trace, 0, 768 * sizeof(long long)); }
    t.sched.trace = trace;
    t.sched.trace_i0 = n_chunks;
    t.sched.dbg_flags = getenv("BT_TC3_FLAGS") ? atoi(getenv("BT_TC3_FLAGS")) : 0;
    cudaError_t err;
    const bool use_dbg = dbg || trace || dump || t.sched.dbg_flags;
#define TC3_LAUNCH(S_, D_)                                                                                                     \
    do {                                                                                                                       \
        err = cudaFuncSetAttribute(mlp_tc3_kernel<S_, D_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_bytes);  \
        if (err == cudaSuccess) mlp_tc3_kernel<S_, D_><<<g, TC3_THREADS, t.smem_bytes, stream>>>(t.sched, net, fm, rows, M, lf, jac); \
    } while (0)
    static const int ko = getenv("BT_TC3_KO") ? atoi(getenv("BT_TC3_KO")) : 0;   // compile-time knock-outs (S = 1 only)
#ifdef BT_KNOCKOUTS     // timing experiments with invalid results: not in the production library (scripts/run_ko.sh builds them)
    if (S == 1 && ko == 4) TC3_LAUNCH(1, 4);
    else if (S == 1 && ko == 8) TC3_LAUNCH(1, 8);
    else if (S == 1 && ko == 256) TC3_LAUNCH(1, 256);
    else if (S == 1 && ko == 268) TC3_LAUNCH(1, 268);
    else if (S == 1 && ko == 260) TC3_LAUNCH(1, 260);
    else if (S == 1 && ko == 1024) TC3_LAUNCH(1, 1024);   // no generic->async proxy fence after the activation stores (results invalid)
    else
#else
    if (ko) return -8;
#endif
    if (t.sched.pre_kb > 0) {
        if (t.sched.pre_kb > 3 || !t.sched.pre_img) return -7;
        err = S == 1 ? cudaFuncSetAttribute(mlp_tc3_kernel<1, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_bytes)
                     : cudaFuncSetAttribute(mlp_tc3_kernel<4, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_bytes);
        if (err == cudaSuccess) {
            if (S == 1) mlp_tc3_kernel<1, 0, true><<<g, TC3_THREADS, t.smem_bytes, stream>>>(t.sched, net, fm, rows, M, lf, jac);
            else        mlp_tc3_kernel<4, 0, true><<<g, TC3_THREADS, t.smem_bytes, stream>>>(t.sched, net, fm, rows, M, lf, jac);
        }
    }
    else if (S == 1) { if (use_dbg) TC3_LAUNCH(1, 1); else TC3_LAUNCH(1, 0); }
    else             { if (use_dbg) TC3_LAUNCH(4, 1); else TC3_LAUNCH(4, 0); }
#undef TC3_LAUNCH
    if (err != cudaSuccess) return -5;
    if (cudaPeekAtLastError() != cudaSuccess) return -6;
    if (dump) {
        std::vector<uint8_t> h(dump_bytes);
        cudaStreamSynchronize(stream);
        cudaMemcpy(h.data(), dump, dump_bytes, cudaMemcpyDeviceToHost);
        cudaFree(dump);
        t.sched.dump = nullptr;
        if (FILE* f = fopen(dump_path, "wb")) { fwrite(h.data(), 1, dump_bytes, f); fclose(f); }
    }
    if (dbg) {
        long long h[384];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        t.sched.dbg = nullptr;
        const long long t0 = h[0];
        if (trace) {
            std::vector<long long> tr(768);
            cudaMemcpy(tr.data(), trace, 768 * sizeof(long long), cudaMemcpyDeviceToHost);
            cudaFree(trace);
            t.sched.trace = nullptr;
            fprintf(stderr, "[tc3 chunk trace, cycles since tile start] chunk: issue(producer) full mmas_retired(flag 2048) poll_ok mma_issued retired_seen(producer)\n");
            int c = 0;
            for (int u = 0; u < t.sched.n_units; ++u)
                for (int kb = t.sched.U[u].k16_begin / 4; kb < (t.sched.U[u].k16_end + 3) / 4; kb += t.sched.U[u].cw, ++c) {
                    if (c >= 128) break;
                    auto rel = [&](int k) { return tr[k * 128 + c] ? tr[k * 128 + c] - t0 : -1; };
                    fprintf(stderr, "  U%d kb%-2d: %7lld %7lld %7lld %7lld %7lld %7lld\n", u, kb, rel(0), rel(1), rel(2), rel(3), rel(4), rel(5));
                }
        }
        fprintf(stderr, "[tc3 timeline S=%d M=%d] MMA thread: ", S, M);
        for (int u = 0; u < t.sched.n_units; ++u) fprintf(stderr, "U%d start %lld ready %lld | ", u, h[1 + 2 * u] - t0, h[2 + 2 * u] ? h[2 + 2 * u] - t0 : -1);
        fprintf(stderr, "end %lld\n[tc3 timeline] epilogue warp4: ", h[63] - t0);
        for (int u = 0; u < t.sched.n_units; ++u) fprintf(stderr, "U%d full %lld done %lld | ", u, h[64 + 2 * u] - t0, h[65 + 2 * u] ? h[65 + 2 * u] - t0 : -1);
        fprintf(stderr, "\n");
    }
    return 0;
}
