// Host-side check of the schedule and weight streams of the column-major pair kernel (beetroots_b200/csrc/mlp_tc4.cuh):
// replays the step / chunk lists on the CPU exactly as the device roles walk them (three tiles, so that the cross-tile
// hazards are exercised), with the TMEM layout measured by scripts/umma3_bench.cu, and checks
//   * every feature a step reads was written by a chunk the issuer has waited for (or comes from the pre-pass),
//   * every accumulator column a step overwrites was drained (issuer waited for the chunk that lived there),
//   * the outputs equal a direct evaluation of the network with the same bf16-rounded weights.
// Test infrastructure (run by tests/test_host.py); no GPU needed.   nvcc -std=c++17 -o tc4_sched_check tc4_sched_check.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../beetroots_b200/csrc/common.cuh"
#include "../beetroots_b200/csrc/mlp_simt.cuh"
#include "../beetroots_b200/csrc/mlp_tc4.cuh"

static float bf(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static float rbf(float x) { return bf(f32_to_bf16_rne(x)); }
static double gelu2(double x) { return x * (1.0 + erf(x / sqrt(2.0))); }

static int check(int n_layers, double growing, int n_out, int nc_max, int mode) {
    std::mt19937 rng(7);
    std::normal_distribution<double> nd(0, 1);
    const int F0 = 34, n_blocks = n_layers - 1;
    std::vector<int32_t> bn;
    int w = F0;
    for (int l = 0; l < n_blocks; ++l) { int nw = std::max(1, (int)(growing * w)); bn.push_back(nw); w += nw; }
    const int n_feat = w;
    std::vector<std::vector<double>> W(n_blocks), B(n_blocks);
    std::vector<const double*> Wp, Bp;
    w = F0;
    for (int l = 0; l < n_blocks; ++l) {
        W[l].resize((size_t)bn[l] * w); B[l].resize(bn[l]);
        for (auto& x : W[l]) x = nd(rng) * sqrt(1.6 / w);
        for (auto& x : B[l]) x = nd(rng) * 0.1;
        Wp.push_back(W[l].data()); Bp.push_back(B[l].data());
        w += bn[l];
    }
    std::vector<double> Wo((size_t)n_out * n_feat);
    for (auto& x : Wo) x = nd(rng) * 0.35 / sqrt((double)n_feat);
    Tc4Net t;
    const int rc = tc4_build(t, F0, n_blocks, bn.data(), Wp.data(), Bp.data(), Wo.data(), n_out, 2.302585092994046, nc_max, mode, false);
    if (rc != 0 || t.h_steps.empty()) { printf("arch (%d, %.1f, %d) nc_max %d mode %d: not taken (rc %d)\n", n_layers, growing, n_out, nc_max, mode, rc); return 0; }
    const Tc4Sched& s = t.sched;
    printf("arch (%d, %.1f, %d) nc_max %3d mode %d: pre %d blocks, %d steps in %d segments, %d chunks, %d slots x %d B, act %d K-blocks, smem %zu, MMA model %.0f cycles/tile\n",
           n_layers, growing, n_out, nc_max, mode, t.n_pre, s.n_steps, s.n_segs, s.n_chunks, s.n_slots, s.slot_bytes, s.act_kb, t.smem_bytes, t.mma_cycles);
    int bad = 0;
    {   // the segments the device roles walk must expand to exactly the step list checked below
        size_t i = 0;
        for (const Tc4Seg& g : t.h_segs) {
            uint32_t w_off = g.w_off;
            for (int kb = g.kb0; kb < g.kb1; ++kb, ++i, w_off += g.rows) {
                if (i >= t.h_steps.size()) { ++bad; break; }
                const Tc4Step& st = t.h_steps[i];
                const bool first = kb == g.kb0, last = kb == g.kb1 - 1;
                const int lo = first ? g.ks_lo : 0, hi = last ? g.ks_hi : 4;
                bool ok = st.w_off == w_off && st.rows == g.rows && st.tcol == g.tcol && st.kb == kb && st.ks_lo == lo && st.ks_hi == hi;
                ok = ok && ((st.flags & TC4_F_OFF) == (g.flags & TC4_F_OFF));
                ok = ok && (((st.flags & TC4_F_ACC0) != 0) == (first && (g.flags & TC4_F_ACC0)));
                ok = ok && (((st.flags & TC4_F_WAIT_IN) != 0) == (first && (g.flags & TC4_F_WAIT_IN)));
                ok = ok && (((st.flags & TC4_F_KBFREE) != 0) == (last && (g.flags & TC4_F_KBFREE)));
                ok = ok && st.wait[0] == (first ? g.wait[0] : -1) && st.wait[1] == (first ? g.wait[1] : -1) && st.wait_prev == (first ? g.wait_prev : -1);
                ok = ok && st.commit == (last ? g.commit : -1);
                if (!ok) { if (bad++ < 5) printf("  segment expansion differs from step %zu\n", i); }
            }
        }
        if (i != t.h_steps.size()) { printf("  segments expand to %zu steps, expected %zu\n", i, t.h_steps.size()); ++bad; }
    }
    // ---- direct evaluation with the rounded weights (activations kept in double)
    for (int trial = 0; trial < 2; ++trial) {
        std::vector<double> h(n_feat);
        for (int f = 0; f < F0; ++f) h[f] = rbf((float)nd(rng));
        {
            int in = F0;
            for (int l = 0; l < n_blocks; ++l) {
                for (int o = 0; o < bn[l]; ++o) {
                    const float b = (float)B[l][o];
                    const float bh = rbf(b), bl = rbf(b - bh);
                    double a = (double)bh + (double)bl;
                    if (l < t.n_pre) a = b;      // the pre-pass keeps its biases in float32
                    for (int k = 0; k < in; ++k) a += (double)rbf((float)((k >= F0 ? 0.5 : 1.0) * W[l][(size_t)o * in + k])) * h[k];
                    h[in + o] = gelu2(a);        // stored as 2 x GELU
                }
                in += bn[l];
            }
        }
        std::vector<double> ref(n_out);
        for (int o = 0; o < n_out; ++o) {
            double a = 0;
            for (int k = 0; k < n_feat; ++k) a += (double)rbf((float)((k >= F0 ? 0.5 : 1.0) * 2.302585092994046 * Wo[(size_t)o * n_feat + k])) * h[k];
            ref[o] = a;
        }
        // ---- replay
        const int NP = s.act_kb * 64;
        std::vector<double> act(NP, 0.0);
        std::vector<int> writer(NP, -1), writer_tile(NP, -1);
        {
            int in = F0;
            for (int l = 0; l < t.n_pre; ++l) in += bn[l];
            for (int k = 0; k < in; ++k) act[k] = h[k];
            act[s.one_pos] = act[s.one_pos + 1] = 1.0;
        }
        std::vector<double> tm[2];
        tm[0].assign(512, 0.0); tm[1].assign(512, 0.0);
        std::vector<int> own(512, -1), own_tile(512, -1);
        std::vector<int> waited(s.n_chunks, -1000);
        std::vector<double> out(32, 0.0), tail(32, 0.0);
        const int n_my = 3;
        for (int it = -1; it < n_my; ++it) {
            for (int si = 0; si < s.n_steps; ++si) {
                const Tc4Step& st = t.h_steps[si];
                const int tl = it + ((st.flags & TC4_F_OFF) ? 1 : 0);
                if (tl < 0 || tl >= n_my) continue;
                for (int k = 0; k < 2; ++k) if (st.wait[k] >= 0) waited[st.wait[k]] = std::max(waited[st.wait[k]], tl);
                if (st.wait_prev >= 0 && tl > 0) waited[st.wait_prev] = std::max(waited[st.wait_prev], tl - 1);
                const int n2 = st.rows;
                if (st.flags & TC4_F_ACC0)
                    for (int q = st.tcol; q < st.tcol + n2; ++q) {
                        if (own[q] >= 0 && waited[own[q]] < own_tile[q]) { if (bad++ < 5) printf("  TMEM hazard: step %d overwrites column %d of chunk %d (tile %d) not waited for\n", si, q, own[q], own_tile[q]); }
                        tm[0][q] = tm[1][q] = 0.0;
                    }
                for (int ks = st.ks_lo; ks < st.ks_hi; ++ks)
                    for (int k16 = 0; k16 < 16; ++k16) {
                        const int kk = ks * 16 + k16, p = st.kb * 64 + kk;
                        // every activation this tile reads: the current tile's value (the replay keeps one buffer per tile life)
                        bool used = false;
                        for (int c = 0; c < 2; ++c)
                            for (int j = 0; j < n2; ++j) {
                                const uint16_t wv = t.h_w[c][(size_t)st.w_off * 64 + sw128_offset((uint32_t)j, (uint32_t)kk) / 2];
                                if (wv & 0x7fff) { used = true; tm[c][st.tcol + j] += (double)bf(wv) * act[p]; }
                            }
                        if (used && writer[p] >= 0 && (writer_tile[p] != tl || waited[writer[p]] < tl)) {
                            if (bad++ < 5) printf("  feature hazard: step %d (tile %d) reads position %d written by chunk %d (tile %d), waited %d\n", si, tl, p, writer[p], writer_tile[p], waited[writer[p]]);
                        }
                    }
                if (st.commit >= 0) {
                    const Tc4Chunk& ch = t.h_chunks[st.commit];
                    if ((int)ch.off != ((st.flags & TC4_F_OFF) ? 1 : 0)) { printf("  chunk/step tile offset mismatch\n"); ++bad; }
                    for (int hh = 0; hh < 2; ++hh)
                        for (int u = 0; u < ch.units[hh]; ++u)
                            for (int j = 0; j < 8; ++j) {
                                const double v = tm[hh][ch.tcol + 8 * u + j];
                                if (ch.kind == 0) { const int p = ch.pos[hh] + 8 * u + j; act[p] = gelu2(v); writer[p] = st.commit; writer_tile[p] = tl; }
                                else { const int f = ch.pos[hh] + 8 * u + j; for (int o = 0; o < 16; ++o) tail[o] += (double)t.h_wtail[(size_t)f * 16 + o] * gelu2(v); }
                            }
                    if (ch.kind == 3)
                        for (int hh = 0; hh < 2; ++hh)
                            for (int j = 0; j < 8; ++j) { const int o = hh * 8 + j; if (o < n_out) out[o] = tm[hh][ch.ocol + j] + tail[o]; tail[o] = 0.0; }
                    for (int q = ch.tcol; q < ch.tcol + (ch.kind == 0 ? 0 : 0); ++q) (void)q;
                    // ownership of the accumulator columns (whole chunk, as the issuer sees it)
                    for (int q = st.tcol; q < st.tcol + n2; ++q) { own[q] = st.commit; own_tile[q] = tl; }
                    if (ch.kind == 3) {
                        double e = 0;
                        for (int o = 0; o < n_out; ++o) e = std::max(e, fabs(out[o] - ref[o]));
                        if (e > 1e-9 * (1 + fabs(ref[0]))) { if (bad++ < 5) printf("  output mismatch tile %d: max abs %.3e (ref[0] %.4f out[0] %.4f)\n", tl, e, ref[0], out[0]); }
                    }
                }
            }
        }
    }
    printf("   -> %s\n", bad ? "FAILED" : "ok");
    return bad;
}

int main() {
    int bad = 0;
    for (int mode = 0; mode < 1; ++mode) {
        for (int nc : {128, 160, 192, 224, 256}) bad += check(10, 0.5, 10, nc, mode);
        bad += check(10, 0.5, 16, 160, mode);
        bad += check(7, 0.5, 5, 160, mode);
        bad += check(6, 1.0, 7, 160, mode);
        bad += check(9, 0.5, 3, 160, mode);
        bad += check(4, 0.5, 10, 160, mode);
    }
    printf(bad ? "TC4 SCHEDULE CHECK FAILED (%d)\n" : "TC4 SCHEDULE CHECK OK\n", bad);
    return bad ? 1 : 0;
}
