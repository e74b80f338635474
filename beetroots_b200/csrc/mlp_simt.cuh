// Forward map kernel, CUDA-core float32 variant (the parity path).
//
// Evaluates, for a batch of rows theta (M, D_sampling):
//     lf'(row, l) = theta_full[0] + LOGE_10 * (W_out h(row))_l          (shifted log-intensity)
// and optionally the forward-mode Jacobian d lf / d x_i for the T sampled NN inputs, where h is
// the dense-concat feature vector of PolynomialNetwork -> DenselyConnected
// (replaces NeuralNetworkApprox.compute_all / evaluate_log,
//  modelling/forward_maps/neural_network_approx.py:107-122,152-305, and nnbma's forward +
//  torch.func.vmap(jacfwd), :83).
//
// Layout: one CTA owns a tile of 32 "columns" (a column = the primal of a row or one of its
// tangents); the whole concat H[F_total][32] lives in shared memory (1296 x 32 x 4 B = 166 KB),
// so activations never touch HBM.  Lane = column; each warp produces 8 output features per pass
// with the transposed weights streamed through L1/L2 as warp-uniform 128-bit loads.
#pragma once
#include "common.cuh"

#define BT_MAX_BLOCKS 16
#define MLP_COLS 32
#define MLP_OT 8
#define MLP_THREADS 512

struct NetDev {
    int n_in, order, F0, n_blocks, F_total, n_out;
    const int* exps;          // [F0][order]
    const float* mean;        // [F0]
    const float* inv_std;     // [F0]
    const float* Wt[BT_MAX_BLOCKS];   // [in_l][ld_l]  transposed, zero padded to ld_l = round_up(new_l, 8)
    const float* bias[BT_MAX_BLOCKS]; // [ld_l]
    int in_[BT_MAX_BLOCKS], new_[BT_MAX_BLOCKS], ld[BT_MAX_BLOCKS];
    const float* Wt_out;      // [F_total][ld_out], scaled by LOGE_10
    int ld_out;
};

struct FmapDev {
    int D;                    // sampled dims
    int d_full;               // n_in + 1
    int T;                    // number of tangents = sampled NN inputs
    int full_idx[BT_MAX_D];   // sampled dim d -> index in theta_full
    int tslot[BT_MAX_D];      // sampled dim d -> tangent slot (or -1 for kappa: d lf / d theta = 1)
    int tinput[BT_MAX_T];     // tangent slot -> NN input index
    float fixed[BT_MAX_D + 1];
};

// rows: (M, D) float32.  lf: (M, L).  jac: (M, T, L) or nullptr (then S = 1).
template <bool WITH_JAC>
__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_simt_kernel(NetDev net, FmapDev fm, const float* __restrict__ rows, int M, float* __restrict__ lf,
                float* __restrict__ jac) {
    extern __shared__ float smem[];
    float* H = smem;                                  // [F_total][32]
    float* xin = smem + (size_t)net.F_total * MLP_COLS;   // [d_full][32] full parameter vector per column
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = MLP_THREADS / 32;
    const int S = WITH_JAC ? 1 + fm.T : 1;
    const int P = MLP_COLS / S;
    const int p = lane / S, s = lane - p * S;
    const int ntiles = (M + P - 1) / P;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row = tile * P + p;
        const bool valid = (p < P) && (row < M);
        __syncthreads();                              // previous tile fully consumed
        // ---- full parameter vector of this column
        if (warp == 0) {
            float xf[BT_MAX_D + 1];
#pragma unroll
            for (int i = 0; i <= BT_MAX_D; ++i) xf[i] = (i < fm.d_full) ? fm.fixed[i] : 0.f;
            if (valid) {
#pragma unroll
                for (int d = 0; d < BT_MAX_D; ++d)
                    if (d < fm.D) {
                        const float v = rows[(size_t)row * fm.D + d];
#pragma unroll
                        for (int i = 0; i <= BT_MAX_D; ++i)
                            if (i == fm.full_idx[d]) xf[i] += v;
                    }
            }
#pragma unroll
            for (int i = 0; i <= BT_MAX_D; ++i)
                if (i < fm.d_full) xin[i * MLP_COLS + lane] = xf[i];
        }
        __syncthreads();
        // ---- polynomial features (+ directional derivative for tangent columns), standardised
        const int tin = (WITH_JAC && s > 0) ? fm.tinput[s - 1] : -1;
        for (int f = warp; f < net.F0; f += nwarps) {
            float val;
            if (s == 0) {
                val = 1.f;
                for (int j = 0; j < net.order; ++j) {
                    const int e = net.exps[f * net.order + j];
                    if (e >= 0) val *= xin[(e + 1) * MLP_COLS + lane];
                }
                val = (val - net.mean[f]) * net.inv_std[f];
            } else {
                val = 0.f;
                for (int j = 0; j < net.order; ++j) {
                    if (net.exps[f * net.order + j] != tin) continue;
                    float term = 1.f;
                    for (int j2 = 0; j2 < net.order; ++j2) {
                        const int e2 = net.exps[f * net.order + j2];
                        if (j2 != j && e2 >= 0) term *= xin[(e2 + 1) * MLP_COLS + lane];
                    }
                    val += term;
                }
                val *= net.inv_std[f];
            }
            H[f * MLP_COLS + lane] = valid ? val : 0.f;
        }
        __syncthreads();
        // ---- dense-concat blocks
        for (int l = 0; l < net.n_blocks; ++l) {
            const int K = net.in_[l], NEW = net.new_[l], LD = net.ld[l];
            const float* __restrict__ Wt = net.Wt[l];
            for (int o0 = warp * MLP_OT; o0 < NEW; o0 += nwarps * MLP_OT) {
                float acc[MLP_OT];
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(net.bias[l] + o0));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(net.bias[l] + o0 + 4));
                // tangent columns carry no bias
                const float bs = (s == 0) ? 1.f : 0.f;
                acc[0] = bs * b0.x; acc[1] = bs * b0.y; acc[2] = bs * b0.z; acc[3] = bs * b0.w;
                acc[4] = bs * b1.x; acc[5] = bs * b1.y; acc[6] = bs * b1.z; acc[7] = bs * b1.w;
                const float* wp = Wt + o0;
#pragma unroll 4
                for (int k = 0; k < K; ++k) {
                    const float h = H[k * MLP_COLS + lane];
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * LD));
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * LD + 4));
                    acc[0] = fmaf(w0.x, h, acc[0]); acc[1] = fmaf(w0.y, h, acc[1]);
                    acc[2] = fmaf(w0.z, h, acc[2]); acc[3] = fmaf(w0.w, h, acc[3]);
                    acc[4] = fmaf(w1.x, h, acc[4]); acc[5] = fmaf(w1.y, h, acc[5]);
                    acc[6] = fmaf(w1.z, h, acc[6]); acc[7] = fmaf(w1.w, h, acc[7]);
                }
#pragma unroll
                for (int i = 0; i < MLP_OT; ++i) {
                    const float pre = acc[i];
                    const float prim = __shfl_sync(0xffffffffu, pre, p * S);   // primal pre-activation of this pixel
                    const float val = (s == 0) ? gelu_f(pre) : gelu_prime_f(prim) * pre;
                    if (o0 + i < NEW) H[(K + o0 + i) * MLP_COLS + lane] = val;
                }
            }
            __syncthreads();
        }
        // ---- output layer (restricted to the fitted lines), natural-log scale
        {
            const int K = net.F_total, NEW = net.n_out, LD = net.ld_out;
            for (int o0 = warp * MLP_OT; o0 < NEW; o0 += nwarps * MLP_OT) {
                float acc[MLP_OT];
#pragma unroll
                for (int i = 0; i < MLP_OT; ++i) acc[i] = 0.f;
                const float* wp = net.Wt_out + o0;
#pragma unroll 4
                for (int k = 0; k < K; ++k) {
                    const float h = H[k * MLP_COLS + lane];
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * LD));
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * LD + 4));
                    acc[0] = fmaf(w0.x, h, acc[0]); acc[1] = fmaf(w0.y, h, acc[1]);
                    acc[2] = fmaf(w0.z, h, acc[2]); acc[3] = fmaf(w0.w, h, acc[3]);
                    acc[4] = fmaf(w1.x, h, acc[4]); acc[5] = fmaf(w1.y, h, acc[5]);
                    acc[6] = fmaf(w1.z, h, acc[6]); acc[7] = fmaf(w1.w, h, acc[7]);
                }
                if (valid) {
                    const float x0 = xin[lane];       // theta_full[0] = scaled ln kappa  (:198-200)
#pragma unroll
                    for (int i = 0; i < MLP_OT; ++i) {
                        const int o = o0 + i;
                        if (o < NEW) {
                            if (s == 0) lf[(size_t)row * NEW + o] = acc[i] + x0;
                            else jac[((size_t)row * fm.T + (s - 1)) * NEW + o] = acc[i];
                        }
                    }
                }
            }
        }
    }
}
