// On-device Markov chain store and posterior summaries.
//
// The reference keeps every saved iterate in host memory, writes it to HDF5 (`list_Theta`,
// sampler/saver/my_saver.py:46-83) and much later reads the file back pixel by pixel to form, per (pixel, parameter),
// the mean and np.percentile of the chain (inversion/results/utils/mc.py:158-215).  At c4 an iterate is 16 MiB, so
// here the thinned chain simply stays in HBM ([T][N][D] float32, 180 GB hold ~10^4 iterates of a 1M-pixel map) and
// the summaries are computed where the data is: one thread per (pixel, parameter) series, the series staged in a
// column of shared memory (or of a global scratch tile when T is too long for 227 KB), heap-sorted in place, and the
// bracketing order statistics of every requested percentile written out.  Order statistics are exact (the sort moves
// float32 values, no arithmetic); the mean is accumulated in float64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BT_CHAIN_MAX_Q 16

struct ChainRanks {
    int n;
    int lo[BT_CHAIN_MAX_Q], hi[BT_CHAIN_MAX_Q];   // 0-based ranks bracketing each percentile (numpy 'linear' method)
};

// element i of the calling thread's series lives at s[i * stride]
__device__ __forceinline__ void chain_sift_down(float* s, size_t stride, int start, int end) {
    const float val = s[(size_t)start * stride];
    int i = start;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= end) break;
        float cv = s[(size_t)c * stride];
        if (c + 1 < end) {
            const float c2 = s[(size_t)(c + 1) * stride];
            if (c2 > cv) { cv = c2; ++c; }
        }
        if (!(cv > val)) break;
        s[(size_t)i * stride] = cv;
        i = c;
    }
    s[(size_t)i * stride] = val;
}

// chain: [T_total][nd] float32; series index = blockIdx.x * blockDim.x + threadIdx.x + series0.
// SMEM: the block's tile [T][blockDim.x] lives in dynamic shared memory; otherwise in scratch[T][scratch_stride]
// (column = series - series0).
template <bool SMEM>
__global__ void chain_stats_kernel(const float* __restrict__ chain, size_t nd, int64_t first, int T, ChainRanks rk,
                                   size_t series0, size_t n_series, float* __restrict__ scratch, size_t scratch_stride,
                                   float* __restrict__ mean, float* __restrict__ var, float* __restrict__ q_lo,
                                   float* __restrict__ q_hi) {
    extern __shared__ float chain_tile[];
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n_series) return;
    const size_t sidx = series0 + col;
    float* s = SMEM ? chain_tile + threadIdx.x : scratch + col;
    const size_t stride = SMEM ? (size_t)blockDim.x : scratch_stride;
    // pass 1: stage the series (coalesced across the warp: consecutive threads read consecutive series of one iterate)
    double sum = 0.0;
    const float* src = chain + (size_t)first * nd + sidx;
    for (int t = 0; t < T; ++t) {
        const float x = src[(size_t)t * nd];
        s[(size_t)t * stride] = x;
        sum += (double)x;
    }
    const double mu = sum / (double)T;
    if (var) {
        double ss = 0.0;
        for (int t = 0; t < T; ++t) { const double dlt = (double)s[(size_t)t * stride] - mu; ss += dlt * dlt; }
        var[sidx] = (float)(ss / (double)T);            // population variance (np.var default)
    }
    if (mean) mean[sidx] = (float)mu;
    if (rk.n == 0) return;
    // pass 2: heapsort of the column (no recursion, no extra storage; every access of a warp hits 32 distinct banks)
    for (int start = T / 2 - 1; start >= 0; --start) chain_sift_down(s, stride, start, T);
    for (int end = T - 1; end > 0; --end) {
        const float top = s[0];
        s[0] = s[(size_t)end * stride];
        s[(size_t)end * stride] = top;
        chain_sift_down(s, stride, 0, end);
    }
    for (int q = 0; q < rk.n; ++q) {
        q_lo[(size_t)q * nd + sidx] = s[(size_t)rk.lo[q] * stride];
        q_hi[(size_t)q * nd + sidx] = s[(size_t)rk.hi[q] * stride];
    }
}
