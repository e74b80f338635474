// Host-side dump of the pair kernel's unit schedule (tc3_build) for the reference architecture: no GPU needed (the
// weight upload at the end of tc3_build fails without a device, the schedule is complete by then).
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -Ibeetroots_b200/csrc -Iinclude -o /tmp/sched_dump scripts/sched_dump.cu
//   /tmp/sched_dump <n_out> <skip>        # skip = leading blocks left to the pre-pass kernel (0 or 4)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include "common.cuh"
#include "mlp_simt.cuh"
#include "sampler_kernels.cuh"
#include "mlp_tc.cuh"
#include "mlp_tc2.cuh"
#include "mlp_tc3.cuh"
int main(int argc, char** argv) {
    int n_out = atoi(argv[1]), skip = atoi(argv[2]);
    int bn[9] = {17, 25, 38, 57, 85, 128, 192, 288, 432};
    std::vector<std::vector<double>> W(9), B(9);
    std::vector<const double*> wp(9), bp(9);
    int in = 34;
    for (int l = 0; l < 9; ++l) { W[l].assign((size_t)bn[l] * in, 0.01); B[l].assign(bn[l], 0.0); wp[l] = W[l].data(); bp[l] = B[l].data(); in += bn[l]; }
    std::vector<double> wo((size_t)n_out * in, 0.01);
    Tc3Net t;
    int rc = tc3_build(t, 34, 9, bn, wp.data(), bp.data(), wo.data(), n_out, 2.302585, skip);
    Tc3Sched& s = t.sched;
    printf("rc=%d n_units=%d n_ready=%d act_kb=%d ring_bytes=%d ring_ext=%d stage_kb=%d pre_kb=%d pre_feats=%d smem=%zu\n", rc, s.n_units, s.n_ready, s.act_kb, s.ring_bytes, s.ring_ext, s.stage_kb, s.pre_kb, s.pre_feats, t.smem_bytes);
    for (int i = 0; i < s.n_ready; ++i) printf(" ready_count[%d]=%d", i, s.ready_count[i]);
    printf("\n");
    for (int u = 0; u < s.n_units; ++u) {
        Tc3Unit& U = s.U[u];
        printf("U%-2d k16 [%3d,%3d) new %3d dep %d sp %d flags %2d ready_idx %2d two %d ring_lo %6d cw %d rep %d groups %2d/%2d gelu0 %3d/%3d rows %3d/%3d feat0 %4d/%4d\n",
               u, U.k16_begin, U.k16_end, U.new_k16, U.dep, U.sp, U.flags, U.ready_idx, U.two, U.ring_lo, U.cw, U.rep, U.groups[0], U.groups[1],
               U.gelu_lane0[0], U.gelu_lane0[1], U.gelu_rows[0], U.gelu_rows[1], U.feat0[0], U.feat0[1]);
    }
    return 0;
}
