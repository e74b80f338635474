// Forward map kernel, tcgen05 / TMEM bf16 variant (placeholder until the tensor-core kernel lands:
// tc_build reports "not ready", so bt_set_mlp_mode(BT_MLP_BF16_TC) fails loudly).
#pragma once
#include "common.cuh"
#include "mlp_simt.cuh"

struct TcNet {
    bool ready = false;
};
static inline void tc_free(TcNet& t) { t.ready = false; }
static inline int tc_build(TcNet& t, int, int, const int32_t*, const double* const*, const double* const*, const double*, int, double) {
    t.ready = false;
    return 0;
}
static inline int tc_launch(TcNet&, const FmapDev&, const NetDev&, const float*, int, float*, float*, int, cudaStream_t) { return -1; }
