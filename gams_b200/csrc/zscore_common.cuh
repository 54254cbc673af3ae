// zscore_common.cuh — the reference's z-score arithmetic for one window (stat.rs:1-14, 36-38), shared by
// zscore.cu and gc_wave_fused.cu.  Translation units including this must be built with -fmad=false.
#pragma once
#include "common.cuh"

namespace gams {

// ---- the reference's arithmetic, one window, elements fetched through `ld(j)` -------------------
template <class Ld>
__device__ __forceinline__ int zscore_decide(Ld ld, int64_t a, uint64_t lag, float flag, float flagm1, float x,
                                             float threshold) {
    float s = 0.0f;
    for (uint64_t t = 0; t < lag; ++t) s = __fadd_rn(s, ld(a + (int64_t)t));
    const float m = __fdiv_rn(s, flag);
    float q = 0.0f;
    for (uint64_t t = 0; t < lag; ++t) {
        const float d = __fsub_rn(ld(a + (int64_t)t), m);
        q = __fadd_rn(q, __fmul_rn(d, d));
    }
    const float sd = __fsqrt_rn(__fdiv_rn(q, flagm1));
    if (fabsf(__fsub_rn(x, m)) > __fmul_rn(threshold, sd)) return (x > m) ? 1 : -1;
    return 0;
}


}  // namespace gams
