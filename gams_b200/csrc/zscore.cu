// zscore.cu — Kernel B: GC-wave peak signalling, bit-exact with gams::thresholding_algo
// (src/libs/stat.rs:16-56; mean stat.rs:1-6, stddev stat.rs:8-14).
//
// The reference's statistics are strictly sequential f32 sums (separate multiply and add, sample
// stddev, IEEE sqrt/div), and in low-variance stretches the outcome is decided by rounding noise,
// so every window replays exactly that order: 4*lag non-fusable flops per window.  This kernel is
// therefore FP32-issue bound, not HBM bound (SURVEY §7 "hard parts"); HBM traffic is 4 B read +
// 1 B written per window.  This file is compiled with -fmad=false and uses the _rn intrinsics so
// nothing can be contracted into an FMA.
//
// Two paths:
//  * influence == 1 (the default and every BASELINE config): filtered == data exactly
//    (1*x + 0*f), so windows are independent.  One thread owns ZS_R consecutive windows; their
//    lag-windows overlap, so one shared-memory load feeds ZS_R independent add chains.
//  * influence != 1: the filtered series depends on earlier decisions.  One warp per ctg walks the
//    ctg in lag order; its 32 lanes evaluate the next 32 windows speculatively ("no signal in
//    between") and the warp commits up to and including the window after the first signal, which
//    is exactly what the sequential loop would have produced.
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

__host__ __device__ __forceinline__ int zs_swz(int p) { return p + (p >> 5); }  // breaks the stride-ZS_R bank pattern

// ------------------------------------------------------------------------------------------------
// influence == 1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ZS_THREADS)
zscore_indep_kernel(const float *__restrict__ x, const uint64_t *__restrict__ off,
                    const TileDesc *__restrict__ tiles, int lag, float threshold, int8_t *__restrict__ signal) {
    extern __shared__ float sm[];  // zs_swz(ZS_TILE + lag + 1) + 1 floats, then ZS_TILE signal bytes
    const int tid = threadIdx.x;
    const TileDesc td = tiles[blockIdx.x];
    const uint64_t base = off[td.item];
    const int64_t n = (int64_t)(off[td.item + 1] - base);
    const int64_t w0 = td.first;
    const int nw = (int)min((int64_t)ZS_TILE, n - w0);
    const int64_t lo0 = w0 - 1 - lag;  // series index held at smem position 0 (may be negative)
    const int n_el = nw + lag + 1;
    int8_t *sg = reinterpret_cast<int8_t *>(sm + zs_swz(ZS_TILE + lag + 1) + 1);

    for (int p = tid; p < n_el; p += ZS_THREADS) {
        const int64_t e = lo0 + p;
        sm[zs_swz(p)] = (e >= 0) ? x[base + e] : 0.0f;
    }
    __syncthreads();

    const float flag = __ull2float_rn((unsigned long long)lag);  // `data.len() as f32`
    const float flagm1 = __fsub_rn(flag, 1.0f);                   // `len - 1.`
    const int p0 = tid * ZS_R;                                    // smem position of chain 0's first element
    const int64_t i0 = w0 + p0;                                   // first window of this thread
    int sig[ZS_R];
#pragma unroll
    for (int r = 0; r < ZS_R; ++r) sig[r] = 0;

    if (p0 < nw) {
        const bool fast = (i0 >= (int64_t)lag + 1) && (p0 + ZS_R <= nw) && (lag >= ZS_R);
        if (fast) {
            // chain r sums smem positions p0+r .. p0+r+lag-1, in order; the window's own value is at p0+r+lag+1
            float s[ZS_R], m[ZS_R], q[ZS_R];
#pragma unroll
            for (int r = 0; r < ZS_R; ++r) s[r] = 0.0f;
#pragma unroll
            for (int t = 0; t < ZS_R - 1; ++t) {
                const float v = sm[zs_swz(p0 + t)];
#pragma unroll
                for (int r = 0; r < ZS_R; ++r)
                    if (r <= t) s[r] = __fadd_rn(s[r], v);
            }
#pragma unroll 4
            for (int t = ZS_R - 1; t < lag; ++t) {
                const float v = sm[zs_swz(p0 + t)];
#pragma unroll
                for (int r = 0; r < ZS_R; ++r) s[r] = __fadd_rn(s[r], v);
            }
#pragma unroll
            for (int k = 0; k < ZS_R - 1; ++k) {
                const float v = sm[zs_swz(p0 + lag + k)];
#pragma unroll
                for (int r = 0; r < ZS_R; ++r)
                    if (r > k) s[r] = __fadd_rn(s[r], v);
            }
#pragma unroll
            for (int r = 0; r < ZS_R; ++r) {
                m[r] = __fdiv_rn(s[r], flag);
                q[r] = 0.0f;
            }
#pragma unroll
            for (int t = 0; t < ZS_R - 1; ++t) {
                const float v = sm[zs_swz(p0 + t)];
#pragma unroll
                for (int r = 0; r < ZS_R; ++r)
                    if (r <= t) {
                        const float d = __fsub_rn(v, m[r]);
                        q[r] = __fadd_rn(q[r], __fmul_rn(d, d));
                    }
            }
#pragma unroll 4
            for (int t = ZS_R - 1; t < lag; ++t) {
                const float v = sm[zs_swz(p0 + t)];
#pragma unroll
                for (int r = 0; r < ZS_R; ++r) {
                    const float d = __fsub_rn(v, m[r]);
                    q[r] = __fadd_rn(q[r], __fmul_rn(d, d));
                }
            }
#pragma unroll
            for (int k = 0; k < ZS_R - 1; ++k) {
                const float v = sm[zs_swz(p0 + lag + k)];
#pragma unroll
                for (int r = 0; r < ZS_R; ++r)
                    if (r > k) {
                        const float d = __fsub_rn(v, m[r]);
                        q[r] = __fadd_rn(q[r], __fmul_rn(d, d));
                    }
            }
#pragma unroll
            for (int r = 0; r < ZS_R; ++r) {
                const float sd = __fsqrt_rn(__fdiv_rn(q[r], flagm1));
                const float xv = sm[zs_swz(p0 + r + lag + 1)];
                sig[r] = (fabsf(__fsub_rn(xv, m[r])) > __fmul_rn(threshold, sd)) ? ((xv > m[r]) ? 1 : -1) : 0;
            }
        } else {
            // ctg head (i <= lag), tile tail, or tiny lag: one window at a time, same arithmetic
            auto ld = [&](int64_t e) { return sm[zs_swz((int)(e - lo0))]; };
            for (int r = 0; r < ZS_R; ++r) {
                const int64_t i = i0 + r;
                if (i >= n || i < (int64_t)lag) continue;
                const int64_t a = (i - 1 - (int64_t)lag > 0) ? (i - 1 - (int64_t)lag) : 0;  // stat.rs:30-31 vs :51-52
                sig[r] = zscore_decide(ld, a, (uint64_t)lag, flag, flagm1, ld(i), threshold);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < ZS_R; ++r)
        if (p0 + r < ZS_TILE) sg[p0 + r] = (int8_t)sig[r];
    __syncthreads();
    for (int p = tid; p < nw; p += ZS_THREADS) signal[base + w0 + p] = sg[p];
}

// ------------------------------------------------------------------------------------------------
// any influence: one warp per series, speculative 32-window look-ahead, `filt` in global scratch
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
zscore_seq_kernel(const float *__restrict__ x_all, const uint64_t *__restrict__ off, int64_t n_series, uint64_t lag,
                  float threshold, float influence, float *filt_all, int8_t *__restrict__ signal) {
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_series) return;
    const uint64_t base = off[c];
    const int64_t n = (int64_t)(off[c + 1] - base);
    const float *x = x_all + base;
    float *f = filt_all + base;
    int8_t *sg = signal + base;
    const int64_t L = (int64_t)lag;

    for (int64_t j = lane; j < L && j < n; j += 32) {
        f[j] = x[j];
        sg[j] = 0;
    }
    __syncwarp();
    const float flag = __ull2float_rn((unsigned long long)lag);
    const float flagm1 = __fsub_rn(flag, 1.0f);
    const float one_minus = __fsub_rn(1.0f, influence);

    int64_t i = L;
    while (i < n) {
        const int64_t w = i + lane;
        const bool active = w < n;
        int my = 0;
        float xv = 0.0f;
        if (active) {
            xv = x[w];
            const int64_t a = (w - 1 - L > 0) ? (w - 1 - L) : 0;
            // committed values come from f; the not-yet-committed ones [i, w) are speculated as x
            auto ld = [&](int64_t e) { return (e < i) ? f[e] : x[e]; };
            my = zscore_decide(ld, a, lag, flag, flagm1, xv, threshold);
        }
        const unsigned ballot = __ballot_sync(0xFFFFFFFFu, my != 0);
        const int first = ballot ? (__ffs(ballot) - 1) : 32;
        // lanes <= first+1 saw only correctly speculated values (the stats window of w ends at w-2)
        int n_commit = (first >= 31) ? 32 : first + 2;
        if ((int64_t)n_commit > n - i) n_commit = (int)(n - i);
        // filtered value of the first signalling lane
        const float prev_x = __shfl_up_sync(0xFFFFFFFFu, xv, 1);
        float fw = xv;
        if (lane == first) {
            const float prev = (lane == 0) ? f[i - 1] : prev_x;  // predecessor did not signal => filtered == x
            fw = __fadd_rn(__fmul_rn(influence, xv), __fmul_rn(one_minus, prev));
        }
        const float f_first = __shfl_sync(0xFFFFFFFFu, fw, first & 31);
        if (lane == first + 1 && my != 0) fw = __fadd_rn(__fmul_rn(influence, xv), __fmul_rn(one_minus, f_first));
        if (lane < n_commit) {
            f[w] = fw;
            sg[w] = (int8_t)my;
        }
        __syncwarp();
        i += n_commit;
    }
}

int32_t launch_zscore(const SeriesLayout &L, cudaStream_t st, const float *d_x, uint64_t lag, float threshold,
                      float influence, float *d_filtered, int8_t *d_signal, uint64_t *launches) {
    if (lag == 0) {
        set_error("thresholding: lag must be >= 1 (the reference underflows lag-1, stat.rs:30)");
        return GAMS_CUDA_EINVAL;
    }
    if (L.n_series == 0) return GAMS_CUDA_OK;
    if (L.min_len < lag) {
        set_error("thresholding: a series has %llu values but lag is %llu (the reference panics, stat.rs:30)",
                  (unsigned long long)L.min_len, (unsigned long long)lag);
        return GAMS_CUDA_ELAG;
    }
    const bool tiled = influence == 1.0f && lag <= 40000;
    const size_t smem_need = tiled ? (size_t)(zs_swz(ZS_TILE + (int)lag + 1) + 1) * sizeof(float) + ZS_TILE : 0;
    if (tiled) {
        if (smem_need > 48 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(zscore_indep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need));
        zscore_indep_kernel<<<(unsigned)L.n_ztiles, ZS_THREADS, smem_need, st>>>(d_x, L.d_off, L.d_ztiles, (int)lag,
                                                                                threshold, d_signal);
    } else {
        if (!d_filtered) {
            set_error("thresholding: influence != 1 (or lag too large for the tiled path) needs the d_filtered scratch");
            return GAMS_CUDA_EINVAL;
        }
        const unsigned grid = (unsigned)((L.n_series + 3) / 4);
        zscore_seq_kernel<<<grid, 128, 0, st>>>(d_x, L.d_off, L.n_series, lag, threshold, influence, d_filtered,
                                                d_signal);
    }
    if (launches) ++*launches;
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
