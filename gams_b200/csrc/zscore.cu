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
#include "zscore_common.cuh"

namespace gams {

__host__ __device__ __forceinline__ int zs_swz(int p) { return p + (p >> 5); }  // breaks the stride-ZS_R bank pattern

// ------------------------------------------------------------------------------------------------
// influence == 1: exact replay of ZS_R consecutive windows by one thread (blocked add chains)
// ------------------------------------------------------------------------------------------------
// chain r sums smem positions p0+r .. p0+r+lag-1, in order; the window's own value is at p0+r+lag+1
__device__ __forceinline__ void replay_block(const float *sm, int p0, int lag, float flag, float flagm1,
                                             float threshold, int (&sig)[ZS_R]) {
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
}

// MODE 0: certified filter, exact replay only where the certificate is inconclusive (default)
// MODE 1: exact replay of every window (the round-1a kernel)
// MODE 2: verify — replay every window AND count certified decisions that disagree (must be 0)
//
// The filter.  For window i with stats range W (n = lag values) the reference computes, in f32 and in a
// fixed order, m^ = fl(sum)/n, sd^ = sqrt(fl(sum (x-m^)^2)/(n-1)) and signals iff |x_i - m^| > thr*sd^.
// Let mu and S = sum (x-mu)^2 be the exact (real) mean and centred sum of squares of the same f32 inputs.
// Standard forward error bounds for recursive summation (u = 2^-24, M = max|x| over the tile) give
//     |m^ - mu| <= E_m = 1.01 (n+2) u M,
//     sd^ in sqrt(Q/(n-1)) [1-R, 1+R],  Q = S + n (mu - m^)^2 in [S, S + n E_m^2],  R = (0.6 n + 8) u,
//     fl|fl(x_i - m^)| in |x_i - m^| [1-u, 1+u],   fl(thr*sd^) in thr*sd^ [1-u, 1+u]   (u's folded into R).
// mu and S are evaluated in f64 from per-tile prefix sums of y = x - x_ref and y^2 (errors err_mu, err_S,
// bounded with N = tile elements and Ymax = max|y|).  If the worst case of the left side still exceeds the
// worst case of the right side the window certainly signals; if the best case of the left side is below the
// best case of the right side it certainly does not; otherwise (a few windows per million on genomes, and
// every window inside a constant-GC run, where the reference compares rounding noise with rounding noise)
// the window is replayed exactly.  Either way the output equals the reference's.
//
// Mapping: the filter runs one window per thread, strided (lane = consecutive windows: conflict-free f64
// prefix lookups); inconclusive windows go to a per-tile worklist replayed one per thread; if the list
// overflows (degenerate tile) the whole tile is replayed with the blocked chains of replay_block.
constexpr int ZS_LIST_CAP = 256;

// Tile-independent factors of the certificate, evaluated once on the host (no f64 divisions on the device).
struct ZsConst {
    double k_hi, k_lo;  // thr^2 (1 +- R)^2 (1 +- 5u) / (n - 1)
    double c_slack;     // slack   = c_slack * M + tiny           (E_m + err_mu)
    double c_errS;      // err_S   = c_errS * M^2 + tiny
    double c_extra;     // extra   = c_extra * M^2 + tiny          (err_S + n E_m^2)
    double inv_n;       // 1 / n (its rounding is inside err_mu)
};

static ZsConst make_zs_const(uint64_t lag, float threshold) {
    const double u = 5.9604644775390625e-8, u64e = 1.1102230246251565e-16;
    const float flag = (float)lag, flagm1 = flag - 1.0f;  // `data.len() as f32`, `len - 1.`
    const double dn = (double)lag, dnm1 = (double)flagm1;
    const double n_el = (double)(ZS_TILE + lag + 1), N2 = n_el * n_el;
    const double cE = 1.01 * (dn + 2.0) * u;                  // E_m    = cE M
    const double cMu = 4.0 * N2 * u64e * 2.0 / dn;            // err_mu = cMu M        (Ymax <= 2 M)
    const double cS = 16.0 * N2 * u64e * 4.0;                 // err_S  = cS M^2
    const double R = (0.6 * dn + 8.0) * u;
    const double t2 = (double)threshold * (double)threshold;
    ZsConst c;
    c.k_hi = t2 * (1.0 + R) * (1.0 + R) * (1.0 + 5.0 * u) / dnm1;
    c.k_lo = t2 * (1.0 - R) * (1.0 - R) * (1.0 - 5.0 * u) / dnm1;
    c.c_slack = (cE + cMu) * (1.0 + 1e-9);
    c.c_errS = cS;
    c.c_extra = (cS + dn * cE * cE) * (1.0 + 1e-9);
    c.inv_n = 1.0 / dn;
    return c;
}

template <int MODE>
__global__ void __launch_bounds__(ZS_THREADS)
zscore_indep_kernel(const float *__restrict__ x, const uint64_t *__restrict__ off,
                    const TileDesc *__restrict__ tiles, int lag, float threshold, const ZsConst zc,
                    int8_t *__restrict__ signal, unsigned long long *__restrict__ stats, int *__restrict__ nonfinite) {
    extern __shared__ __align__(16) unsigned char zs_smem[];
    const int n_el_max = ZS_TILE + lag + 1;
    double2 *YY = reinterpret_cast<double2 *>(zs_smem);          // n_el_max + 1: exclusive prefixes of (y, y^2)
    float *sm = reinterpret_cast<float *>(YY + (n_el_max + 1));  // zs_swz(n_el_max) + 1 values
    int8_t *sg = reinterpret_cast<int8_t *>(sm + zs_swz(n_el_max) + 1);  // ZS_TILE decisions
    int8_t *cg = sg + ZS_TILE;                                   // ZS_TILE certified decisions (MODE 2), 2 = none
    __shared__ double red[ZS_THREADS / 32][4];
    __shared__ int list[ZS_LIST_CAP];
    __shared__ int list_n;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileDesc td = tiles[blockIdx.x];
    const uint64_t base = off[td.item];
    const int64_t n = (int64_t)(off[td.item + 1] - base);
    const int64_t w0 = td.first;
    const int nw = (int)min((int64_t)ZS_TILE, n - w0);
    const int64_t lo0 = w0 - 1 - lag;  // series index held at smem position 0 (may be negative)
    const int n_el = nw + lag + 1;

    if (tid == 0) list_n = 0;
    const int first_valid = (lo0 < 0) ? (int)(-lo0) : 0;  // smem position of series index max(lo0, 0)
    {
        const float *xt = x + base + (lo0 < 0 ? 0 : lo0);  // element at smem position first_valid
        const float x0 = xt[0];                             // n_el > first_valid always (nw >= 1)
        int nf = 0;
        for (int p = tid; p < n_el; p += ZS_THREADS) {
            const int q = p - first_valid;
            const float v = (q >= 0) ? xt[q] : x0;  // positions before the series start are never part of a stats window
            sm[zs_swz(p)] = v;
            nf |= !(fabsf(v) <= 3.4028234664e38f);
        }
        // Inf / NaN: `filtered == data` no longer holds at influence 1 (stat.rs:42: 1 * x + 0 * filtered[i-1] is NaN behind a
        // non-finite value), so the series is walked again in order by the sequential kernel (launch_zscore)
        if (__syncthreads_or(nf) && tid == 0 && nonfinite) nonfinite[td.item] = 1;
    }

    const float flag = __ull2float_rn((unsigned long long)lag);  // `data.len() as f32`
    const float flagm1 = __fsub_rn(flag, 1.0f);                   // `len - 1.`
    // The certificate rests on the standard model fl(a op b) = (a op b)(1 + d), |d| <= u, which holds only while nothing the
    // reference computes in f32 overflows or underflows: thresholds in [2^-16, 2^30] here, and per tile max|x| in
    // [2^-30, 2^50] below (then sum (x - m)^2 <= 4 n M^2 < 2^116 stays finite, squares that flush to zero are below
    // 2^-126 against an error budget `extra` >= 1e-7 M^2, and thr * sd is either exactly 0 or a normal number).  Everything
    // else — huge or tiny magnitudes, Inf, NaN — is replayed exactly (stat.rs:8-14 then gives sd = inf or NaN: no signal).
    const bool use_filter = (MODE != 1) && lag >= 2 && threshold >= 1.52587890625e-5f && threshold <= 1073741824.0f;
    bool replay_all = !use_filter || MODE == 2;
    bool cert_ok = use_filter;

    if (use_filter) {
        // ---- f64 prefix sums of y = x - xref and y^2 over the tile's elements (thread-contiguous chunks) ----
        const double xref = (double)sm[zs_swz(first_valid)];
        const int per = ((n_el + ZS_THREADS - 1) / ZS_THREADS) | 1;  // odd stride: conflict-free chunk starts
        const int pb = min(tid * per, n_el), pe = min(pb + per, n_el);
        double a1 = 0.0, a2 = 0.0;
        float mxf = 0.0f;
#pragma unroll 3
        for (int p = pb; p < pe; ++p) {
            const float xf = sm[zs_swz(p)];
            const double y = (double)xf - xref;
            a1 += y;
            a2 = fma(y, y, a2);
            mxf = fmaxf(mxf, (fabsf(xf) <= 3.4028234664e38f) ? fabsf(xf) : __int_as_float(0x7F800000));   // NaN counts as Inf
        }
        double s1 = a1, s2 = a2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double t1 = __shfl_up_sync(0xFFFFFFFFu, s1, d), t2 = __shfl_up_sync(0xFFFFFFFFu, s2, d);
            if (lane >= d) { s1 += t1; s2 += t2; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mxf = fmaxf(mxf, __shfl_xor_sync(0xFFFFFFFFu, mxf, d));
        if (lane == 31) { red[warp][0] = s1; red[warp][1] = s2; }
        if (lane == 0) red[warp][2] = (double)mxf;
        __syncthreads();
        double b1 = s1 - a1, b2 = s2 - a2, M = 0.0;  // exclusive within the warp; M = max|x| over the tile
#pragma unroll
        for (int w = 0; w < ZS_THREADS / 32; ++w) {
            if (w < warp) { b1 += red[w][0]; b2 += red[w][1]; }
            M = fmax(M, red[w][2]);
        }
#pragma unroll 3
        for (int p = pb; p < pe; ++p) {
            YY[p] = make_double2(b1, b2);
            const double y = (double)sm[zs_swz(p)] - xref;
            b1 += y;
            b2 = fma(y, y, b2);
        }
        if (pe == n_el && pb < pe) YY[n_el] = make_double2(b1, b2);
        __syncthreads();
        cert_ok = M >= 9.313225746154785e-10 && M <= 1125899906842624.0;   // block-uniform; false for M = Inf (or NaN data)
        if (!cert_ok) replay_all = true;

        // ---- certificate, one window per thread, branch-free except for the rare inconclusive case ----
        const double slack = fma(zc.c_slack, M, 1e-40);
        const double err_S = fma(zc.c_errS, M * M, 1e-40);
        const double extra = fma(zc.c_extra, M * M, 2e-40);
        const int k_min = (w0 >= (int64_t)lag) ? 0 : (int)((int64_t)lag - w0);  // windows before `lag` never signal
        unsigned unc = 0;
#pragma unroll 2
        for (int k = tid; k < nw && cert_ok; k += ZS_THREADS) {
            // a - lo0 with a = max(i-1-lag, 0) (stat.rs:30-31 vs :51-52); the min() only bites for windows before `lag`
            const int pa = min(max(k, first_valid), n_el - lag);
            const double2 A = YY[pa], B = YY[pa + lag];
            const double s1w = B.x - A.x, s2w = B.y - A.y;
            const double mu_y = s1w * zc.inv_n;
            const double S = fma(-s1w, mu_y, s2w);
            const double d = ((double)sm[zs_swz(k + lag + 1)] - xref) - mu_y;
            const double dev = fabs(d);
            const double lo = dev - slack, hi = dev + slack;
            const bool yes = (lo > 0.0) & (lo * lo > zc.k_hi * (S + extra));
            const bool no = hi * hi < zc.k_lo * fmax(S - err_S, 0.0);
            int dec = yes ? (d > 0.0 ? 1 : -1) : (no ? 0 : 2);
            if (k < k_min) dec = 0;
            if (dec == 2) {
                ++unc;
                if (MODE == 0) {
                    const int slot = atomicAdd(&list_n, 1);
                    if (slot < ZS_LIST_CAP) list[slot] = k;
                }
            }
            if (MODE == 0) sg[k] = (int8_t)dec; else cg[k] = (int8_t)dec;
        }
        if (stats) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) unc += __shfl_xor_sync(0xFFFFFFFFu, unc, d);
            if (lane == 0 && unc) atomicAdd(&stats[0], (unsigned long long)unc);
        }
        __syncthreads();
        if (MODE == 0 && cert_ok) {
            const int cnt = list_n;
            if (cnt > ZS_LIST_CAP) replay_all = true;
            else {
                auto ld = [&](int64_t e) { return sm[zs_swz((int)(e - lo0))]; };
                for (int t = tid; t < cnt; t += ZS_THREADS) {
                    const int k = list[t];
                    const int64_t i = w0 + k;
                    const int64_t a = (i - 1 - (int64_t)lag > 0) ? (i - 1 - (int64_t)lag) : 0;  // stat.rs:30-31 vs :51-52
                    sg[k] = (int8_t)zscore_decide(ld, a, (uint64_t)lag, flag, flagm1, ld(i), threshold);
                }
            }
        }
    }

    if (replay_all) {  // block-uniform
        const int p0 = tid * ZS_R;   // smem position of chain 0's first element
        const int64_t i0 = w0 + p0;  // first window of this thread
        if (p0 < nw) {
            int rs[ZS_R];
#pragma unroll
            for (int r = 0; r < ZS_R; ++r) rs[r] = 0;
            if ((i0 >= (int64_t)lag + 1) && (p0 + ZS_R <= nw) && (lag >= ZS_R)) {
                replay_block(sm, p0, lag, flag, flagm1, threshold, rs);
            } else {  // ctg head (i <= lag), tile tail, or tiny lag: one window at a time, same arithmetic
                auto ld = [&](int64_t e) { return sm[zs_swz((int)(e - lo0))]; };
                for (int r = 0; r < ZS_R; ++r) {
                    const int64_t i = i0 + r;
                    if (i >= n || i < (int64_t)lag) continue;
                    const int64_t a = (i - 1 - (int64_t)lag > 0) ? (i - 1 - (int64_t)lag) : 0;
                    rs[r] = zscore_decide(ld, a, (uint64_t)lag, flag, flagm1, ld(i), threshold);
                }
            }
            unsigned bad = 0;
#pragma unroll
            for (int r = 0; r < ZS_R; ++r) {
                if (p0 + r >= nw) continue;
                if (MODE == 2 && cert_ok && cg[p0 + r] != 2 && cg[p0 + r] != (int8_t)rs[r]) ++bad;
                sg[p0 + r] = (int8_t)rs[r];
            }
            if (MODE == 2 && bad && stats) atomicAdd(&stats[1], (unsigned long long)bad);
        }
    }
    __syncthreads();
    for (int p = tid; p < nw; p += ZS_THREADS) signal[base + w0 + p] = sg[p];
}

// ------------------------------------------------------------------------------------------------
// any influence: one warp per series, speculative 32-window look-ahead, `filt` in global scratch
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
zscore_seq_kernel(const float *__restrict__ x_all, const uint64_t *__restrict__ off, int64_t s_begin, int64_t s_end,
                  uint64_t lag, float threshold, float influence, float *filt_all, int8_t *__restrict__ signal,
                  const int *__restrict__ only_flagged) {
    const int lane = threadIdx.x & 31;
    const int64_t c = s_begin + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (c >= s_end) return;
    if (only_flagged && !only_flagged[c]) return;
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

// zscore_decide on a ring of P floats, the window [a, a + L) cut where the ring wraps: two runs of consecutive addresses,
// so an element costs one LDS + one FADD in the first pass and LDS + FSUB + FMUL + FADD in the second (same order, same
// operations as zscore_decide)
template <int P>
__device__ __forceinline__ int ring_decide(const float *ring, int64_t a, uint32_t l32, float flag, float flagm1, float xv, float threshold) {
    const uint32_t r0 = (uint32_t)a & (uint32_t)(P - 1);
    const uint32_t n1 = min(l32, (uint32_t)P - r0), n2 = l32 - n1;
    const float *p1 = ring + r0;
    float s = 0.0f;
#pragma unroll 4
    for (uint32_t t = 0; t < n1; ++t) s = __fadd_rn(s, p1[t]);
#pragma unroll 4
    for (uint32_t t = 0; t < n2; ++t) s = __fadd_rn(s, ring[t]);
    const float m = __fdiv_rn(s, flag);
    float q = 0.0f;
#pragma unroll 4
    for (uint32_t t = 0; t < n1; ++t) {
        const float d = __fsub_rn(p1[t], m);
        q = __fadd_rn(q, __fmul_rn(d, d));
    }
#pragma unroll 4
    for (uint32_t t = 0; t < n2; ++t) {
        const float d = __fsub_rn(ring[t], m);
        q = __fadd_rn(q, __fmul_rn(d, d));
    }
    const float sd = __fsqrt_rn(__fdiv_rn(q, flagm1));
    if (fabsf(__fsub_rn(xv, m)) > __fmul_rn(threshold, sd)) return (xv > m) ? 1 : -1;
    return 0;
}

// The same walk with the history in SHARED memory: a ring of P floats per warp (P a power of two >= lag + 65) holds the
// committed filtered values of the last lag + 1 positions and the speculated x of the next 32, so a lane's 2 * lag
// reads per window are shared-memory loads of consecutive addresses (conflict-free across the lanes) and nothing but x is read
// from, and nothing but the signals written to, global memory (the kernel above moves ~400 global loads per window).
constexpr int ZSEQ_WARPS = 4;
template <int P>
__global__ void __launch_bounds__(ZSEQ_WARPS * 32)
zscore_seq_ring_kernel(const float *__restrict__ x_all, const uint64_t *__restrict__ off, int64_t s_begin, int64_t s_end,
                       uint64_t lag, float threshold, float influence, int8_t *__restrict__ signal,
                       const int *__restrict__ only_flagged) {
    __shared__ float ring_all[ZSEQ_WARPS][P];
    const int lane = threadIdx.x & 31;
    const int64_t c = s_begin + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (c >= s_end) return;
    if (only_flagged && !only_flagged[c]) return;
    float *ring = ring_all[threadIdx.x >> 5];
    const uint64_t base = off[c];
    const int64_t n = (int64_t)(off[c + 1] - base);
    const float *x = x_all + base;
    int8_t *sg = signal + base;
    const int64_t L = (int64_t)lag;
    for (int64_t j = lane; j < L && j < n; j += 32) {
        ring[j & (P - 1)] = x[j];           // filtered == data for the first lag values (stat.rs:24-27)
        sg[j] = 0;
    }
    const float flag = __ull2float_rn((unsigned long long)lag);
    const float flagm1 = __fsub_rn(flag, 1.0f);
    const float one_minus = __fsub_rn(1.0f, influence);
    int64_t i = L, front = L;               // ring holds positions [i - L - 1 (or 0), front): committed below i, speculated x from i on
    while (i < n) {
        for (int64_t pos = front + lane; pos < i + 32 && pos < n; pos += 32) ring[pos & (P - 1)] = x[pos];
        front = (i + 32 < n) ? i + 32 : n;
        __syncwarp();
        const int64_t w = i + lane;
        const bool active = w < n;
        int my = 0;
        float xv = 0.0f;
        if (active) {
            xv = ring[w & (P - 1)];
            const int64_t a = (w - 1 - L > 0) ? (w - 1 - L) : 0;
            my = ring_decide<P>(ring, a, (uint32_t)lag, flag, flagm1, xv, threshold);
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
            const float prev = (lane == 0) ? ring[(i - 1) & (P - 1)] : prev_x;  // predecessor did not signal => filtered == x
            fw = __fadd_rn(__fmul_rn(influence, xv), __fmul_rn(one_minus, prev));
        }
        const float f_first = __shfl_sync(0xFFFFFFFFu, fw, first & 31);
        if (lane == first + 1 && my != 0) fw = __fadd_rn(__fmul_rn(influence, xv), __fmul_rn(one_minus, f_first));
        __syncwarp();                        // every lane has read the ring
        if (lane < n_commit) {
            ring[w & (P - 1)] = fw;          // committed: the filtered value replaces the speculated x
            sg[w] = (int8_t)my;
        }
        __syncwarp();
        i += n_commit;
    }
}

// influence != 1 without walking every window.  filtered differs from data only AT signals (stat.rs:40-48), so a window whose
// stats window holds no signal sees plain data and its decision is the influence-1 decision D1 (a pure function of the data,
// computed for all windows in parallel by zscore_indep_kernel beforehand, in `signal`).  The walk therefore alternates:
//   clean  - no signal among the last lag + 1 windows: jump to the next window with D1 != 0 (32 windows per ballot); every
//            window passed keeps D1 = 0, and the one found signals exactly as D1 says (its stats window is clean too);
//   dirty  - from there on windows are replayed in order with the true filtered history (the ring of the kernel above),
//            until lag + 1 windows in a row have not signalled.
// Only the dirty stretches (~40 % of a genome at threshold 3) pay the sequential replay.
// ---- the walk, cut into SEGMENTS -------------------------------------------------------------------------------------------
// One warp per series would leave the GPU to the longest series (a chromosome's last ctg is up to twice the others: warps active
// 44 %).  The walk of a series can be cut wherever the process is CLEAN — no signal among the last lag + 1 windows: the
// history is plain data there, whatever happened before.  Where that is cannot be known without walking, so it is assumed
// at every multiple of `seg` windows and checked afterwards:
//   zscore_seg_walk_kernel  one warp per segment [s, e): the walk above, started clean at s, stopped at e; it leaves the window
//                           of its last signal and, if that is within lag + 1 windows of e, the filtered history of
//                           [e - lag - 1, e)
//   zscore_seg_fix_kernel   one warp per series, boundary by boundary.  Segment 0 started truly clean, so what it left is the
//                           true state at the first boundary.  If that state is clean the next segment's assumption held and
//                           its results (and what IT left) are true as well.  If not, the warp takes the true history over
//                           and replays the windows behind the boundary in order — reading the assumed walk's signals before
//                           it overwrites them — until the true process AND the assumed one are clean at the same window:
//                           from there on both see plain data, so the rest of the segment is right as it stands.  A replay
//                           that reaches the end of its segment simply goes on into the next one.
// The replays are short (the rest of one dirty stretch per failed boundary), the walk itself runs on 6 - 12 times as many warps.
constexpr int64_t ZS_NEVER = -(1ll << 40);

// one step of the replay: the next <= 32 windows from i on, evaluated as in zscore_seq_ring_kernel; commits n_commit <= room windows.
// Returns n_commit; last_sig / spec_last are updated (spec_last from the signals that stood in sg before: track_old).
template <int P, bool TRACK_OLD>
__device__ __forceinline__ int zs_replay_step(float *ring, const float *__restrict__ x, int8_t *sg, int64_t n, int64_t i, int64_t &front,
                                              int64_t room, int64_t L, uint32_t lag32, float flag, float flagm1, float threshold,
                                              float influence, float one_minus, int lane, int64_t &last_sig, int64_t &spec_last) {
    for (int64_t pos = front + lane; pos < i + 32 && pos < n; pos += 32) ring[pos & (P - 1)] = x[pos];
    front = (i + 32 < n) ? i + 32 : n;
    __syncwarp();
    const int64_t w = i + lane;
    const bool active = w < n;
    int my = 0;
    float xv = 0.0f;
    if (active) {
        xv = ring[w & (P - 1)];
        const int64_t a = (w - 1 - L > 0) ? (w - 1 - L) : 0;
        my = ring_decide<P>(ring, a, lag32, flag, flagm1, xv, threshold);
    }
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, my != 0);
    const int first = ballot ? (__ffs(ballot) - 1) : 32;
    int n_commit = (first >= 31) ? 32 : first + 2;     // lanes <= first + 1 saw only correctly speculated values
    if ((int64_t)n_commit > room) n_commit = (int)room;
    const float prev_x = __shfl_up_sync(0xFFFFFFFFu, xv, 1);
    float fw = xv;
    if (lane == first) {
        const float prev = (lane == 0) ? ring[(i - 1) & (P - 1)] : prev_x;
        fw = __fadd_rn(__fmul_rn(influence, xv), __fmul_rn(one_minus, prev));
    }
    const float f_first = __shfl_sync(0xFFFFFFFFu, fw, first & 31);
    if (lane == first + 1 && my != 0) fw = __fadd_rn(__fmul_rn(influence, xv), __fmul_rn(one_minus, f_first));
    if (first < n_commit) {
        last_sig = i + first;
        if (first + 1 < n_commit && ((ballot >> (first + 1)) & 1u)) last_sig = i + first + 1;
    }
    if (TRACK_OLD) {
        const unsigned ob = __ballot_sync(0xFFFFFFFFu, lane < n_commit && sg[w] != 0);
        if (ob) spec_last = i + (31 - __clz(ob));
    }
    __syncwarp();
    if (lane < n_commit) {
        ring[w & (P - 1)] = fw;
        sg[w] = (int8_t)my;
    }
    __syncwarp();
    return n_commit;
}

template <int P>
__global__ void __launch_bounds__(ZSEQ_WARPS * 32)
zscore_seg_walk_kernel(const float *__restrict__ x_all, const uint64_t *__restrict__ off, const uint32_t *__restrict__ seg_off,
                       int64_t s_begin, int64_t s_end, int64_t seg, uint64_t lag, float threshold, float influence, int8_t *signal,
                       long long *__restrict__ seg_last, float *__restrict__ seg_ring) {
    __shared__ float ring_all[ZSEQ_WARPS][P];
    const int lane = threadIdx.x & 31;
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= (int64_t)seg_off[s_end - s_begin]) return;
    int64_t lo = 0, hi = s_end - s_begin;            // the series of segment g
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)seg_off[mid] <= g) lo = mid; else hi = mid;
    }
    const int64_t c = s_begin + lo;
    float *ring = ring_all[threadIdx.x >> 5];
    const uint64_t base = off[c];
    const int64_t n = (int64_t)(off[c + 1] - base);
    const float *x = x_all + base;
    int8_t *sg = signal + base;
    const int64_t L = (int64_t)lag;
    const int64_t s = (g - (int64_t)seg_off[lo]) * seg, e = (s + seg < n) ? s + seg : n;
    const float flag = __ull2float_rn((unsigned long long)lag);
    const float flagm1 = __fsub_rn(flag, 1.0f);
    const float one_minus = __fsub_rn(1.0f, influence);
    int64_t i = s > L ? s : L;              // next window to decide; windows < L never signal (D1 wrote 0 there)
    int64_t last_sig = ZS_NEVER, spec_unused = ZS_NEVER;
    while (i < e) {
        // ---- clean: the next window whose D1 is a signal
        int64_t p = -1;
        while (i < e) {
            const int64_t w = i + lane;
            const unsigned b = __ballot_sync(0xFFFFFFFFu, w < e && sg[w] != 0);
            if (b) { p = i + (__ffs(b) - 1); break; }
            i += 32;
        }
        if (p < 0) break;
        // window p signals as D1 says; the ring takes the data of [p - L - 1, p) and the filtered value of p
        {
            const int64_t a = (p - L - 1 > 0) ? p - L - 1 : 0;
            for (int64_t q = a + lane; q < p; q += 32) ring[q & (P - 1)] = x[q];
            if (lane == 0) ring[p & (P - 1)] = __fadd_rn(__fmul_rn(influence, x[p]), __fmul_rn(one_minus, x[p - 1]));
        }
        last_sig = p;
        int64_t front = p + 1;
        i = p + 1;
        __syncwarp();
        // ---- dirty: sequential replay until the stats window of window i is clean again
        while (i < e && last_sig >= i - 1 - L)
            i += zs_replay_step<P, false>(ring, x, sg, n, i, front, e - i, L, (uint32_t)lag, flag, flagm1, threshold, influence, one_minus,
                                          lane, last_sig, spec_unused);
    }
    // ---- what the next segment has to know
    if (lane == 0) seg_last[g] = last_sig;
    if (e < n && last_sig >= e - 1 - L) {       // (the ring holds the filtered values of [e - L - 1, e): the replay ran up to e)
        float *dst = seg_ring + (size_t)g * (size_t)(L + 1);
        for (int64_t j = lane; j <= L; j += 32) dst[j] = ring[(e - L - 1 + j) & (P - 1)];
    }
}

template <int P>
__global__ void __launch_bounds__(ZSEQ_WARPS * 32)
zscore_seg_fix_kernel(const float *__restrict__ x_all, const uint64_t *__restrict__ off, const uint32_t *__restrict__ seg_off,
                      int64_t s_begin, int64_t s_end, int64_t seg, uint64_t lag, float threshold, float influence, int8_t *signal,
                      const long long *__restrict__ seg_last, const float *__restrict__ seg_ring) {
    __shared__ float ring_all[ZSEQ_WARPS][P];
    const int lane = threadIdx.x & 31;
    const int64_t ci = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ci >= s_end - s_begin) return;
    const int64_t c = s_begin + ci, g0 = (int64_t)seg_off[ci], n_seg = (int64_t)seg_off[ci + 1] - g0;
    if (n_seg < 2) return;
    float *ring = ring_all[threadIdx.x >> 5];
    const uint64_t base = off[c];
    const int64_t n = (int64_t)(off[c + 1] - base);
    const float *x = x_all + base;
    int8_t *sg = signal + base;
    const int64_t L = (int64_t)lag;
    const float flag = __ull2float_rn((unsigned long long)lag);
    const float flagm1 = __fsub_rn(flag, 1.0f);
    const float one_minus = __fsub_rn(1.0f, influence);
    int64_t k = 1;
    while (k < n_seg) {
        const int64_t s = k * seg;
        int64_t last_sig = seg_last[g0 + k - 1];
        if (last_sig < s - 1 - L) { ++k; continue; }       // the true process is clean at s: segment k is right as it stands
        // the true filtered history of [s - L - 1, s)
        {
            const float *src = seg_ring + (size_t)(g0 + k - 1) * (size_t)(L + 1);
            for (int64_t j = lane; j <= L; j += 32) ring[(s - L - 1 + j) & (P - 1)] = src[j];
        }
        __syncwarp();
        int64_t i = s, front = s, kk = k, seg_end = (s + seg < n) ? s + seg : n, spec_last = ZS_NEVER;
        while (i < n) {
            if (i == seg_end) {             // into the next segment with the TRUE state: its walk assumed a clean one
                ++kk;
                seg_end = (seg_end + seg < n) ? seg_end + seg : n;
                spec_last = ZS_NEVER;
            }
            if (last_sig < i - 1 - L && spec_last < i - 1 - L) break;     // both clean: identical from here on
            i += zs_replay_step<P, true>(ring, x, sg, n, i, front, seg_end - i, L, (uint32_t)lag, flag, flagm1, threshold, influence,
                                         one_minus, lane, last_sig, spec_last);
        }
        k = kk + 1;     // segment kk is right from i on, and so is what its walk left for the boundary behind it
    }
}

// ---- the segmented walk with ONE THREAD per segment and O(1) certified steps (GC fractions) ---------------------------------------
// The warp-per-segment walk above pays 2 * lag shared-memory loads and 4 * lag flops per replayed window, and inside a run of
// signals 32 evaluated windows commit two.  For data known to lie in [0, 1] (GC fractions) and 0 <= influence <= 1 a window needs
// none of that: the filtered history is bounded by M = 1.125 (zw_const), so the certificate of zscore_indep_kernel applies with a
// constant M, and its exact quantities — sum and sum of squares of the lag filtered values in the stats window — SLIDE: one value
// leaves, one enters (f64; the rounding of at most 3 * seg + 2 * lag + 64 such updates is part of the error budget).  A thread
// therefore walks its segment window by window: ring of the last lag + 1 filtered values in shared memory (column of the
// thread: conflict-free), S1 / S2 in registers, certificate, and only where that is inconclusive (a few windows per million) the
// reference's sums over the ring in the reference's order.  While the process is clean the decision is D1, as above.
//   zscore_tseg_walk_kernel  one thread per segment, assumed clean at its start; leaves seg_last / seg_ring as the kernels above
//   zscore_tseg_fix_kernel   one thread per boundary whose true state (left by the previous segment's walk) is dirty: replays
//                            behind the boundary until the true and the assumed process are clean at the same window.  A replay
//                            that is still dirty at the end of its segment would have to carry on into the next one, whose own
//                            boundary thread meanwhile started from what the assumed walk left: the series is flagged instead
//                            and walked again from scratch by zscore_seq_ring_kernel (one warp, the plain sequential walk).
//                            That takes a stretch of a whole segment (4096 windows) without lag + 1 quiet windows in a row.
// The certificate of zscore_indep_kernel multiplied through by n (D = n (x - mu) = n x - S1 and V = n S = n S2 - S1^2 are one
// operation away from the sliding sums, so a decision is four dependent f64 operations behind S1):
//   signal certainly      (|D| - ns)^2 > nk_hi V + c_hi   and |D| > ns           ns = n slack, nk_hi = n k_hi, c_hi = n^2 k_hi extra
//   certainly none        (|D| + ns)^2 < nk_lo max(V - n_errS, 0)                n_errS = n err_S
struct ZwConst {
    double dn, ns, nk_hi, c_hi, nk_lo, n_errS;
};

static ZwConst make_zw_const(uint64_t lag, float threshold, int64_t seg) {
    const double u = 5.9604644775390625e-8, u64e = 1.1102230246251565e-16;
    const float flag = (float)lag, flagm1 = flag - 1.0f;  // `data.len() as f32`, `len - 1.`
    const double dn = (double)lag, dnm1 = (double)flagm1;
    // M bounds every value of the filtered series: x in [0, 1], and f = fl(fl(a x) + fl(b f')) with a = influence in [2^-20, 1] or 0,
    // b = fl(1 - a): f <= (a + b f')(1 + u)^2 has its fixed point below 1 + 2^-4 + 2^-20
    const double M = 1.125;
    const double K = 3.0 * (double)seg + 2.0 * dn + 64.0;     // f64 operations that went into S1 / S2 since they were summed afresh
    const double cE = 1.01 * (dn + 2.0) * u;                  // E_m    = cE M
    const double cMu = 1.01 * (K * (dn + 1.0) / dn + 6.0) * u64e;          // err_mu = cMu M   (partial sums <= (n + 1) M)
    const double cS = (4.0 * K * (dn + 1.0) + 16.0 * dn) * u64e;           // err_S  = cS M^2
    const double R = (0.6 * dn + 8.0) * u;
    const double t2 = (double)threshold * (double)threshold;
    const double k_hi = t2 * (1.0 + R) * (1.0 + R) * (1.0 + 5.0 * u) / dnm1;
    const double k_lo = t2 * (1.0 - R) * (1.0 - R) * (1.0 - 5.0 * u) / dnm1;
    const double slack = (cE + cMu) * (1.0 + 1e-9) * M + 1e-40;      // on |x - mu|: E_m + the f64 error of mu (and of D / n)
    const double err_S = cS * M * M + 1e-40;                         // the f64 error of S (and of V / n)
    const double extra = (cS + dn * cE * cE) * (1.0 + 1e-9) * M * M + 2e-40;
    // scaled by n; every product below is rounded once more (relative 2^-53): widened by 1e-12 in the safe direction
    ZwConst c;
    c.dn = dn;
    c.ns = dn * slack * (1.0 + 1e-12);
    c.nk_hi = dn * k_hi * (1.0 + 1e-12);
    c.c_hi = dn * dn * k_hi * extra * (1.0 + 1e-12);
    c.nk_lo = dn * k_lo * (1.0 - 1e-12);
    c.n_errS = dn * err_S * (1.0 + 1e-12);
    return c;
}

// what a walking thread carries from window to window (window indices are relative to the series: < 2^31)
constexpr int ZW_NEVER = -(1 << 30);
struct ZwState {
    double S1, S2;       // sum / sum of squares of the filtered values in the stats window of the next window
    float prev;          // filtered value of the window before the next one
    double dprev;        // ... as a double
    int h;               // ring slot of the oldest value = the one that leaves next
    int last_sig;
};

// Window i (>= lag + 1): decision, then the slide to the stats window of i + 1.  HAVE_D1: while the process is clean the decision is
// d1 (the influence-1 decision), otherwise (and always without HAVE_D1) the certificate, and where that is inconclusive the
// reference's own sums in its order over the ring (oldest value first).  ring: this thread's column (stride NT), RP = lag + 1
// slots; `old` / `dold`: the value in slot st.h, which leaves (the caller may have loaded it long before: a slot is written only
// by the step that retires it).  One basic block but for the rare replay, and nothing of the next window's certificate waits
// for this window's decision except through S1 / S2 two windows on: consecutive steps overlap.
// f32 -> f64 by integer operations (the conversion unit takes 3 such conversions per window and is the slowest pipe of the walk):
// exact for zero and for normal positive numbers; anything else (a subnormal: a filtered value that decayed through a hundred
// signals in a row; negative, Inf, NaN: not GC data) raises `redo`
__device__ __forceinline__ double zw_f2d(float f, bool &redo) {
    const uint32_t b = __float_as_uint(f);
    redo |= ((b - 0x00800000u) >= 0x7F000000u) & (b != 0u);
    return __hiloint2double((int)(b ? (b >> 3) + 0x38000000u : 0u), (int)(b << 29));
}

// FAST: an inconclusive certificate only raises `redo` (the caller rolls the block back and walks it again with FAST = false): the
// step is one basic block, and eight of them in a row leave the compiler free to run the certificates side by side.
template <int NT, bool HAVE_D1, bool FAST = false>
__device__ __forceinline__ int zw_step(float *ring, int RP, ZwState &st, int i, int L, float xi, double dxi, float old, double dold, int d1,
                                       const ZwConst &zc, float flag, float flagm1, float threshold, float influence, float one_minus,
                                       bool *redo = nullptr) {
    const double D = fma(zc.dn, dxi, -st.S1);
    const double V = fma(-st.S1, st.S1, zc.dn * st.S2);
    const double dev = fabs(D);
    const double lo = dev - zc.ns, hi = dev + zc.ns;
    const bool yes = (lo > 0.0) & (lo * lo > fma(zc.nk_hi, V, zc.c_hi));
    const bool no = hi * hi < zc.nk_lo * fmax(V - zc.n_errS, 0.0);
    int dec = yes ? (D > 0.0 ? 1 : -1) : (no ? 0 : 2);
    if (HAVE_D1 && st.last_sig < i - 1 - L) dec = d1;
    if (FAST) {
        *redo |= dec == 2;
        dec = dec == 2 ? 0 : dec;
    } else if (dec == 2) {
        const int h = st.h;
        auto ld = [&](int64_t t) {
            int sl = h + (int)t;
            if (sl >= RP) sl -= RP;
            return ring[sl * NT];
        };
        dec = zscore_decide(ld, 0, (uint64_t)L, flag, flagm1, xi, threshold);
    }
    const float blend = __fadd_rn(__fmul_rn(influence, xi), __fmul_rn(one_minus, st.prev));
    double dblend;
    if (FAST) dblend = zw_f2d(blend, *redo);
    else dblend = (double)blend;
    const float fi = dec ? blend : xi;
    st.S1 = (st.S1 - dold) + st.dprev;
    st.S2 = fma(st.dprev, st.dprev, fma(-dold, dold, st.S2));
    ring[st.h * NT] = fi;
    st.h = (st.h + 1 == RP) ? 0 : st.h + 1;
    st.prev = fi;
    st.dprev = dec ? dblend : dxi;
    st.last_sig = dec ? i : st.last_sig;
    (void)old;
    return dec;
}
// the same with the leaving value loaded here
template <int NT, bool HAVE_D1>
__device__ __forceinline__ int zw_step1(float *ring, int RP, ZwState &st, int i, int L, float xi, int d1, const ZwConst &zc, float flag,
                                        float flagm1, float threshold, float influence, float one_minus) {
    const float old = ring[st.h * NT];
    return zw_step<NT, HAVE_D1>(ring, RP, st, i, L, xi, (double)xi, old, (double)old, d1, zc, flag, flagm1, threshold, influence, one_minus);
}

// series of segment g (seg_off: first segment of every series of the range)
__device__ __forceinline__ int64_t zw_series_of(const uint32_t *__restrict__ seg_off, int64_t n_series, int64_t g) {
    int64_t lo = 0, hi = n_series;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)seg_off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

template <int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
zscore_tseg_walk_kernel(const float *__restrict__ x_all, const uint64_t *__restrict__ off, const uint32_t *__restrict__ seg_off,
                        int64_t s_begin, int64_t s_end, int64_t seg, uint64_t lag, float threshold, float influence, const ZwConst zc,
                        int8_t *signal, long long *__restrict__ seg_last, float *__restrict__ seg_ring) {
    extern __shared__ float zw_ring[];      // (lag + 1) * NT
    const int64_t g = (int64_t)blockIdx.x * NT + threadIdx.x;
    if (g >= (int64_t)seg_off[s_end - s_begin]) return;
    const int64_t lo = zw_series_of(seg_off, s_end - s_begin, g);
    const int64_t c = s_begin + lo;
    const uint64_t base = off[c];
    const int64_t n = (int64_t)(off[c + 1] - base);
    const float *x = x_all + base;
    int8_t *sg = signal + base;
    const int L = (int)lag;
    const int RP = (int)lag + 1;
    float *ring = zw_ring + threadIdx.x;
    const int s = (int)((g - (int64_t)seg_off[lo]) * seg), e = (int)(((int64_t)s + seg < n) ? (int64_t)s + seg : n);
    const float flag = __ull2float_rn((unsigned long long)lag);
    const float flagm1 = __fsub_rn(flag, 1.0f);
    const float one_minus = __fsub_rn(1.0f, influence);
    ZwState st;
    st.S1 = 0.0, st.S2 = 0.0, st.h = 0, st.last_sig = ZW_NEVER, st.prev = 0.0f, st.dprev = 0.0;
    int i;
    if (s > L) {            // clean at s by assumption: the history [s - 1 - L, s - 1] is plain data
        if (s >= e) { seg_last[g] = ZS_NEVER; return; }
        const float *xs = x + (s - 1 - L);
        for (int j = 0; j <= L; ++j) {
            const float v = xs[j];
            ring[j * NT] = v;
            if (j < L) {
                st.S1 += (double)v;
                st.S2 = fma((double)v, (double)v, st.S2);
            }
            st.prev = v;
        }
        st.dprev = (double)st.prev;
        i = s;
    } else {                // the first segment of a series: windows < L never signal, window L sees [0, L) (stat.rs:30-31) and is D1
        if (L >= e) { seg_last[g] = ZS_NEVER; return; }
        for (int j = 0; j < L; ++j) {
            const float v = x[j];
            ring[j * NT] = v;
            st.S1 += (double)v;
            st.S2 = fma((double)v, (double)v, st.S2);
        }
        const float xl = x[L];
        const int dl = sg[L];
        const float fl = dl ? __fadd_rn(__fmul_rn(influence, xl), __fmul_rn(one_minus, x[L - 1])) : xl;
        ring[L * NT] = fl;
        st.prev = fl;
        st.dprev = (double)fl;
        if (dl) st.last_sig = L;
        i = L + 1;          // its stats window is [0, L) as well
    }
    // windows one by one up to an 8-aligned place of the arrays, then eight per 32-byte load of x and 8-byte load / store of the
    // signals.  The next eight are requested before the current eight are walked (a thread has nothing else to hide the latency
    // with), and the eight values that leave the ring during a block are loaded and converted up front: slot h + j is written
    // only by step j itself (RP >= 8).
    const uint64_t gi0 = base + (uint64_t)i;
    int head = (int)((8u - (uint32_t)(gi0 & 7u)) & 7u);
    if (head > e - i) head = e - i;
    for (const int stop = i + head; i < stop; ++i)
        sg[i] = (int8_t)zw_step1<NT, true>(ring, RP, st, i, L, x[i], sg[i], zc, flag, flagm1, threshold, influence, one_minus);
    float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa;
    uint2 dd = make_uint2(0u, 0u);
    if (i + 8 <= e) {
        xa = __ldg(reinterpret_cast<const float4 *>(x + i)), xb = __ldg(reinterpret_cast<const float4 *>(x + i) + 1);
        dd = *reinterpret_cast<const uint2 *>(sg + i);
    }
    for (; i + 8 <= e; i += 8) {
        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const uint2 dcur = dd;
        if (i + 16 <= e) {
            xa = __ldg(reinterpret_cast<const float4 *>(x + i + 8)), xb = __ldg(reinterpret_cast<const float4 *>(x + i + 8) + 1);
            dd = *reinterpret_cast<const uint2 *>(sg + i + 8);
        }
        float ov[8];
        double dxv[8], dov[8];
        // (volatile: the compiler cannot tell that the steps' stores go to other slots, and would leave every load next to its use)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int sl = st.h + j;
            if (sl >= RP) sl -= RP;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ov[j]) : "r"((uint32_t)__cvta_generic_to_shared(ring + sl * NT)) : "memory");
        }
        bool redo = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) dxv[j] = zw_f2d(xv[j], redo), dov[j] = zw_f2d(ov[j], redo);
        uint32_t o0 = 0u, o1 = 0u;
        const ZwState save = st;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int d1 = (int)(int8_t)(((j < 4) ? dcur.x : dcur.y) >> (8 * (j & 3)));
            const uint32_t dec = (uint32_t)zw_step<NT, true, true>(ring, RP, st, i + j, L, xv[j], dxv[j], ov[j], dov[j], d1, zc, flag, flagm1,
                                                                   threshold, influence, one_minus, &redo) & 0xFFu;
            if (j < 4) o0 |= dec << (8 * j); else o1 |= dec << (8 * (j - 4));
        }
        if (redo) {          // a few blocks per million: back to the state and the ring of the block's start, and once more with replays
            st = save;
            o0 = 0u, o1 = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int sl = save.h + j;
                if (sl >= RP) sl -= RP;
                ring[sl * NT] = ov[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d1 = (int)(int8_t)(((j < 4) ? dcur.x : dcur.y) >> (8 * (j & 3)));
                const uint32_t dec = (uint32_t)zw_step1<NT, true>(ring, RP, st, i + j, L, xv[j], d1, zc, flag, flagm1, threshold, influence, one_minus) & 0xFFu;
                if (j < 4) o0 |= dec << (8 * j); else o1 |= dec << (8 * (j - 4));
            }
        }
        *reinterpret_cast<uint2 *>(sg + i) = make_uint2(o0, o1);
    }
    for (; i < e; ++i) sg[i] = (int8_t)zw_step1<NT, true>(ring, RP, st, i, L, x[i], sg[i], zc, flag, flagm1, threshold, influence, one_minus);
    // ---- what the next segment has to know: the ring holds the filtered values of [e - L - 1, e), oldest at slot h
    seg_last[g] = st.last_sig == ZW_NEVER ? ZS_NEVER : (long long)st.last_sig;
    if (e < n && st.last_sig >= e - 1 - L) {
        float *dst = seg_ring + (size_t)g * (size_t)(L + 1);
        int sl = st.h;
        for (int j = 0; j <= L; ++j) {
            dst[j] = ring[sl * NT];
            sl = (sl + 1 == RP) ? 0 : sl + 1;
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT)
zscore_tseg_fix_kernel(const float *__restrict__ x_all, const uint64_t *__restrict__ off, const uint32_t *__restrict__ seg_off,
                       int64_t s_begin, int64_t s_end, int64_t seg, uint64_t lag, float threshold, float influence, const ZwConst zc,
                       int8_t *signal, const long long *__restrict__ seg_last, const float *__restrict__ seg_ring, int *__restrict__ redo) {
    extern __shared__ float zw_ring[];      // (lag + 1) * NT
    const int64_t g = (int64_t)blockIdx.x * NT + threadIdx.x;
    if (g >= (int64_t)seg_off[s_end - s_begin]) return;
    const int64_t lo = zw_series_of(seg_off, s_end - s_begin, g);
    const int64_t k = g - (int64_t)seg_off[lo];
    if (k == 0) return;                     // the first segment of a series started truly clean
    const int L = (int)lag;
    const int s = (int)(k * seg);
    ZwState st;
    const long long ls = seg_last[g - 1];
    if (ls < (long long)s - 1 - L) return;
    st.last_sig = (int)ls;    // the true process is clean at s: the segment is right as it stands
    const int64_t c = s_begin + lo;
    const uint64_t base = off[c];
    const int64_t n = (int64_t)(off[c + 1] - base);
    const float *x = x_all + base;
    int8_t *sg = signal + base;
    const int RP = (int)lag + 1;
    float *ring = zw_ring + threadIdx.x;
    const int e = (int)(((int64_t)s + seg < n) ? (int64_t)s + seg : n);
    const float flag = __ull2float_rn((unsigned long long)lag);
    const float flagm1 = __fsub_rn(flag, 1.0f);
    const float one_minus = __fsub_rn(1.0f, influence);
    // the true filtered history of [s - L - 1, s)
    st.S1 = 0.0, st.S2 = 0.0, st.h = 0, st.prev = 0.0f, st.dprev = 0.0;
    {
        const float *src = seg_ring + (size_t)(g - 1) * (size_t)(L + 1);
        for (int j = 0; j <= L; ++j) {
            const float v = src[j];
            ring[j * NT] = v;
            if (j < L) {
                st.S1 += (double)v;
                st.S2 = fma((double)v, (double)v, st.S2);
            }
            st.prev = v;
        }
        st.dprev = (double)st.prev;
    }
    int i = s, spec_last = ZW_NEVER;
    for (; i < e; ++i) {
        if (st.last_sig < i - 1 - L && spec_last < i - 1 - L) break;      // both clean: identical from here on
        if (sg[i] != 0) spec_last = i;                                     // what the assumed walk decided here
        sg[i] = (int8_t)zw_step1<NT, false>(ring, RP, st, i, L, x[i], 2, zc, flag, flagm1, threshold, influence, one_minus);
    }
    // still dirty at the end of the segment: the next boundary's thread started from what the assumed walk left, which is not the
    // true state — the series is walked again from scratch
    if (i == e && e < n && !(st.last_sig < e - 1 - L && spec_last < e - 1 - L)) redo[c] = 1;
}

// the walk alone: d_signal holds the influence-1 decisions of every window (from zscore_indep_kernel or the fused GC-wave kernel)
bool zscore_skip_ok(uint64_t lag) { return lag >= 1 && lag + 65 <= 1024; }
int32_t launch_zscore_skip(const SeriesLayout &L, cudaStream_t st, const float *d_x, uint64_t lag, float threshold, float influence,
                           int8_t *d_signal, uint64_t *launches, int64_t s0, int64_t s1, bool gc_values) {
    if (L.n_series == 0 || s1 <= s0) return GAMS_CUDA_OK;
    const char *seg_e = getenv("GAMS_CUDA_ZSEG");       // (tests: boundaries every few hundred windows)
    const int64_t seg_env = seg_e ? atoll(seg_e) : 0;
    int64_t seg = std::max<int64_t>(seg_env > 0 ? seg_env : 4096, (int64_t)lag + 34);   // (16384: 8.3 ms, 4096: 7.5, 1024: 7.8 per 3.1e8 windows)
    int32_t rc;
    // GC fractions (values in [0, 1]), an influence the bound of zw_const covers, a certificate-friendly threshold: one thread per segment
    const int nt = (lag + 1) * 4 * 128 <= 56 * 1024 ? 128 : ((lag + 1) * 4 * 64 <= 75 * 1024 ? 64 : 0);
    const bool threads = gc_values && nt && lag >= 8 && (influence == 0.0f || (influence >= 9.5367431640625e-7f && influence <= 1.0f)) &&
                         threshold >= 1.52587890625e-5f && threshold <= 1073741824.0f &&
                         (reinterpret_cast<uintptr_t>(d_x) & 31u) == 0 && (reinterpret_cast<uintptr_t>(d_signal) & 7u) == 0 &&
                         !getenv("GAMS_CUDA_ZWALK_WARPS");
    const size_t smem = (size_t)(lag + 1) * 4 * (size_t)nt;
    if (threads && seg_env <= 0) {
        // Every thread walks one segment and all segments take about as long, so the kernel's time is (waves of CTAs) x (segment
        // length): the segments are made just long enough for the CTAs to fit the device in one wave (4096 windows gave 593
        // CTAs on 592 places for the bench genome: two waves), or in the fewest waves that keep a segment below 8192 windows;
        // not below 4096 windows: a replay that is still dirty at the end of its segment sends the whole series to the sequential
        // kernel, and stretches of 1000 windows without lag + 1 quiet ones in a row do occur (a few per cent of them at threshold 3).
        int dev = 0, n_sm = 0, occ = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        if (nt == 128) {
            CUDA_TRY(cudaFuncSetAttribute(zscore_tseg_walk_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, zscore_tseg_walk_kernel<128>, nt, smem));
        } else {
            CUDA_TRY(cudaFuncSetAttribute(zscore_tseg_walk_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, zscore_tseg_walk_kernel<64>, nt, smem));
        }
        const int64_t places = (int64_t)std::max(occ, 1) * n_sm * nt;
        const int64_t seg_min = std::max<int64_t>(4096, (int64_t)lag + 34);
        auto count = [&](int64_t sg_len) {
            int64_t k = 0;
            for (int64_t c = s0; c < s1; ++c) k += std::max<int64_t>(((int64_t)(L.h_off[c + 1] - L.h_off[c]) + sg_len - 1) / sg_len, 1);
            return k;
        };
        for (int64_t waves = 1;; ++waves) {
            int64_t lo = seg_min, hi = 8192;             // the shortest segment whose CTAs fit `waves` waves
            if (count(hi) > places * waves) continue;    // (ends: count(8192) is finite)
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (count(mid) <= places * waves) hi = mid; else lo = mid + 1;
            }
            seg = lo;
            break;
        }
    }
    if ((rc = L.seg_tables(s0, s1, seg, lag, st))) return rc;
    const int64_t n_seg = (int64_t)L.h_seg_off.back();
    if (threads) {
        const ZwConst zc = make_zw_const(lag, threshold, seg);
        const unsigned grid = (unsigned)((n_seg + nt - 1) / nt), grid3 = (unsigned)((s1 - s0 + ZSEQ_WARPS - 1) / ZSEQ_WARPS);
        CUDA_TRY(cudaMemsetAsync(L.d_nonfinite + s0, 0, sizeof(int) * (size_t)(s1 - s0), st));
        auto go = [&](auto walk, auto fix) -> int32_t {
            CUDA_TRY(cudaFuncSetAttribute(walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CUDA_TRY(cudaFuncSetAttribute(fix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            walk<<<grid, nt, smem, st>>>(d_x, L.d_off, L.d_seg_off, s0, s1, seg, lag, threshold, influence, zc, d_signal, L.d_seg_last, L.d_seg_ring);
            fix<<<grid, nt, smem, st>>>(d_x, L.d_off, L.d_seg_off, s0, s1, seg, lag, threshold, influence, zc, d_signal, L.d_seg_last, L.d_seg_ring,
                                        L.d_nonfinite);
            return GAMS_CUDA_OK;
        };
        if ((rc = nt == 128 ? go(zscore_tseg_walk_kernel<128>, zscore_tseg_fix_kernel<128>) : go(zscore_tseg_walk_kernel<64>, zscore_tseg_fix_kernel<64>))) return rc;
        // series whose replay ran through a whole segment: the plain sequential walk, one warp each (warps of other series leave at once)
        if (lag + 65 <= 256)
            zscore_seq_ring_kernel<256><<<grid3, ZSEQ_WARPS * 32, 0, st>>>(d_x, L.d_off, s0, s1, lag, threshold, influence, d_signal, L.d_nonfinite);
        else
            zscore_seq_ring_kernel<1024><<<grid3, ZSEQ_WARPS * 32, 0, st>>>(d_x, L.d_off, s0, s1, lag, threshold, influence, d_signal, L.d_nonfinite);
        if (launches) *launches += 3;
        CUDA_TRY(cudaGetLastError());
        return GAMS_CUDA_OK;
    }
    const unsigned grid = (unsigned)((n_seg + ZSEQ_WARPS - 1) / ZSEQ_WARPS), grid2 = (unsigned)((s1 - s0 + ZSEQ_WARPS - 1) / ZSEQ_WARPS);
    if (lag + 65 <= 256) {
        zscore_seg_walk_kernel<256><<<grid, ZSEQ_WARPS * 32, 0, st>>>(d_x, L.d_off, L.d_seg_off, s0, s1, seg, lag, threshold, influence, d_signal,
                                                                     L.d_seg_last, L.d_seg_ring);
        zscore_seg_fix_kernel<256><<<grid2, ZSEQ_WARPS * 32, 0, st>>>(d_x, L.d_off, L.d_seg_off, s0, s1, seg, lag, threshold, influence, d_signal,
                                                                     L.d_seg_last, L.d_seg_ring);
    } else {
        zscore_seg_walk_kernel<1024><<<grid, ZSEQ_WARPS * 32, 0, st>>>(d_x, L.d_off, L.d_seg_off, s0, s1, seg, lag, threshold, influence, d_signal,
                                                                      L.d_seg_last, L.d_seg_ring);
        zscore_seg_fix_kernel<1024><<<grid2, ZSEQ_WARPS * 32, 0, st>>>(d_x, L.d_off, L.d_seg_off, s0, s1, seg, lag, threshold, influence, d_signal,
                                                                      L.d_seg_last, L.d_seg_ring);
    }
    if (launches) *launches += 2;
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

int32_t launch_zscore(const SeriesLayout &L, cudaStream_t st, const float *d_x, uint64_t lag, float threshold,
                      float influence, float *d_filtered, int8_t *d_signal, uint64_t *launches, int mode,
                      unsigned long long *d_stats, int64_t s0, int64_t s1, bool gc_values) {
    if (lag == 0) {
        set_error("thresholding: lag must be >= 1 (the reference underflows lag-1, stat.rs:30)");
        return GAMS_CUDA_EINVAL;
    }
    if (L.n_series == 0 || s1 <= s0) return GAMS_CUDA_OK;
    if (L.min_len < lag) {
        set_error("thresholding: a series has %llu values but lag is %llu (the reference panics, stat.rs:30)",
                  (unsigned long long)L.min_len, (unsigned long long)lag);
        return GAMS_CUDA_ELAG;
    }
    const bool tiled = influence == 1.0f && lag <= ZS_MAX_TILED_LAG;
    // influence != 1 with a lag the shared-memory ring holds: the influence-1 decisions of every window first (parallel),
    // then the walk that replays only the stretches behind a signal
    const bool skip = !tiled && lag <= ZS_MAX_TILED_LAG && zscore_skip_ok(lag) && mode == 0;
    if (tiled && !gc_values) CUDA_TRY(cudaMemsetAsync(L.d_nonfinite + s0, 0, sizeof(int) * (size_t)(s1 - s0), st));
    if (tiled || skip) {
        const int n_el_max = ZS_TILE + (int)lag + 1;
        const size_t smem_need = (size_t)(n_el_max + 1) * 16 + (size_t)(zs_swz(n_el_max) + 1) * sizeof(float) + 2 * ZS_TILE;
        const ZsConst zc = make_zs_const(lag, threshold);
        auto go = [&](auto kern) -> int32_t {
            if (smem_need > 48 * 1024)
                CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need));
            const uint64_t t0 = L.h_ztile_off[s0], t1 = L.h_ztile_off[s1];
            if (t1 > t0)
                kern<<<(unsigned)(t1 - t0), ZS_THREADS, smem_need, st>>>(d_x, L.d_off, L.d_ztiles + t0, (int)lag, threshold, zc,
                                                                        d_signal, d_stats, tiled ? L.d_nonfinite : nullptr);
            return GAMS_CUDA_OK;
        };
        int32_t rc = mode == 1 ? go(zscore_indep_kernel<1>) : mode == 2 ? go(zscore_indep_kernel<2>) : go(zscore_indep_kernel<0>);
        if (rc) return rc;
        if (skip && (rc = launch_zscore_skip(L, st, d_x, lag, threshold, influence, d_signal, launches, s0, s1, gc_values))) return rc;
        if (tiled && !gc_values) {
            // series flagged as holding Inf / NaN are walked again in the reference's order (a warp per series; warps of
            // clean series leave at once).  Data known to be GC fractions cannot hold any.
            const unsigned grid = (unsigned)((s1 - s0 + ZSEQ_WARPS - 1) / ZSEQ_WARPS);
            if (lag + 65 <= 256)
                zscore_seq_ring_kernel<256><<<grid, ZSEQ_WARPS * 32, 0, st>>>(d_x, L.d_off, s0, s1, lag, threshold, influence, d_signal, L.d_nonfinite);
            else if (lag + 65 <= 1024)
                zscore_seq_ring_kernel<1024><<<grid, ZSEQ_WARPS * 32, 0, st>>>(d_x, L.d_off, s0, s1, lag, threshold, influence, d_signal, L.d_nonfinite);
            else if (d_filtered)
                zscore_seq_kernel<<<grid, 128, 0, st>>>(d_x, L.d_off, s0, s1, lag, threshold, influence, d_filtered, d_signal, L.d_nonfinite);
            // (lag > 959 without the d_filtered scratch: series with Inf / NaN keep the filtered == data decisions)
            if (launches) ++*launches;
        }
    } else {
        if (!d_filtered) {
            set_error("thresholding: influence != 1 (or lag too large for the tiled path) needs the d_filtered scratch");
            return GAMS_CUDA_EINVAL;
        }
        const unsigned grid = (unsigned)((s1 - s0 + ZSEQ_WARPS - 1) / ZSEQ_WARPS);
        if (lag + 65 <= 256)
            zscore_seq_ring_kernel<256><<<grid, ZSEQ_WARPS * 32, 0, st>>>(d_x, L.d_off, s0, s1, lag, threshold, influence, d_signal, nullptr);
        else if (lag + 65 <= 1024)
            zscore_seq_ring_kernel<1024><<<grid, ZSEQ_WARPS * 32, 0, st>>>(d_x, L.d_off, s0, s1, lag, threshold, influence, d_signal, nullptr);
        else
            zscore_seq_kernel<<<grid, 128, 0, st>>>(d_x, L.d_off, s0, s1, lag, threshold, influence, d_filtered, d_signal, nullptr);
    }
    if (launches) ++*launches;
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
