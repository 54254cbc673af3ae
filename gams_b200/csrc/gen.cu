// gen.cu — the ambiguity scan of `gams gen`, src/cmd_gams/gen.rs:80-104.
//
// gen walks every base of a chromosome, collects the positions that are not one of A C G T a c g t into an IntSpan
// (gen.rs:86-93) and derives the valid regions as its complement (gen.rs:100-102).  Here the sequence is classified
// 32 bytes per lane per load (the same streaming access as Kernel A) into one valid bit per base, and only the
// boundaries of the valid runs leave the device: run k is [start[k], end[k]] in 0-based byte offsets of the
// concatenated sequence, never crossing a chromosome boundary.  fill / excise / piece cutting (gen.rs:103-126) work on
// that short list on the host (gams_cuda_gen_ctgs in api.cu).
//
// Bound: HBM by design (1 B/base read, a few bytes per run written); the four-way byte compare makes it ALU-heavy.
#include "gc_common.cuh"

namespace gams {

constexpr int GEN_THREADS = 256;
constexpr int GEN_TILE = GEN_THREADS * 32;   // bytes per CTA

// bit 7 of each byte <=> the byte is one of A C G T a c g t
__device__ __forceinline__ uint32_t valid_flags(uint32_t w) {
    const uint32_t t = w & 0xDFDFDFDFu;                               // fold case: a-z -> A-Z (and 0x60-0x7F -> 0x40-0x5F)
    uint32_t nz = 0xFFFFFFFFu;                                        // bit 7: byte differs from every target so far
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = t ^ (k == 0 ? 0x41414141u : k == 1 ? 0x43434343u : k == 2 ? 0x47474747u : 0x54545454u);
        nz &= ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;                  // bit 7 set <=> x != 0
    }
    // folding maps `{`-class bytes (0x61^0x20...) fine, but 0x01-0x1F|0x20.. cannot alias: only bit 5 was cleared, so a
    // byte matches a target iff it is the target or the target | 0x20, i.e. the lower-case letter
    return ~nz & 0x80808080u;
}

__device__ __forceinline__ bool valid_byte(uint8_t b) {
    const uint8_t t = b & 0xDFu;
    return t == 0x41 || t == 0x43 || t == 0x47 || t == 0x54;
}

// MODE 0: count run starts / ends per tile; MODE 1: write them (base = exclusive scans of the counts)
template <int MODE>
__global__ void __launch_bounds__(GEN_THREADS)
valid_runs_kernel(const uint8_t *__restrict__ seq, uint64_t total, const uint64_t *__restrict__ chr_off, int64_t n_chr,
                  uint64_t *__restrict__ cnt_s, uint64_t *__restrict__ cnt_e, const uint64_t *__restrict__ base_s,
                  const uint64_t *__restrict__ base_e, uint64_t cap, uint64_t *__restrict__ out_s, uint64_t *__restrict__ out_e) {
    __shared__ uint32_t ws[GEN_THREADS / 32], we[GEN_THREADS / 32];
    const uint64_t p = (uint64_t)blockIdx.x * GEN_TILE + (uint64_t)threadIdx.x * 32;
    uint32_t v = 0, starts = 0, ends = 0;
    if (p < total) {
        const Bytes32 x = ld_stream32(seq + p);                       // the allocation is padded to a multiple of 32 (+32)
#pragma unroll
        for (int q = 0; q < 8; ++q) v |= ((valid_flags(x.w[q]) * 0x00204081u) >> 28) << (4 * q);
        if (total - p < 32) v &= (1u << (uint32_t)(total - p)) - 1u;  // bytes past the end are not bases
        // chromosome starts inside (p, p + 32]: a run never continues across them
        int64_t lo = 0, hi = n_chr;                                   // last chromosome starting at or before p
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (chr_off[mid] <= p) lo = mid; else hi = mid;
        }
        uint32_t brk = (chr_off[lo] == p) ? 1u : 0u;                  // bit i: position p + i starts a chromosome
        bool brk_next = false;                                        // position p + 32 starts a chromosome
        for (int64_t c = lo + 1; c <= n_chr; ++c) {
            const uint64_t b = chr_off[c];
            if (b > p + 32) break;
            if (b == p + 32) brk_next = true;
            else brk |= 1u << (uint32_t)(b - p);
        }
        const bool prev_valid = p > 0 && !(brk & 1u) && valid_byte(seq[p - 1]);
        const bool next_valid = p + 32 < total && !brk_next && valid_byte(seq[p + 32]);
        const uint32_t prev = ((v << 1) | (prev_valid ? 1u : 0u)) & ~brk;                  // bit i: byte i-1 is valid and in the same chromosome
        const uint32_t next = ((v >> 1) | (next_valid ? 0x80000000u : 0u)) & ~(brk >> 1);  // bit i: byte i+1 likewise
        starts = v & ~prev;
        ends = v & ~next;
    }
    const uint32_t ms = __popc(starts), me = __popc(ends);
    uint32_t is = ms, ie = me;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, is, d), b = __shfl_up_sync(0xFFFFFFFFu, ie, d);
        if (lane >= d) { is += a; ie += b; }
    }
    if (lane == 31) { ws[warp] = is; we[warp] = ie; }
    __syncthreads();
    uint32_t bs = is - ms, be = ie - me;
    for (int w = 0; w < warp; ++w) { bs += ws[w]; be += we[w]; }
    if (MODE == 0) {
        if (threadIdx.x == GEN_THREADS - 1) { cnt_s[blockIdx.x] = bs + ms; cnt_e[blockIdx.x] = be + me; }
        return;
    }
    uint64_t ps = base_s[blockIdx.x] + bs, pe = base_e[blockIdx.x] + be;
    for (uint32_t m = starts; m; m &= m - 1, ++ps)
        if (ps < cap) out_s[ps] = p + (uint32_t)(__ffs(m) - 1);
    for (uint32_t m = ends; m; m &= m - 1, ++pe)
        if (pe < cap) out_e[pe] = p + (uint32_t)(__ffs(m) - 1);
}

__global__ void __launch_bounds__(1024) scan2_u64_kernel(const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, uint64_t n,
                                                         uint64_t *__restrict__ oa, uint64_t *__restrict__ ob) {
    __shared__ uint64_t wsum[2][32];
    __shared__ uint64_t carry[2];
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t va = i < n ? a[i] : 0, vb = i < n ? b[i] : 0;
        uint64_t ia = va, ib = vb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t ta = __shfl_up_sync(0xFFFFFFFFu, ia, d), tb = __shfl_up_sync(0xFFFFFFFFu, ib, d);
            if ((threadIdx.x & 31) >= d) { ia += ta; ib += tb; }
        }
        if ((threadIdx.x & 31) == 31) { wsum[0][threadIdx.x >> 5] = ia; wsum[1][threadIdx.x >> 5] = ib; }
        __syncthreads();
        uint64_t ba = carry[0], bb = carry[1];
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) { ba += wsum[0][w]; bb += wsum[1][w]; }
        if (i < n) { oa[i] = ba + ia - va; ob[i] = bb + ib - vb; }
        __syncthreads();
        if (threadIdx.x == 1023) { carry[0] = ba + ia; carry[1] = bb + ib; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { oa[n] = carry[0]; ob[n] = carry[1]; }
}

// d_tmp: 4 * (tiles + 1) u64.  *d_n_runs (device) receives the number of valid runs.
int32_t launch_valid_runs(cudaStream_t st, const uint8_t *d_seq, uint64_t total, const uint64_t *d_chr_off, int64_t n_chr,
                          uint64_t *d_tmp, uint64_t cap, uint64_t *d_run_start, uint64_t *d_run_end, uint64_t *d_n_runs) {
    const uint64_t tiles = (total + GEN_TILE - 1) / GEN_TILE;
    if (tiles == 0) {
        CUDA_TRY(cudaMemsetAsync(d_n_runs, 0, sizeof(uint64_t), st));
        return GAMS_CUDA_OK;
    }
    if (tiles >= (1ull << 31) || (reinterpret_cast<uintptr_t>(d_seq) & 31u)) {
        set_error("valid_runs: sequence too long for one launch or not 32-byte aligned");
        return GAMS_CUDA_EINVAL;
    }
    uint64_t *cs = d_tmp, *ce = cs + (tiles + 1), *bs = ce + (tiles + 1), *be = bs + (tiles + 1);
    valid_runs_kernel<0><<<(unsigned)tiles, GEN_THREADS, 0, st>>>(d_seq, total, d_chr_off, n_chr, cs, ce, nullptr, nullptr, 0, nullptr, nullptr);
    scan2_u64_kernel<<<1, 1024, 0, st>>>(cs, ce, tiles, bs, be);
    valid_runs_kernel<1><<<(unsigned)tiles, GEN_THREADS, 0, st>>>(d_seq, total, d_chr_off, n_chr, nullptr, nullptr, bs, be, cap, d_run_start, d_run_end);
    CUDA_TRY(cudaMemcpyAsync(d_n_runs, bs + tiles, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
