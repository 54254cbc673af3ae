// parse.cu — the lines of a range file, parsed on the device.
//
// The front end of `gams locate` (src/cmd_gams/locate.rs:111-117) and of read_range (src/libs/utils.rs:39-52): every
// line goes through intspan::Range::from_str (intspan 0.7.7, a regex) and is dropped unless is_valid().  At 10^8 lines that
// parse is what the reference spends its time on.  Here: newline scan -> line offsets (two small passes
// and a prefix sum), then one thread per line.
//
// Grammar.  Range::from_str searches  (?:(\w+)\.)?([\w/-]+)(?:\((.+)\))?:(\d+)[_-]?(\d+)?  anywhere in the line.  The
// device accepts the STRICT form gams' own tools write, anchored at the start of the line,
//     [name.]chr[(+)|(-)]:start[-end | _end]          (anything after it is ignored, as by the regex)
// and classifies every line:  >= 0 index of chr in the caller's name table;  -1 chromosome not in the table;
// -2 not a valid range (no ':' at all, start == 0, end < start: the reference skips these);  -3 outside the strict form
// (the regex might still match somewhere: the caller passes exactly these lines to Range::from_str).  Lines of the
// strict form get the values the regex gives — that is what tests/ checks against the oracle's restatement.
//
// Bound: HBM (the text is read twice, ~25 B per line, + 16 B per line written).
#include "common.cuh"
#include <memory>

namespace gams {

constexpr int PR_THREADS = 256;
constexpr int PR_BYTES = 16;                       // bytes per thread in the newline passes
constexpr int PR_TILE = PR_THREADS * PR_BYTES;

__device__ __forceinline__ uint32_t nl_mask16(const uint8_t *__restrict__ text, uint64_t base, uint64_t n) {
    uint32_t m = 0;
    if (base + PR_BYTES <= n && ((reinterpret_cast<uintptr_t>(text + base) & 15u) == 0)) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + base));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // byte == '\n' (0x0A): zero-byte test on w ^ 0x0A0A0A0A, exact per byte (no cross-byte borrow)
            const uint32_t x = w[k] ^ 0x0A0A0A0Au;
            const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;   // bit 7 of each zero byte
            m |= ((z * 0x00204081u) >> 28) << (4 * k);
        }
    } else {
        for (int k = 0; k < PR_BYTES; ++k)
            if (base + k < n && text[base + k] == '\n') m |= 1u << k;
    }
    return m;
}

// pass 1: newlines per tile of PR_TILE bytes
__global__ void __launch_bounds__(PR_THREADS)
nl_count_kernel(const uint8_t *__restrict__ text, uint64_t n, uint32_t *__restrict__ tile_cnt) {
    __shared__ uint32_t ws[PR_THREADS / 32];
    const uint64_t base = ((uint64_t)blockIdx.x * PR_THREADS + threadIdx.x) * PR_BYTES;
    uint32_t c = base < n ? (uint32_t)__popc(nl_mask16(text, base, n)) : 0u;
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < PR_THREADS / 32; ++w) t += ws[w];
        tile_cnt[blockIdx.x] = t;
    }
}

// exclusive prefix over the tile counts (one CTA; n_tiles is n_bytes / 4096)
__global__ void __launch_bounds__(1024)
nl_scan_kernel(const uint32_t *__restrict__ tile_cnt, uint64_t n_tiles, uint64_t *__restrict__ tile_base, uint64_t *__restrict__ total) {
    __shared__ uint64_t ws[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t b = 0; b < n_tiles; b += 1024) {
        const uint64_t i = b + threadIdx.x;
        const uint64_t v = i < n_tiles ? tile_cnt[i] : 0;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((threadIdx.x & 31) >= d) inc += t;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint64_t before = carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += ws[w];
        if (i < n_tiles) tile_base[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// pass 2: line k + 1 starts behind the k-th newline (line 0 starts at byte 0)
__global__ void __launch_bounds__(PR_THREADS)
nl_scatter_kernel(const uint8_t *__restrict__ text, uint64_t n, const uint64_t *__restrict__ tile_base, uint64_t cap_lines,
                  uint64_t *__restrict__ line_off) {
    __shared__ uint32_t ws[PR_THREADS / 32];
    const uint64_t base = ((uint64_t)blockIdx.x * PR_THREADS + threadIdx.x) * PR_BYTES;
    const uint32_t m = base < n ? nl_mask16(text, base, n) : 0u;
    const uint32_t c = (uint32_t)__popc(m);
    uint32_t inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if ((threadIdx.x & 31) >= d) inc += t;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint64_t k = tile_base[blockIdx.x] + inc - c;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) k += ws[w];
    uint32_t mm = m;
    while (mm) {
        const int b = __ffs(mm) - 1;
        mm &= mm - 1;
        ++k;                                            // this newline ends line k - 1 and starts line k
        if (k <= cap_lines) line_off[k] = base + b + 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) line_off[0] = 0;
}

__device__ __forceinline__ bool is_word(uint8_t c) {
    return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
}

// one thread per line
__global__ void __launch_bounds__(256)
parse_lines_kernel(const uint8_t *__restrict__ text, uint64_t n_bytes, const uint64_t *__restrict__ line_off, uint64_t n_lines,
                   const uint8_t *__restrict__ names, const uint32_t *__restrict__ name_off, const int32_t *__restrict__ htab,
                   uint32_t hmask, int32_t *__restrict__ q_chr, uint32_t *__restrict__ q_start, uint32_t *__restrict__ q_end) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += stride) {
        uint64_t p = line_off[i], e = (i + 1 < n_lines) ? line_off[i + 1] - 1 : n_bytes;   // [p, e): without the '\n'
        if (i + 1 == n_lines && e > p && text[e - 1] == '\n') --e;                         // the last line may or may not end in '\n'
        if (e > p && text[e - 1] == '\r') --e;
        int32_t res = -3;
        uint32_t start = 0, end = 0;
        // no ':' anywhere: the regex cannot match
        bool colon = false;
        for (uint64_t k = p; k < e; ++k) colon |= text[k] == ':';
        if (!colon) res = -2;
        else {
            uint64_t t0 = p;
            bool plain = true;                              // the token holds \w only (a name may not contain '/' or '-')
            while (p < e && (is_word(text[p]) || text[p] == '/' || text[p] == '-')) { plain &= is_word(text[p]); ++p; }
            bool ok = p > t0;
            if (ok && p < e && text[p] == '.') {            // name.chr
                ok = plain;
                ++p;
                t0 = p;
                while (p < e && (is_word(text[p]) || text[p] == '/' || text[p] == '-')) ++p;
                ok = ok && p > t0;
            }
            const uint64_t c0 = t0, c1 = p;                 // the chr token
            if (ok && p < e && text[p] == '(') {            // strand: (+) or (-)
                ok = p + 2 < e && (text[p + 1] == '+' || text[p + 1] == '-') && text[p + 2] == ')';
                p += 3;
                // the regex's strand group is greedy: a later ')' could pull the match further down the line
                for (uint64_t k = p; k < e && ok; ++k) ok = text[k] != ')';
            }
            ok = ok && p < e && text[p] == ':';
            ++p;
            uint64_t d0 = p, v = 0;
            while (ok && p < e && text[p] >= '0' && text[p] <= '9' && p - d0 < 10) { v = v * 10 + (text[p] - '0'); ++p; }
            ok = ok && p > d0 && !(p < e && text[p] >= '0' && text[p] <= '9') && v <= 0x7FFFFFFFull;   // i32 in intspan
            start = (uint32_t)v;
            if (ok && p < e && (text[p] == '-' || text[p] == '_')) ++p;
            d0 = p, v = 0;
            while (ok && p < e && text[p] >= '0' && text[p] <= '9' && p - d0 < 10) { v = v * 10 + (text[p] - '0'); ++p; }
            ok = ok && !(p < e && text[p] >= '0' && text[p] <= '9') && v <= 0x7FFFFFFFull;
            end = p > d0 ? (uint32_t)v : start;
            if (ok) {
                if (start == 0 || end < start) res = -2;
                else {
                    // chromosome name -> index: FNV-1a into the open-addressing table the host built
                    uint32_t h = 2166136261u;
                    for (uint64_t k = c0; k < c1; ++k) h = (h ^ text[k]) * 16777619u;
                    res = -1;
                    for (uint32_t s = h & hmask;; s = (s + 1) & hmask) {
                        const int32_t c = htab[s];
                        if (c < 0) break;
                        const uint32_t a = name_off[c], len = name_off[c + 1] - a;
                        if (len == (uint32_t)(c1 - c0)) {
                            bool same = true;
                            for (uint32_t k = 0; k < len && same; ++k) same = names[a + k] == text[c0 + k];
                            if (same) { res = c; break; }
                        }
                    }
                }
            }
        }
        q_chr[i] = res;
        q_start[i] = start;
        q_end[i] = end;
    }
}

}  // namespace gams

using namespace gams;

GAMS_API int32_t gams_cuda_parse_ranges(gams_cuda_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int64_t n_chr,
                                        const uint8_t *chr_names, const uint32_t *chr_name_off, uint64_t cap_lines,
                                        uint64_t *n_lines, uint64_t *line_off, int32_t *q_chr, uint32_t *q_start, uint32_t *q_end) {
    ARG_CHECK(ctx && n_lines && n_chr >= 0 && n_chr < (1ll << 30), "parse_ranges: bad argument");
    ARG_CHECK(n_chr == 0 || (chr_names && chr_name_off), "parse_ranges: NULL name table");
    *n_lines = 0;
    if (n_bytes == 0) return GAMS_CUDA_OK;
    ARG_CHECK(text != nullptr, "parse_ranges: text is NULL");
    // the name table: open addressing, FNV-1a, load factor <= 1/2 (host work: n_chr names)
    uint32_t hsize = 16;
    while (hsize < 2 * (uint64_t)n_chr + 2) hsize <<= 1;
    std::vector<int32_t> htab(hsize, -1);
    for (int64_t c = 0; c < n_chr; ++c) {
        ARG_CHECK(chr_name_off[c + 1] >= chr_name_off[c], "parse_ranges: chr_name_off decreases at %lld", (long long)c);
        uint32_t h = 2166136261u;
        for (uint32_t k = chr_name_off[c]; k < chr_name_off[c + 1]; ++k) h = (h ^ chr_names[k]) * 16777619u;
        uint32_t s = h & (hsize - 1);
        while (htab[s] >= 0) s = (s + 1) & (hsize - 1);
        htab[s] = (int32_t)c;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    cudaStream_t st = pd.stream;
    const uint64_t n_tiles = (n_bytes + PR_TILE - 1) / PR_TILE;
    ARG_CHECK(n_tiles < (1ull << 31), "parse_ranges: text too large for one call");
    const uint64_t names_bytes = n_chr ? chr_name_off[n_chr] : 0;
    int32_t rc;
    if ((rc = pd.scratch[0].reserve(n_bytes + 64))) return rc;
    if ((rc = pd.scratch[1].reserve(n_tiles * 12 + 64))) return rc;
    if ((rc = pd.scratch[2].reserve(names_bytes + (size_t)(n_chr + 1) * 4 + (size_t)hsize * 4 + 64))) return rc;
    uint8_t *d_text = pd.scratch[0].as<uint8_t>();
    uint64_t *d_tile_base = pd.scratch[1].as<uint64_t>(), *d_total = d_tile_base + n_tiles;
    uint32_t *d_tile_cnt = reinterpret_cast<uint32_t *>(d_total + 1);
    int32_t *d_htab = pd.scratch[2].as<int32_t>();
    uint32_t *d_noff = reinterpret_cast<uint32_t *>(d_htab + hsize);
    uint8_t *d_names = reinterpret_cast<uint8_t *>(d_noff + (n_chr + 1));
    CUDA_TRY(cudaMemcpyAsync(d_text, text, n_bytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_htab, htab.data(), (size_t)hsize * 4, cudaMemcpyHostToDevice, st));
    if (n_chr) {
        CUDA_TRY(cudaMemcpyAsync(d_noff, chr_name_off, (size_t)(n_chr + 1) * 4, cudaMemcpyHostToDevice, st));
        if (names_bytes) CUDA_TRY(cudaMemcpyAsync(d_names, chr_names, names_bytes, cudaMemcpyHostToDevice, st));
    } else {
        const uint32_t zero = 0;
        CUDA_TRY(cudaMemcpyAsync(d_noff, &zero, 4, cudaMemcpyHostToDevice, st));
    }
    nl_count_kernel<<<(unsigned)n_tiles, PR_THREADS, 0, st>>>(d_text, n_bytes, d_tile_cnt);
    nl_scan_kernel<<<1, 1024, 0, st>>>(d_tile_cnt, n_tiles, d_tile_base, d_total);
    uint64_t n_nl = 0;
    uint8_t last = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_nl, d_total, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // also: htab is a local
    last = text[n_bytes - 1];
    const uint64_t nl = n_nl + (last == '\n' ? 0 : 1);
    *n_lines = nl;
    if (nl > cap_lines) {
        set_error("parse_ranges: %llu lines, room for %llu (call again with larger arrays)", (unsigned long long)nl,
                  (unsigned long long)cap_lines);
        return GAMS_CUDA_ECAP;
    }
    ARG_CHECK(line_off && q_chr && q_start && q_end, "parse_ranges: NULL output");
    if ((rc = pd.scratch[3].reserve((nl + 2) * 8 + nl * 12 + 64))) return rc;
    uint64_t *d_loff = pd.scratch[3].as<uint64_t>();
    int32_t *d_qc = reinterpret_cast<int32_t *>(d_loff + nl + 2);
    uint32_t *d_qs = reinterpret_cast<uint32_t *>(d_qc + nl), *d_qe = d_qs + nl;
    nl_scatter_kernel<<<(unsigned)n_tiles, PR_THREADS, 0, st>>>(d_text, n_bytes, d_tile_base, nl, d_loff);
    const unsigned pb = (unsigned)std::min<uint64_t>((nl + 255) / 256, 148ull * 32);
    parse_lines_kernel<<<pb, 256, 0, st>>>(d_text, n_bytes, d_loff, nl, d_names, d_noff, d_htab, hsize - 1, d_qc, d_qs, d_qe);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(line_off, d_loff, nl * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(q_chr, d_qc, nl * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(q_start, d_qs, nl * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(q_end, d_qe, nl * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    line_off[nl] = n_bytes;
    return GAMS_CUDA_OK;
}
