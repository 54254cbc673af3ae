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
                   uint32_t hmask, int32_t first_field, int32_t *__restrict__ q_chr, uint32_t *__restrict__ q_start,
                   uint32_t *__restrict__ q_end) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += stride) {
        uint64_t p = line_off[i], e = (i + 1 < n_lines) ? line_off[i + 1] - 1 : n_bytes;   // [p, e): without the '\n'
        if (i + 1 == n_lines && e > p && text[e - 1] == '\n') --e;                         // the last line may or may not end in '\n'
        if (e > p && text[e - 1] == '\r') --e;
        if (first_field)                                                                   // locate.rs:87-90: line.split('\t').first()
            for (uint64_t k = p; k < e; ++k)
                if (text[k] == '\t') { e = k; break; }
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

// ---- `locate --count -f` output: "{rg}\t{cnt}\n" per located line (locate.rs:136-139), rg = the first TAB-separated field of
// the input line (locate.rs:87-90) ----------------
__device__ __forceinline__ uint32_t dec_digits(uint32_t v) {
    uint32_t d = 1;
    while (v >= 10u) { v /= 10u; ++d; }
    return d;
}
// [p, e): the first TAB-separated field of line i (the line without its '\n' / '\r\n' when it has no TAB)
__device__ __forceinline__ void line_span(const uint8_t *__restrict__ text, uint64_t n_bytes, const uint64_t *__restrict__ line_off,
                                          uint64_t n_lines, uint64_t i, uint64_t &p, uint64_t &e) {
    p = line_off[i];
    e = (i + 1 < n_lines) ? line_off[i + 1] - 1 : n_bytes;
    if (i + 1 == n_lines && e > p && text[e - 1] == '\n') --e;
    if (e > p && text[e - 1] == '\r') --e;
    for (uint64_t k = p; k < e; ++k)
        if (text[k] == '\t') { e = k; break; }
}
constexpr int FM_THREADS = 256;

// bytes each tile of FM_THREADS lines prints; counters[0] += printed lines, counters[1] += lines handed back (q_chr == -3)
__global__ void __launch_bounds__(FM_THREADS)
fmt_len_kernel(const uint8_t *__restrict__ text, uint64_t n_bytes, const uint64_t *__restrict__ line_off, uint64_t n_lines,
               const int32_t *__restrict__ q_chr, const int32_t *__restrict__ seg, const uint32_t *__restrict__ cnt,
               uint32_t *__restrict__ tile_len, uint32_t *__restrict__ tile_printed, unsigned long long *__restrict__ counters) {
    __shared__ uint32_t ws[FM_THREADS / 32], wp[FM_THREADS / 32];
    const uint64_t i = (uint64_t)blockIdx.x * FM_THREADS + threadIdx.x;
    uint32_t len = 0;
    bool printed = false, back = false;
    if (i < n_lines) {
        back = q_chr[i] == -3;
        if (seg[i] >= 0) {
            uint64_t p, e;
            line_span(text, n_bytes, line_off, n_lines, i, p, e);
            len = (uint32_t)(e - p) + 2u + dec_digits(cnt[i]);
            printed = true;
        }
    }
    const uint32_t np = __popc(__ballot_sync(0xFFFFFFFFu, printed)), nb = __popc(__ballot_sync(0xFFFFFFFFu, back));
    const uint32_t sum = __reduce_add_sync(0xFFFFFFFFu, len);
    if ((threadIdx.x & 31) == 0) {
        ws[threadIdx.x >> 5] = sum;
        wp[threadIdx.x >> 5] = np;
        if (np) atomicAdd(&counters[0], (unsigned long long)np);
        if (nb) atomicAdd(&counters[1], (unsigned long long)nb);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0, tp = 0;
        for (int w = 0; w < FM_THREADS / 32; ++w) t += ws[w], tp += wp[w];
        tile_len[blockIdx.x] = t;
        tile_printed[blockIdx.x] = tp;
    }
}

__global__ void __launch_bounds__(FM_THREADS)
fmt_write_kernel(const uint8_t *__restrict__ text, uint64_t n_bytes, const uint64_t *__restrict__ line_off, uint64_t n_lines,
                 const int32_t *__restrict__ seg, const uint32_t *__restrict__ cnt, const uint64_t *__restrict__ tile_base,
                 const uint64_t *__restrict__ tile_pbase, uint8_t *__restrict__ out, uint64_t *__restrict__ out_line) {
    __shared__ uint32_t ws[FM_THREADS / 32], wp[FM_THREADS / 32];
    const uint64_t i = (uint64_t)blockIdx.x * FM_THREADS + threadIdx.x;
    uint32_t len = 0, c = 0, nd = 0;
    uint64_t p = 0, e = 0;
    if (i < n_lines && seg[i] >= 0) {
        line_span(text, n_bytes, line_off, n_lines, i, p, e);
        c = cnt[i];
        nd = dec_digits(c);
        len = (uint32_t)(e - p) + 2u + nd;
    }
    uint32_t inc = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if ((threadIdx.x & 31) >= d) inc += t;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    const uint32_t pm = __ballot_sync(0xFFFFFFFFu, len != 0);
    if ((threadIdx.x & 31) == 0) wp[threadIdx.x >> 5] = (uint32_t)__popc(pm);
    __syncthreads();
    uint64_t o = tile_base[blockIdx.x] + inc - len;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) o += ws[w];
    if (len && out_line) {   // which input line each printed line came from
        uint64_t r = tile_pbase[blockIdx.x] + (uint32_t)__popc(pm & ((1u << (threadIdx.x & 31)) - 1u));
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) r += wp[w];
        out_line[r] = i;
    }
    if (len) {
        for (uint64_t k = p; k < e; ++k) out[o++] = text[k];
        out[o++] = '\t';
        for (uint32_t k = nd; k-- > 0;) {
            out[o + k] = (uint8_t)('0' + c % 10u);
            c /= 10u;
        }
        out[o + nd] = '\n';
    }
}

}  // namespace gams

using namespace gams;

namespace {
// the text, its line table and the parsed columns, resident on one device
struct ParsedText {
    DevBuf text, tiles, names, lines;
    uint64_t n_lines = 0, n_tiles = 0;
    uint8_t *d_text = nullptr;
    uint64_t *d_loff = nullptr, *d_tile_base = nullptr, *d_total = nullptr;
    int32_t *d_qc = nullptr;
    uint32_t *d_qs = nullptr, *d_qe = nullptr;
    void release() { text.release(); tiles.release(); names.release(); lines.release(); }
};

int32_t parse_on_device(PerDevice &pd, cudaStream_t st, const uint8_t *text, uint64_t n_bytes, int64_t n_chr, const uint8_t *chr_names,
                        const uint32_t *chr_name_off, bool first_field, ParsedText &P) {
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
    const uint64_t n_tiles = (n_bytes + PR_TILE - 1) / PR_TILE;
    ARG_CHECK(n_tiles < (1ull << 31), "parse_ranges: text too large for one call");
    const uint64_t names_bytes = n_chr ? chr_name_off[n_chr] : 0;
    P.text.cache = P.tiles.cache = P.names.cache = P.lines.cache = pd.cache;
    P.n_tiles = n_tiles;
    int32_t rc;
    if ((rc = P.text.reserve(n_bytes + 64))) return rc;
    if ((rc = P.tiles.reserve(n_tiles * 12 + 64))) return rc;
    if ((rc = P.names.reserve(names_bytes + (size_t)(n_chr + 1) * 4 + (size_t)hsize * 4 + 64))) return rc;
    P.d_text = P.text.as<uint8_t>();
    P.d_tile_base = P.tiles.as<uint64_t>();
    P.d_total = P.d_tile_base + n_tiles;
    uint32_t *d_tile_cnt = reinterpret_cast<uint32_t *>(P.d_total + 1);
    int32_t *d_htab = P.names.as<int32_t>();
    uint32_t *d_noff = reinterpret_cast<uint32_t *>(d_htab + hsize);
    uint8_t *d_names = reinterpret_cast<uint8_t *>(d_noff + (n_chr + 1));
    CUDA_TRY(cudaMemcpyAsync(P.d_text, text, n_bytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_htab, htab.data(), (size_t)hsize * 4, cudaMemcpyHostToDevice, st));
    if (n_chr) {
        CUDA_TRY(cudaMemcpyAsync(d_noff, chr_name_off, (size_t)(n_chr + 1) * 4, cudaMemcpyHostToDevice, st));
        if (names_bytes) CUDA_TRY(cudaMemcpyAsync(d_names, chr_names, names_bytes, cudaMemcpyHostToDevice, st));
    } else {
        const uint32_t zero = 0;
        CUDA_TRY(cudaMemcpyAsync(d_noff, &zero, 4, cudaMemcpyHostToDevice, st));
    }
    nl_count_kernel<<<(unsigned)n_tiles, PR_THREADS, 0, st>>>(P.d_text, n_bytes, d_tile_cnt);
    nl_scan_kernel<<<1, 1024, 0, st>>>(d_tile_cnt, n_tiles, P.d_tile_base, P.d_total);
    uint64_t n_nl = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_nl, P.d_total, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // also: htab is a local
    const uint64_t nl = n_nl + (text[n_bytes - 1] == '\n' ? 0 : 1);
    P.n_lines = nl;
    if ((rc = P.lines.reserve((nl + 2) * 8 + nl * 12 + 64))) return rc;
    P.d_loff = P.lines.as<uint64_t>();
    P.d_qc = reinterpret_cast<int32_t *>(P.d_loff + nl + 2);
    P.d_qs = reinterpret_cast<uint32_t *>(P.d_qc + nl);
    P.d_qe = P.d_qs + nl;
    nl_scatter_kernel<<<(unsigned)n_tiles, PR_THREADS, 0, st>>>(P.d_text, n_bytes, P.d_tile_base, nl, P.d_loff);
    const unsigned pb = (unsigned)std::min<uint64_t>((nl + 255) / 256, 148ull * 32);
    parse_lines_kernel<<<pb, 256, 0, st>>>(P.d_text, n_bytes, P.d_loff, nl, d_names, d_noff, d_htab, hsize - 1, first_field ? 1 : 0, P.d_qc,
                                         P.d_qs, P.d_qe);
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}
}  // namespace

GAMS_API int32_t gams_cuda_parse_ranges(gams_cuda_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int64_t n_chr,
                                        const uint8_t *chr_names, const uint32_t *chr_name_off, uint64_t cap_lines,
                                        uint64_t *n_lines, uint64_t *line_off, int32_t *q_chr, uint32_t *q_start, uint32_t *q_end) {
    ARG_CHECK(ctx && n_lines && n_chr >= 0 && n_chr < (1ll << 30), "parse_ranges: bad argument");
    ARG_CHECK(n_chr == 0 || (chr_names && chr_name_off), "parse_ranges: NULL name table");
    *n_lines = 0;
    if (n_bytes == 0) return GAMS_CUDA_OK;
    ARG_CHECK(text != nullptr, "parse_ranges: text is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    cudaStream_t st = pd.stream;
    ParsedText P;
    struct Rel { ParsedText &p; ~Rel() { cudaDeviceSynchronize(); p.release(); } } rel{P};
    int32_t rc = parse_on_device(pd, st, text, n_bytes, n_chr, chr_names, chr_name_off, false, P);
    if (rc) return rc;
    const uint64_t nl = P.n_lines;
    *n_lines = nl;
    if (nl > cap_lines) {
        set_error("parse_ranges: %llu lines, room for %llu (call again with larger arrays)", (unsigned long long)nl,
                  (unsigned long long)cap_lines);
        return GAMS_CUDA_ECAP;
    }
    ARG_CHECK(line_off && q_chr && q_start && q_end, "parse_ranges: NULL output");
    CUDA_TRY(cudaMemcpyAsync(line_off, P.d_loff, nl * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(q_chr, P.d_qc, nl * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(q_start, P.d_qs, nl * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(q_end, P.d_qe, nl * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    line_off[nl] = n_bytes;
    return GAMS_CUDA_OK;
}

// The whole loop of `gams locate --count -f file` (locate.rs:111-141) on the device: text in, text out.
static int32_t locate_count_text_impl(gams_cuda_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int64_t n_chr,
                                      const uint8_t *chr_names, const uint32_t *chr_name_off, const uint32_t *ctg_start,
                                      const uint32_t *ctg_stop, const uint64_t *chr_off, const uint32_t *starts,
                                      const uint32_t *stops, const uint64_t *iv_off, char **out_text, size_t *out_len,
                                      uint64_t **out_line, uint64_t *n_lines, uint64_t *n_printed, uint64_t *n_handed_back) {
    ARG_CHECK(ctx && out_text && out_len && n_chr >= 0 && n_chr < (1ll << 30) && chr_off && iv_off, "locate_count_text: bad argument");
    ARG_CHECK(n_chr == 0 || (chr_names && chr_name_off), "locate_count_text: NULL name table");
    *out_text = nullptr;
    *out_len = 0;
    if (out_line) *out_line = nullptr;
    if (n_lines) *n_lines = 0;
    if (n_printed) *n_printed = 0;
    if (n_handed_back) *n_handed_back = 0;
    if (n_bytes == 0) return GAMS_CUDA_OK;
    ARG_CHECK(text != nullptr, "locate_count_text: text is NULL");
    ARG_CHECK(chr_off[0] == 0 && iv_off[0] == 0, "locate_count_text: chr_off[0] and iv_off[0] must be 0");
    const uint64_t n_ctg = chr_off[n_chr], n_iv = iv_off[n_ctg];
    ARG_CHECK(n_ctg < (1ull << 31), "locate_count_text: too many ctgs");
    ARG_CHECK((ctg_start && ctg_stop) || n_ctg == 0, "locate_count_text: NULL ctg index");
    ARG_CHECK((starts && stops) || n_iv == 0, "locate_count_text: NULL rg index");
    std::vector<uint32_t> maxlen((size_t)std::max<int64_t>(n_chr, 1), 0);  // Lapper::new: max(stop - start)
    for (int64_t c = 0; c < n_chr; ++c)
        for (uint64_t k = chr_off[c]; k < chr_off[c + 1]; ++k)
            if (ctg_stop[k] > ctg_start[k]) maxlen[c] = std::max(maxlen[c], ctg_stop[k] - ctg_start[k]);
    // the lock first: everything below allocates from the device's block cache; the plan goes before the lock is released
    std::lock_guard<std::mutex> lk(ctx->mu);
    gams_cuda_count_plan *plan = nullptr;
    int32_t rc = gams_cuda_count_plan_create(ctx, 0, (int64_t)n_ctg, iv_off, &plan);
    if (rc) return rc;
    std::unique_ptr<gams_cuda_count_plan, void (*)(gams_cuda_count_plan *)> pguard(plan, gams_cuda_count_plan_destroy);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    cudaStream_t st = pd.stream;
    ParsedText P;
    DevBuf idx, work, outb;
    idx.cache = work.cache = outb.cache = pd.cache;
    struct Rel { ParsedText &p; DevBuf &a, &b, &c; ~Rel() { cudaDeviceSynchronize(); p.release(); a.release(); b.release(); c.release(); } } rel{P, idx, work, outb};
    if ((rc = parse_on_device(pd, st, text, n_bytes, n_chr, chr_names, chr_name_off, true, P))) return rc;
    const uint64_t nl = P.n_lines;
    if (n_lines) *n_lines = nl;
    // the two indexes
    const uint64_t nc = std::max<uint64_t>(n_ctg, 1), ni = std::max<uint64_t>(n_iv, 1);
    if ((rc = idx.reserve((uint64_t)(n_chr + 1) * 8 + nc * 8 + (uint64_t)std::max<int64_t>(n_chr, 1) * 4 + ni * 8 + 64))) return rc;
    uint64_t *d_chr_off = idx.as<uint64_t>();
    uint32_t *d_cs = reinterpret_cast<uint32_t *>(d_chr_off + (n_chr + 1)), *d_ce = d_cs + nc, *d_ml = d_ce + nc;
    uint32_t *d_st = d_ml + std::max<int64_t>(n_chr, 1), *d_sp = d_st + ni;
    CUDA_TRY(cudaMemcpyAsync(d_chr_off, chr_off, (uint64_t)(n_chr + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_ctg) {
        CUDA_TRY(cudaMemcpyAsync(d_cs, ctg_start, n_ctg * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_ce, ctg_stop, n_ctg * 4, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(d_ml, maxlen.data(), maxlen.size() * 4, cudaMemcpyHostToDevice, st));
    if (n_iv) {
        CUDA_TRY(cudaMemcpyAsync(d_st, starts, n_iv * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_sp, stops, n_iv * 4, cudaMemcpyHostToDevice, st));
    }
    // locate -> count -> lengths -> offsets -> bytes
    const uint64_t n_ft = (nl + FM_THREADS - 1) / FM_THREADS;
    if ((rc = work.reserve(nl * 8 + 8 + n_ft * 24 + 48 + 64))) return rc;
    int32_t *d_seg = work.as<int32_t>();
    uint32_t *d_cnt = reinterpret_cast<uint32_t *>(d_seg + nl);
    uint64_t *d_fbase = reinterpret_cast<uint64_t *>(d_cnt + nl), *d_ftotal = d_fbase + n_ft;   // 8 nl bytes in: aligned
    uint64_t *d_pbase = d_ftotal + 1, *d_ptotal = d_pbase + n_ft;
    unsigned long long *d_counters = reinterpret_cast<unsigned long long *>(d_ptotal + 1);
    uint32_t *d_flen = reinterpret_cast<uint32_t *>(d_counters + 2), *d_fprinted = d_flen + n_ft;
    CUDA_TRY(cudaMemsetAsync(d_counters, 0, 16, st));
    if ((rc = launch_locate_ctg(st, n_chr, d_cs, d_ce, d_chr_off, d_ml, (int64_t)nl, P.d_qc, P.d_qs, P.d_qe, d_seg))) return rc;
    if ((rc = launch_count_overlaps(plan, st, d_st, d_sp, (int64_t)nl, d_seg, P.d_qs, P.d_qe, 0, d_cnt))) return rc;
    fmt_len_kernel<<<(unsigned)n_ft, FM_THREADS, 0, st>>>(P.d_text, n_bytes, P.d_loff, nl, P.d_qc, d_seg, d_cnt, d_flen, d_fprinted, d_counters);
    nl_scan_kernel<<<1, 1024, 0, st>>>(d_flen, n_ft, d_fbase, d_ftotal);
    if (out_line) nl_scan_kernel<<<1, 1024, 0, st>>>(d_fprinted, n_ft, d_pbase, d_ptotal);
    uint64_t h[3] = {0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(&h[0], d_ftotal, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&h[1], d_counters, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (n_printed) *n_printed = h[1];
    if (n_handed_back) *n_handed_back = h[2];
    const uint64_t total = h[0], n_pr = h[1];
    char *host = static_cast<char *>(malloc(total + 1));
    uint64_t *host_line = out_line ? static_cast<uint64_t *>(malloc((n_pr + 1) * sizeof(uint64_t))) : nullptr;
    if (!host || (out_line && !host_line)) {
        free(host);
        free(host_line);
        set_error("locate_count_text: cannot allocate %llu bytes for the output", (unsigned long long)total);
        return GAMS_CUDA_ENOMEM;
    }
    if (total) {
        if ((rc = outb.reserve(total + 64 + (out_line ? n_pr * 8 + 8 : 0)))) { free(host); free(host_line); return rc; }
        uint64_t *d_line = out_line ? reinterpret_cast<uint64_t *>(outb.as<uint8_t>() + ((total + 15) & ~15ull)) : nullptr;
        fmt_write_kernel<<<(unsigned)n_ft, FM_THREADS, 0, st>>>(P.d_text, n_bytes, P.d_loff, nl, d_seg, d_cnt, d_fbase, d_pbase, outb.as<uint8_t>(),
                                                                d_line);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(host, outb.p, total, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && out_line && n_pr) e = cudaMemcpyAsync(host_line, d_line, n_pr * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { free(host); free(host_line); return cuda_fail(e, "locate_count_text output", __FILE__, __LINE__); }
    }
    host[total] = 0;
    *out_text = host;
    *out_len = (size_t)total;
    if (out_line) *out_line = host_line;
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_locate_count_text(gams_cuda_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int64_t n_chr,
                                             const uint8_t *chr_names, const uint32_t *chr_name_off, const uint32_t *ctg_start,
                                             const uint32_t *ctg_stop, const uint64_t *chr_off, const uint32_t *starts,
                                             const uint32_t *stops, const uint64_t *iv_off, char **out_text, size_t *out_len,
                                             uint64_t *n_lines, uint64_t *n_printed, uint64_t *n_handed_back) {
    return locate_count_text_impl(ctx, text, n_bytes, n_chr, chr_names, chr_name_off, ctg_start, ctg_stop, chr_off, starts, stops, iv_off,
                                  out_text, out_len, nullptr, n_lines, n_printed, n_handed_back);
}

GAMS_API int32_t gams_cuda_locate_count_lines(gams_cuda_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int64_t n_chr,
                                              const uint8_t *chr_names, const uint32_t *chr_name_off, const uint32_t *ctg_start,
                                              const uint32_t *ctg_stop, const uint64_t *chr_off, const uint32_t *starts,
                                              const uint32_t *stops, const uint64_t *iv_off, char **out_text, size_t *out_len,
                                              uint64_t **out_line, uint64_t *n_lines, uint64_t *n_printed, uint64_t *n_handed_back) {
    ARG_CHECK(out_line != nullptr, "locate_count_lines: out_line is NULL");
    return locate_count_text_impl(ctx, text, n_bytes, n_chr, chr_names, chr_name_off, ctg_start, ctg_stop, chr_off, starts, stops, iv_off,
                                  out_text, out_len, out_line, n_lines, n_printed, n_handed_back);
}
