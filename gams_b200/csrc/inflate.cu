// inflate.cu — gunzip of a batch of `seq:{ctg}` blobs on the device.
//
// Replaces the feed step of every compute subcommand: Conn::get_seq -> decode_gz (src/libs/redis.rs:142-161):
//     GET seq:{ctg}  ->  flate2::read::GzDecoder::read_to_end  ->  String::from_utf8
// The blobs are gzip members (RFC 1952) written by encode_gz with flate2 Compression::fast() (redis.rs:149-154); the
// algorithm lives in flate2 1.0.30 / miniz_oxide (not vendored in the reference) and is restated here from RFC 1951:
// stored, fixed-Huffman and dynamic-Huffman blocks, 32 KB window.  Like GzDecoder, only the first member of a blob is
// read; the ISIZE trailer is checked, the CRC-32 is not (documented in include/gams_cuda.h).
//
// One WARP per blob (a deflate stream is bit-serial; the batch is the parallelism: one human genome is 6 200 streams).
// Every lane keeps the same bit reader and decodes the same symbol (broadcast loads, no divergence, no shuffles): one
// primary-table lookup per alphabet in shared memory (9 / 7 index bits; an entry carries code length, extra-bit count and
// base, so a length or distance is one load, one shift and one add), canonical bit-by-bit decoding for the rare longer
// codes.  The lanes differ only in the copies: the 32 lanes copy a match together (level-1 DNA is almost all matches of
// ~4 bases); stored blocks and table builds are cooperative as well.  History is read back from the output buffer itself
// (L1/L2), no window is staged.
// Bound: instruction issue (one symbol = one dependent chain), not HBM.
#include "common.cuh"
#include "inflate_common.cuh"

namespace gams {

constexpr int INF_WARPS = 4;        // per CTA: small CTAs free their slots soon (streams of a batch end at different times)
constexpr int INF_LBITS = 9;        // primary table index bits, literal/length alphabet
constexpr int INF_DBITS = 7;        // primary table index bits, distance alphabet

// Table entry (both alphabets): bits 0-3 code length (0: the code is longer than the primary index -> canonical decoding),
// bits 4-7 number of extra bits, bits 8-9 kind (0 literal, 1 length, 2 end of block, 3 invalid), bits 16-31 literal / base.
constexpr uint32_t INF_KLIT = 0u << 8, INF_KLEN = 1u << 8, INF_KEOB = 2u << 8, INF_KBAD = 3u << 8;

struct InfTables {                  // per warp, shared memory (3.4 KB: 48 warps per SM)
    uint32_t ltab[1 << INF_LBITS];
    uint32_t dtab[1 << INF_DBITS];
    uint16_t lcnt[16], dcnt[16];    // codes per length
    uint16_t lsym[288], dsym[32];   // symbols in canonical order
    uint8_t lens[320];              // code lengths of the block being set up
    uint16_t ccnt[8];               // the code-length alphabet of a dynamic header
    uint8_t csym[20], clen[20];
};

__constant__ uint16_t INF_LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t INF_LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t INF_DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t INF_DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t INF_CLORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__device__ __forceinline__ uint32_t inf_lit_entry(int s) {     // without the code length
    if (s < 256) return INF_KLIT | ((uint32_t)s << 16);
    if (s == 256) return INF_KEOB;
    if (s > 285) return INF_KBAD;
    return INF_KLEN | ((uint32_t)INF_LEXT[s - 257] << 4) | ((uint32_t)INF_LBASE[s - 257] << 16);
}
__device__ __forceinline__ uint32_t inf_dist_entry(int s) {
    if (s > 29) return INF_KBAD;
    return ((uint32_t)INF_DEXT[s] << 4) | ((uint32_t)INF_DBASE[s] << 16);
}

// The bit reader: two consecutive aligned 32-bit words of the blob and a bit offset into the first.  EVERY lane of the
// warp keeps an identical copy and decodes the same symbols (same addresses: broadcast loads), so the decode loop has no
// divergence, no shuffles and no predication; only the copies use the lanes differently.
struct BitReader {
    const uint32_t *w;              // the blob's first byte lies `mis` bytes into w[0]
    uint32_t n_words;               // words that hold blob bytes; zeros are read beyond
    uint32_t mis;
    uint32_t w0, w1, o, wi;         // w0 = w[wi - 2], w1 = w[wi - 1], o < 32 bits of w0 are consumed
    __device__ __forceinline__ uint32_t ld(uint32_t i) const { return i < n_words ? __ldg(w + i) : 0u; }
    __device__ __forceinline__ void seek(uint32_t byte) {
        const uint32_t a = mis + byte;
        wi = (a >> 2) + 2;
        o = (a & 3u) * 8u;
        w0 = ld(wi - 2);
        w1 = ld(wi - 1);
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(w0, w1, o); }   // the next 32 bits
    __device__ __forceinline__ void drop(uint32_t n) {                                        // n <= 32
        o += n;
        if (o >= 32u) {
            w0 = w1;
            w1 = ld(wi);
            ++wi;
            o -= 32u;
        }
    }
    __device__ __forceinline__ uint32_t bits(uint32_t n) { const uint32_t v = peek() & ((1u << n) - 1u); drop(n); return v; }   // n < 32
    __device__ __forceinline__ uint64_t consumed_bits() const { return 32ull * (wi - 2) + o - 8ull * mis; }
};

// bits [pos, pos + len) of x, len <= 15 (BFE)
__device__ __forceinline__ uint32_t inf_bfe(uint32_t x, uint32_t pos, uint32_t len) {
    uint32_t r;
    asm("bfe.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(pos), "r"(len));
    return r;
}

// canonical decoding of a code longer than the primary index (or of the code-length alphabet), one bit at a time from the
// peeked 32 bits; returns the symbol (or -1) and the code length
__device__ __forceinline__ int inf_decode_slow(uint32_t pk, const uint16_t *cnt, const uint16_t *sym, int max_sym, uint32_t *code_len) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= (int)((pk >> (len - 1)) & 1u);
        const int c = cnt[len];
        if (code - c < first) {
            const int k = index + (code - first);
            *code_len = (uint32_t)len;
            return k < max_sym ? (int)sym[k] : -1;
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    *code_len = 15;
    return -1;
}

// Warp-cooperative table build from code lengths lens[0 .. n): lane L (1..15) owns the codes of length L.
// Returns false for an over-subscribed set of lengths.
template <int PBITS, bool DIST>
__device__ bool inf_build(const uint8_t *lens, int n, uint16_t *cnt, uint16_t *sym, uint32_t *tab, int lane) {
    for (int i = lane; i < (1 << PBITS); i += 32) tab[i] = 0;
    int mine = 0;
    if (lane >= 1 && lane <= 15)
        for (int s = 0; s < n; ++s) mine += (lens[s] == lane) ? 1 : 0;
    if (lane < 16) cnt[lane] = (uint16_t)(lane ? mine : 0);
    __syncwarp();
    bool ok = true;
    if (lane >= 1 && lane <= 15) {
        int offs = 0, code = 0, left = 1;
        for (int l = 1; l <= 15; ++l) {          // Kraft check over all lengths (every lane computes it: tiny)
            left <<= 1;
            left -= cnt[l];
            if (left < 0) ok = false;
        }
        for (int l = 1; l < lane; ++l) {
            offs += cnt[l];
            code = (code + cnt[l]) << 1;
        }
        // code = first canonical code of length `lane`; symbols of this length in increasing order
        int k = 0;
        for (int s = 0; s < n && ok; ++s) {
            if (lens[s] != lane) continue;
            sym[offs + k] = (uint16_t)s;
            if (lane <= PBITS) {
                const uint32_t rev = __brev((uint32_t)(code + k)) >> (32 - lane);
                const uint32_t e = (DIST ? inf_dist_entry(s) : inf_lit_entry(s)) | (uint32_t)lane;
                for (uint32_t j = rev; j < (1u << PBITS); j += (1u << lane)) tab[j] = e;
            }
            ++k;
        }
    }
    ok = __all_sync(0xFFFFFFFFu, ok);
    __syncwarp();
    return ok;
}


__global__ void __launch_bounds__(INF_WARPS * 32, 12)
gunzip_kernel(const uint8_t *__restrict__ gz, const uint64_t *__restrict__ gz_off, int64_t n_blob, uint8_t *out_base,
              const uint64_t *__restrict__ out_off, int32_t *__restrict__ status, int32_t only_status,
              uint32_t *__restrict__ crc_want) {
    __shared__ InfTables tabs[INF_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t blob = (int64_t)blockIdx.x * INF_WARPS + wib;
    if (blob >= n_blob) return;
    if (only_status >= 0 && status[blob] != only_status) return;   // (warp-uniform)
    InfTables &T = tabs[wib];
    uint8_t *out = out_base + out_off[blob];
    const uint64_t cap64 = out_off[blob + 1] - out_off[blob], len64 = gz_off[blob + 1] - gz_off[blob];
    const uint8_t *in = gz + gz_off[blob];
    int err = INF_OK;
    if (cap64 > 0xFFFFFF00ull || len64 > 0xFFFFFF00ull) err = INF_ESIZE;     // positions are 32-bit
    const uint32_t cap = (uint32_t)cap64, in_len = (uint32_t)len64;
    BitReader br;
    br.mis = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 3u);
    br.w = reinterpret_cast<const uint32_t *>(in - br.mis);
    br.n_words = (br.mis + in_len + 3u) >> 2;
    // ---- gzip header (RFC 1952)
    uint32_t p = 10;
    if (!err) {
        if (in_len < 18 || in[0] != 0x1F || in[1] != 0x8B || in[2] != 8) err = INF_EHEADER;
        else {
            const uint32_t flg = in[3];
            if (flg & 4u) p += 2 + (p + 2 <= in_len ? (uint32_t)in[p] | ((uint32_t)in[p + 1] << 8) : 0u);   // FEXTRA
            if (flg & 8u) { while (p < in_len && in[p]) ++p; ++p; }                                           // FNAME
            if (flg & 16u) { while (p < in_len && in[p]) ++p; ++p; }                                          // FCOMMENT
            if (flg & 2u) p += 2;                                                                             // FHCRC
            if (p + 8 > in_len) err = INF_EHEADER;
        }
    }
    br.seek(err ? 0u : p);
    uint32_t pos = 0;               // bytes written; the same in every lane
    bool last = false;
    while (!err && !last) {
        const uint32_t hdr = br.bits(3);
        last = hdr & 1u;
        const uint32_t type = hdr >> 1;
        if (type == 0) {
            // stored: skip to the byte boundary, LEN / NLEN, copy
            br.drop((8u - (br.o & 7u)) & 7u);
            const uint32_t ln = br.peek();
            br.drop(32);
            const uint32_t len = ln & 0xFFFFu;
            if ((len ^ 0xFFFFu) != (ln >> 16)) { err = INF_ETRUNC; break; }
            const uint32_t src = (uint32_t)(br.consumed_bits() >> 3);
            if ((uint64_t)src + len > in_len) { err = INF_ETRUNC; break; }
            if ((uint64_t)pos + len > cap) { err = INF_EOVERFLOW; break; }
            for (uint32_t i = lane; i < len; i += 32) out[pos + i] = __ldg(in + src + i);
            pos += len;
            br.seek(src + len);
            continue;
        }
        if (type == 3) { err = INF_ECODE; break; }
        int nlen = 288, ndist = 30;
        if (type == 1) {
            for (int s = lane; s < 288; s += 32) T.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
            if (lane < 30) T.lens[288 + lane] = 5;
        } else {
            // dynamic: code-length code, then HLIT + HDIST lengths.  Sequential (~300 short symbols per block); every lane
            // runs it and writes the same values.
            const int hlit = (int)br.bits(5) + 257, hdist = (int)br.bits(5) + 1, hclen = (int)br.bits(4) + 4;
            bool bad = hlit > 286 || hdist > 30;
            if (lane < 19) T.clen[lane] = 0;
            if (lane < 8) T.ccnt[lane] = 0;
            __syncwarp();
            for (int i = 0; i < hclen; ++i) {
                const uint32_t v = br.bits(3);
                if (lane == 0) T.clen[INF_CLORDER[i]] = (uint8_t)v;
            }
            __syncwarp();
            if (lane == 0) {
                uint32_t cc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int i = 0; i < 19; ++i) cc[T.clen[i]]++;
                cc[0] = 0;
                uint32_t offs[8];
                offs[1] = 0;
                for (int l = 1; l < 7; ++l) offs[l + 1] = offs[l] + cc[l];
                for (int l = 0; l < 8; ++l) T.ccnt[l] = (uint16_t)cc[l];
                for (int i = 0; i < 19; ++i)
                    if (T.clen[i]) T.csym[offs[T.clen[i]]++] = (uint8_t)i;
            }
            __syncwarp();
            int idx = 0;
            const int total = hlit + hdist;
            int prev = 0;
            while (idx < total && !bad) {
                // canonical decoding over the (at most 7-bit) code-length codes
                const uint32_t pk = br.peek();
                int code = 0, first = 0, index = 0, s = -1;
                uint32_t cl = 7;
                for (int len = 1; len <= 7; ++len) {
                    code |= (int)((pk >> (len - 1)) & 1u);
                    const int c = T.ccnt[len];
                    if (code - c < first) { s = T.csym[index + (code - first)]; cl = (uint32_t)len; break; }
                    index += c; first += c; first <<= 1; code <<= 1;
                }
                if (s < 0) { bad = true; break; }
                br.drop(cl);
                if (s < 16) {
                    if (lane == 0) T.lens[idx] = (uint8_t)s;
                    prev = s;
                    ++idx;
                } else {
                    int rep, v = 0;
                    if (s == 16) {
                        if (idx == 0) { bad = true; break; }
                        v = prev;
                        rep = 3 + (int)br.bits(2);
                    } else if (s == 17) rep = 3 + (int)br.bits(3);
                    else rep = 11 + (int)br.bits(7);
                    if (idx + rep > total) { bad = true; break; }
                    if (lane == 0)
                        for (int k = 0; k < rep; ++k) T.lens[idx + k] = (uint8_t)v;
                    idx += rep;
                    prev = v;
                }
            }
            __syncwarp();
            if (!bad && T.lens[256] == 0) bad = true;      // no end-of-block code
            if (bad) { err = INF_ECODE; break; }
            nlen = hlit;
            ndist = hdist;
        }
        __syncwarp();
        const bool ok1 = inf_build<INF_LBITS, false>(T.lens, nlen, T.lcnt, T.lsym, T.ltab, lane);
        const bool ok2 = inf_build<INF_DBITS, true>(T.lens + (type == 1 ? 288 : nlen), ndist, T.dcnt, T.dsym, T.dtab, lane);
        if (!ok1 || !ok2) { err = INF_ECODE; break; }
        // ---- symbols
        uint8_t *const out_lane = out + lane;
        for (;;) {
            uint32_t pk = br.peek();
            uint32_t e = T.ltab[pk & ((1u << INF_LBITS) - 1u)];
            if ((e & 15u) == 0u) {
                uint32_t cl;
                const int s = inf_decode_slow(pk, T.lcnt, T.lsym, nlen, &cl);
                e = (s < 0 ? INF_KBAD : inf_lit_entry(s)) | cl;
            }
            const uint32_t cl = e & 15u, ex = (e >> 4) & 15u;
            const uint32_t val = (e >> 16) + inf_bfe(pk, cl, ex);
            br.drop(cl + ex);
            if (e & INF_KLEN) {         // a length (kind 1) or an invalid code (kind 3)
                if (e & (2u << 8)) { err = INF_ECODE; break; }
                pk = br.peek();
                uint32_t d = T.dtab[pk & ((1u << INF_DBITS) - 1u)];
                if ((d & 15u) == 0u) {
                    uint32_t dl;
                    const int s = inf_decode_slow(pk, T.dcnt, T.dsym, ndist, &dl);
                    d = (s < 0 ? INF_KBAD : inf_dist_entry(s)) | dl;
                }
                const uint32_t dl = d & 15u, dx = (d >> 4) & 15u;
                const uint32_t dist = (d >> 16) + inf_bfe(pk, dl, dx);
                br.drop(dl + dx);
                const uint32_t len = val;
                if ((d & (3u << 8)) != 0u) { err = INF_ECODE; break; }
                if (dist > pos) { err = INF_EDIST; break; }
                if (len > cap - pos) { err = INF_EOVERFLOW; break; }
                __syncwarp();           // the bytes written so far (lane 0's literals, every lane's copies) are visible
                uint8_t *dst = out_lane + pos;
                const uint8_t *src = dst - dist;
                if (dist >= len) {
                    if (lane < len) dst[0] = src[0];
                    if (len > 32u)
                        for (uint32_t i = 32; i + lane < len; i += 32) dst[i] = src[i];
                } else if (dist == 1) {
                    const uint8_t v = out[pos - 1];
                    for (uint32_t i = lane; i < len; i += 32) out[pos + i] = v;
                } else {
                    for (uint32_t i = lane; i < len; i += 32) out[pos + i] = out[pos - dist + i % dist];   // the period repeats
                }
                pos += len;
                continue;
            }
            if (e & (2u << 8)) break;   // end of block
            if (pos >= cap) { err = INF_EOVERFLOW; break; }
            if (lane == 0) out[pos] = (uint8_t)val;
            ++pos;
        }
    }
    // ---- trailer: ISIZE (RFC 1952: size of the original input modulo 2^32) and the size the caller expects
    if (lane == 0) {
        if (!err) {
            const uint64_t t = (br.consumed_bits() + 7) >> 3;      // first byte after the deflate stream = CRC32
            if (t + 8 > in_len) err = INF_ETRUNC;
            else {
                const uint8_t *q = in + t + 4;
                const uint32_t isize = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
                if (isize != pos || pos != cap) err = INF_ESIZE;
                if (crc_want) crc_want[blob] = (uint32_t)q[-4] | ((uint32_t)q[-3] << 8) | ((uint32_t)q[-2] << 16) | ((uint32_t)q[-1] << 24);
            }
        }
        status[blob] = err;
    }
}

int32_t launch_gunzip_only(cudaStream_t st, const uint8_t *d_gz, const uint64_t *d_gz_off, int64_t n_blob, uint8_t *d_out,
                           const uint64_t *d_out_off, int32_t *d_status, int32_t only_status, uint32_t *d_crc_want) {
    if (n_blob == 0) return GAMS_CUDA_OK;
    const int64_t blocks = (n_blob + INF_WARPS - 1) / INF_WARPS;
    if (blocks >= (1ll << 31)) {
        set_error("gunzip: too many blobs for one launch");
        return GAMS_CUDA_EINVAL;
    }
    gunzip_kernel<<<(unsigned)blocks, INF_WARPS * 32, 0, st>>>(d_gz, d_gz_off, n_blob, d_out, d_out_off, d_status, only_status, d_crc_want);
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams

using namespace gams;

// ISIZE of every blob (host): the uncompressed length modulo 2^32, from the gzip trailer
GAMS_API int32_t gams_cuda_gz_sizes(int64_t n_blob, const uint8_t *gz_concat, const uint64_t *gz_off, uint64_t *out_len) {
    ARG_CHECK(n_blob >= 0 && gz_off && out_len && (gz_concat || n_blob == 0), "gz_sizes: NULL argument");
    for (int64_t i = 0; i < n_blob; ++i) {
        ARG_CHECK(gz_off[i + 1] >= gz_off[i] + 18, "gz_sizes: blob %lld is shorter than a gzip member", (long long)i);
        const uint8_t *b = gz_concat + gz_off[i], *t = gz_concat + gz_off[i + 1] - 4;
        ARG_CHECK(b[0] == 0x1F && b[1] == 0x8B && b[2] == 8, "gz_sizes: blob %lld is not gzip/deflate", (long long)i);
        out_len[i] = (uint64_t)t[0] | ((uint64_t)t[1] << 8) | ((uint64_t)t[2] << 16) | ((uint64_t)t[3] << 24);
    }
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_gunzip_dev(gams_cuda_ctx *ctx, int32_t dev_index, void *stream, int64_t n_blob, const uint8_t *d_gz,
                                      const uint64_t *d_gz_off, uint8_t *d_out, const uint64_t *d_out_off, int32_t *d_status) {
    ARG_CHECK(ctx && dev_index >= 0 && dev_index < (int32_t)ctx->devs.size(), "gunzip_dev: bad context / device index");
    ARG_CHECK(n_blob >= 0 && (n_blob == 0 || (d_gz && d_gz_off && d_out && d_out_off && d_status)), "gunzip_dev: NULL device pointer");
    if (n_blob == 0) return GAMS_CUDA_OK;
    PerDevice &pd = ctx->devs[dev_index];
    DeviceGuard g(pd.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the token buffer of the two-phase decoder is sized from the total output: two offsets come back (the one sync of this call)
    uint64_t ends[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(&ends[0], d_out_off, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&ends[1], d_out_off + n_blob, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::lock_guard<std::mutex> lk(ctx->mu);   // the decoder's scratch belongs to the context
    return launch_gunzip2(pd, st, d_gz, d_gz_off, n_blob, d_out, d_out_off, d_status, ends[1] - ends[0], ctx->gunzip_mode, ctx->gunzip_crc != 0);
}

static int32_t gunzip_one(gams_cuda_ctx *ctx, int d, int64_t n_blob, const uint8_t *gz_concat, const uint64_t *gz_off,
                          const uint64_t *out_off, uint8_t *out_concat, int32_t *status) {
    if (n_blob == 0) return GAMS_CUDA_OK;
    const uint64_t n_in = gz_off[n_blob] - gz_off[0], n_out = out_off[n_blob] - out_off[0];
    std::vector<uint64_t> go((size_t)n_blob + 1), oo((size_t)n_blob + 1);
    for (int64_t i = 0; i <= n_blob; ++i) {
        go[i] = gz_off[i] - gz_off[0];
        oo[i] = out_off[i] - out_off[0];
    }
    PerDevice &pd = ctx->devs[d];
    DeviceGuard g(pd.device);
    cudaStream_t st = pd.stream;
    int32_t rc;
    if ((rc = pd.scratch[0].reserve(n_in + 64))) return rc;
    if ((rc = pd.scratch[1].reserve(std::max<uint64_t>(n_out, 1) + 64))) return rc;
    if ((rc = pd.scratch[2].reserve((size_t)(n_blob + 1) * 16 + (size_t)n_blob * 4 + 64))) return rc;
    uint8_t *d_gz = pd.scratch[0].as<uint8_t>(), *d_out = pd.scratch[1].as<uint8_t>();
    uint64_t *d_go = pd.scratch[2].as<uint64_t>(), *d_oo = d_go + (n_blob + 1);
    int32_t *d_status = reinterpret_cast<int32_t *>(d_oo + (n_blob + 1));
    // Large batches go in pieces (two by default, GAMS_CUDA_GZ_PIECES): every piece has its own upload, the decoder takes them one
    // after the other, and the bytes of a piece go down while the next is decoded — the output is three times the input, so the
    // download is the longest leg.  (More pieces do not pay: a piece costs the decoder its serial chains, whatever its size.)
    const char *pe = getenv("GAMS_CUDA_GZ_PIECES");
    int64_t n_piece = pe ? std::max<long>(1, std::min<long>(atol(pe), 16)) : std::min<int64_t>(2, (int64_t)(n_out >> 27));
    n_piece = std::max<int64_t>(1, std::min<int64_t>(n_piece, n_blob));
    if (n_piece == 1) {
        CUDA_TRY(cudaMemcpyAsync(d_gz, gz_concat + gz_off[0], n_in, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_go, go.data(), (size_t)(n_blob + 1) * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_oo, oo.data(), (size_t)(n_blob + 1) * 8, cudaMemcpyHostToDevice, st));
        if ((rc = launch_gunzip2(pd, st, d_gz, d_go, n_blob, d_out, d_oo, d_status, n_out, ctx->gunzip_mode, ctx->gunzip_crc != 0))) return rc;
        if (n_out) CUDA_TRY(cudaMemcpyAsync(out_concat + out_off[0], d_out, n_out, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(status, d_status, (size_t)n_blob * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));   // go / oo are locals
        return GAMS_CUDA_OK;
    }
    std::vector<int64_t> cut((size_t)n_piece + 1, 0);      // pieces of about equal output
    for (int64_t q = 1, i = 0; q < n_piece; ++q) {
        const uint64_t want = n_out / (uint64_t)n_piece * (uint64_t)q;
        while (i < n_blob && oo[i] < want) ++i;
        cut[q] = std::max(i, cut[q - 1]);
    }
    cut[n_piece] = n_blob;
    struct Guard {      // an early return must not leave copies from / to the caller's buffers in flight
        PerDevice &pd;
        std::vector<cudaEvent_t> ev;
        ~Guard() {
            cudaStreamSynchronize(pd.copy_in);
            cudaStreamSynchronize(pd.stream);
            cudaStreamSynchronize(pd.copy_out);
            for (auto e : ev) cudaEventDestroy(e);
        }
    } guard{pd, {}};
    guard.ev.assign((size_t)n_piece * 2, nullptr);
    for (auto &e : guard.ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    size_t by = 0;
    if ((rc = gunzip2_scratch_bytes(pd, n_blob, n_out, ctx->gunzip_mode, ctx->gunzip_crc != 0, &by))) return rc;
    if ((rc = pd.scratch[7].reserve(by))) return rc;       // no piece needs more: none re-allocates it under an earlier one
    CUDA_TRY(cudaMemcpyAsync(d_go, go.data(), (size_t)(n_blob + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_oo, oo.data(), (size_t)(n_blob + 1) * 8, cudaMemcpyHostToDevice, st));
    for (int64_t q = 0; q < n_piece; ++q) {
        const int64_t a = cut[q], b = cut[q + 1];
        if (go[b] > go[a]) CUDA_TRY(cudaMemcpyAsync(d_gz + go[a], gz_concat + gz_off[0] + go[a], go[b] - go[a], cudaMemcpyHostToDevice, pd.copy_in));
        CUDA_TRY(cudaEventRecord(guard.ev[2 * q], pd.copy_in));
    }
    for (int64_t q = 0; q < n_piece; ++q) {
        const int64_t a = cut[q], b = cut[q + 1], all = b - a;
        if (all == 0) continue;
        cudaEvent_t up = guard.ev[2 * q], done = guard.ev[2 * q + 1];
        if ((rc = launch_gunzip2(pd, st, d_gz, d_go + a, all, d_out, d_oo + a, d_status + a, oo[b] - oo[a], ctx->gunzip_mode, ctx->gunzip_crc != 0,
                                 1, &all, &up, done)))
            return rc;
        CUDA_TRY(cudaStreamWaitEvent(pd.copy_out, done, 0));
        if (oo[b] > oo[a]) CUDA_TRY(cudaMemcpyAsync(out_concat + out_off[0] + oo[a], d_out + oo[a], oo[b] - oo[a], cudaMemcpyDeviceToHost, pd.copy_out));
    }
    CUDA_TRY(cudaMemcpyAsync(status, d_status, (size_t)n_blob * 4, cudaMemcpyDeviceToHost, st));   // behind the last CRC pass
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaStreamSynchronize(pd.copy_out));
    return GAMS_CUDA_OK;
}

// Host entry point: blobs are split over the context's devices by compressed bytes.
GAMS_API int32_t gams_cuda_gunzip_batch(gams_cuda_ctx *ctx, int64_t n_blob, const uint8_t *gz_concat, const uint64_t *gz_off,
                                        const uint64_t *out_off, uint8_t *out_concat, int32_t *status) {
    ARG_CHECK(ctx && n_blob >= 0 && gz_off && out_off, "gunzip_batch: bad argument");
    if (n_blob == 0) return GAMS_CUDA_OK;
    ARG_CHECK(gz_concat && status, "gunzip_batch: NULL argument");
    ARG_CHECK(out_concat || out_off[n_blob] == out_off[0], "gunzip_batch: out_concat is NULL");
    for (int64_t i = 1; i <= n_blob; ++i)
        ARG_CHECK(gz_off[i] >= gz_off[i - 1] && out_off[i] >= out_off[i - 1], "gunzip_batch: offsets decrease at %lld", (long long)i);
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int n_dev = (int)std::min<int64_t>((int64_t)ctx->devs.size(), std::max<int64_t>(1, n_blob / 64));
    std::vector<int64_t> part((size_t)n_dev + 1);
    int32_t rc = gams_cuda_partition(n_blob, gz_off, n_dev, part.data());
    if (rc) return rc;
    rc = run_on_devices(n_dev, [&](int d) {
        const int64_t b0 = part[d], b1 = part[d + 1];
        return gunzip_one(ctx, d, b1 - b0, gz_concat, gz_off + b0, out_off + b0, out_concat, status + b0);
    });
    if (rc) return rc;
    for (int64_t i = 0; i < n_blob; ++i)
        if (status[i] != 0) {
            set_error("gunzip_batch: blob %lld failed with inflate status %d", (long long)i, (int)status[i]);
            return GAMS_CUDA_EINVAL;
        }
    return GAMS_CUDA_OK;
}
