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
// Lane 0 owns the bit reader and decodes one symbol per iteration — a 9-bit primary lookup table per alphabet in shared
// memory, canonical bit-by-bit decoding for the rare longer codes; the symbol is broadcast and the warp acts on it
// together: the 32 lanes copy a match (level-1 DNA is almost all matches of ~4 bases), stored blocks and table builds are
// cooperative as well.  History is read back from the output buffer itself (L1/L2), no window is staged.
// Bound: instruction issue / latency (~40 warp instructions per symbol), not HBM.
#include "common.cuh"

namespace gams {

constexpr int INF_WARPS = 8;        // per CTA
constexpr int INF_PBITS = 9;        // primary table index bits

struct InfTables {                  // per warp, shared memory (3 KB)
    uint16_t lcnt[16], dcnt[16];    // codes per length
    uint16_t lsym[288], dsym[32];   // symbols in canonical order
    uint16_t ltab[1 << INF_PBITS], dtab[1 << INF_PBITS];   // (symbol << 4) | length, 0 = longer than INF_PBITS
    uint8_t lens[320];              // code lengths of the block being set up
};

__constant__ uint16_t INF_LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t INF_LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t INF_DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t INF_DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t INF_CLORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct BitReader {                  // lives in lane 0's registers
    const uint8_t *in;
    uint64_t pos, end;              // next byte to load, end of the blob
    uint64_t buf;
    int cnt;
    __device__ __forceinline__ void refill() {   // at least 56 bits afterwards (zeros past the end)
        while (cnt <= 56) {
            const uint64_t b = pos < end ? (uint64_t)__ldg(in + pos) : 0ull;
            ++pos;
            buf |= b << cnt;
            cnt += 8;
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(int n) { buf >>= n; cnt -= n; }
    __device__ __forceinline__ uint32_t bits(int n) { const uint32_t v = peek(n); drop(n); return v; }   // n <= 16, caller refilled
    __device__ __forceinline__ bool overrun() const { return pos - (uint64_t)(cnt >> 3) > end; }          // consumed bytes past the end
};

// canonical decoding, one bit at a time (codes longer than the primary table; the code-length alphabet)
__device__ __forceinline__ int inf_decode_slow(BitReader &br, const uint16_t *cnt, const uint16_t *sym, int max_sym) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= (int)br.bits(1);
        const int c = cnt[len];
        if (code - c < first) {
            const int k = index + (code - first);
            return k < max_sym ? sym[k] : -1;
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}
__device__ __forceinline__ int inf_decode(BitReader &br, const uint16_t *tab, const uint16_t *cnt, const uint16_t *sym, int max_sym) {
    const uint32_t e = tab[br.peek(INF_PBITS)];
    if (e & 15u) {
        br.drop((int)(e & 15u));
        return (int)(e >> 4);
    }
    return inf_decode_slow(br, cnt, sym, max_sym);
}

// Warp-cooperative table build from code lengths lens[0 .. n): lane L (1..15) owns the codes of length L.
// Returns false for an over-subscribed set of lengths.
__device__ bool inf_build(const uint8_t *lens, int n, uint16_t *cnt, uint16_t *sym, uint16_t *tab, int lane) {
    for (int i = lane; i < (1 << INF_PBITS); i += 32) tab[i] = 0;
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
            if (lane <= INF_PBITS) {
                const uint32_t rev = __brev((uint32_t)(code + k)) >> (32 - lane);
                for (uint32_t j = rev; j < (1u << INF_PBITS); j += (1u << lane)) tab[j] = (uint16_t)((s << 4) | lane);
            }
            ++k;
        }
    }
    ok = __all_sync(0xFFFFFFFFu, ok);
    __syncwarp();
    return ok;
}

enum { INF_OK = 0, INF_EHEADER = 1, INF_ECODE = 2, INF_EOVERFLOW = 3, INF_EDIST = 4, INF_ESIZE = 5, INF_ETRUNC = 6 };

__global__ void __launch_bounds__(INF_WARPS * 32)
gunzip_kernel(const uint8_t *__restrict__ gz, const uint64_t *__restrict__ gz_off, int64_t n_blob, uint8_t *out_base,
              const uint64_t *__restrict__ out_off, int32_t *__restrict__ status) {
    __shared__ InfTables tabs[INF_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t blob = (int64_t)blockIdx.x * INF_WARPS + wib;
    if (blob >= n_blob) return;
    InfTables &T = tabs[wib];
    uint8_t *out = out_base + out_off[blob];
    const uint64_t cap = out_off[blob + 1] - out_off[blob];
    BitReader br;
    br.in = gz + gz_off[blob];
    br.pos = 0, br.end = gz_off[blob + 1] - gz_off[blob], br.buf = 0, br.cnt = 0;
    int err = INF_OK;
    // ---- gzip header (RFC 1952): every lane parses it (same bytes), only lane 0's reader is used afterwards
    {
        const uint8_t *h = br.in;
        const uint64_t n = br.end;
        if (n < 18 || h[0] != 0x1F || h[1] != 0x8B || h[2] != 8) err = INF_EHEADER;
        else {
            const uint32_t flg = h[3];
            uint64_t p = 10;
            if (flg & 4u) p += 2 + (p + 2 <= n ? (uint64_t)h[p] | ((uint64_t)h[p + 1] << 8) : 0);   // FEXTRA
            if (flg & 8u) { while (p < n && h[p]) ++p; ++p; }                                        // FNAME
            if (flg & 16u) { while (p < n && h[p]) ++p; ++p; }                                       // FCOMMENT
            if (flg & 2u) p += 2;                                                                    // FHCRC
            if (p + 8 > n) err = INF_EHEADER;
            br.pos = p;
        }
    }
    uint64_t pos = 0;               // bytes written; uniform over the warp
    bool last = false;
    while (!err && !last) {
        // ---- block header: lane 0 reads, everything the warp needs is broadcast
        uint32_t hdr = 0;
        if (lane == 0) {
            br.refill();
            hdr = br.bits(3);
        }
        hdr = __shfl_sync(0xFFFFFFFFu, hdr, 0);
        last = hdr & 1u;
        const uint32_t type = hdr >> 1;
        if (type == 0) {
            // stored: skip to the byte boundary, LEN / NLEN, copy
            uint64_t src = 0;
            uint32_t len = 0, bad = 0;
            if (lane == 0) {
                br.drop(br.cnt & 7);
                br.refill();
                len = br.bits(16);
                const uint32_t nlen = br.bits(16);
                bad = (len ^ 0xFFFFu) != nlen;
                src = br.pos - (uint64_t)(br.cnt >> 3);     // first byte not yet consumed
                br.pos = src + len;
                br.buf = 0, br.cnt = 0;
                if (br.pos > br.end) bad = 1;
            }
            len = __shfl_sync(0xFFFFFFFFu, len, 0);
            bad = __shfl_sync(0xFFFFFFFFu, bad, 0);
            src = __shfl_sync(0xFFFFFFFFu, src, 0);
            if (bad) { err = INF_ETRUNC; break; }
            if (pos + len > cap) { err = INF_EOVERFLOW; break; }
            for (uint32_t i = lane; i < len; i += 32) out[pos + i] = __ldg(br.in + src + i);
            pos += len;
            continue;
        }
        if (type == 3) { err = INF_ECODE; break; }
        int nlen = 288, ndist = 30;
        if (type == 1) {
            for (int s = lane; s < 288; s += 32) T.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
            if (lane < 30) T.lens[288 + lane] = 5;
        } else {
            // dynamic: code-length code, then HLIT + HDIST lengths (lane 0, sequential: ~300 short symbols per block)
            uint32_t info = 0;
            if (lane == 0) {
                br.refill();
                const int hlit = (int)br.bits(5) + 257, hdist = (int)br.bits(5) + 1, hclen = (int)br.bits(4) + 4;
                bool bad = hlit > 286 || hdist > 30;
                uint8_t cl[19];
                for (int i = 0; i < 19; ++i) cl[i] = 0;
                for (int i = 0; i < hclen; ++i) {
                    br.refill();
                    cl[INF_CLORDER[i]] = (uint8_t)br.bits(3);
                }
                // code-length alphabet: canonical tables in the (still unused) distance arrays
                uint16_t *ccnt = T.dcnt, *csym = T.dsym;
                for (int l = 0; l < 16; ++l) ccnt[l] = 0;
                for (int i = 0; i < 19; ++i) ccnt[cl[i]]++;
                ccnt[0] = 0;
                int offs[16];
                offs[1] = 0;
                for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + ccnt[l];
                for (int i = 0; i < 19; ++i)
                    if (cl[i]) csym[offs[cl[i]]++] = (uint16_t)i;
                int idx = 0;
                const int total = hlit + hdist;
                while (idx < total && !bad) {
                    br.refill();
                    const int s = inf_decode_slow(br, ccnt, csym, 19);
                    if (s < 0) { bad = true; break; }
                    if (s < 16) T.lens[idx++] = (uint8_t)s;
                    else {
                        int rep, v = 0;
                        if (s == 16) {
                            if (idx == 0) { bad = true; break; }
                            v = T.lens[idx - 1];
                            rep = 3 + (int)br.bits(2);
                        } else if (s == 17) rep = 3 + (int)br.bits(3);
                        else rep = 11 + (int)br.bits(7);
                        if (idx + rep > total) { bad = true; break; }
                        while (rep--) T.lens[idx++] = (uint8_t)v;
                    }
                }
                if (!bad && T.lens[256] == 0) bad = true;      // no end-of-block code
                info = bad ? 0xFFFFFFFFu : ((uint32_t)hlit | ((uint32_t)hdist << 16));
            }
            info = __shfl_sync(0xFFFFFFFFu, info, 0);
            if (info == 0xFFFFFFFFu) { err = INF_ECODE; break; }
            nlen = (int)(info & 0xFFFFu);
            ndist = (int)(info >> 16);
        }
        __syncwarp();
        const bool ok1 = inf_build(T.lens, nlen, T.lcnt, T.lsym, T.ltab, lane);
        const bool ok2 = inf_build(T.lens + (type == 1 ? 288 : nlen), ndist, T.dcnt, T.dsym, T.dtab, lane);
        if (!ok1 || !ok2) { err = INF_ECODE; break; }
        // ---- symbols: lane 0 decodes one, the warp executes it
        for (;;) {
            uint32_t a = 0, b = 0;      // a: symbol (0..255 literal, 256 end, 257 match, 0xFFFF error); b: length | distance << 16
            if (lane == 0) {
                br.refill();
                int s = inf_decode(br, T.ltab, T.lcnt, T.lsym, nlen);
                if (s < 0 || s > 285) a = 0xFFFFu;
                else if (s <= 256) a = (uint32_t)s;
                else {
                    s -= 257;
                    const uint32_t len = INF_LBASE[s] + br.bits(INF_LEXT[s]);
                    const int d = inf_decode(br, T.dtab, T.dcnt, T.dsym, ndist);
                    if (d < 0 || d > 29) a = 0xFFFFu;
                    else {
                        const uint32_t dist = INF_DBASE[d] + br.bits(INF_DEXT[d]);
                        a = 257u;
                        b = len | (dist << 16);
                    }
                }
                if (br.overrun()) a = 0xFFFFu;
            }
            a = __shfl_sync(0xFFFFFFFFu, a, 0);
            if (a < 256u) {
                if (pos >= cap) { err = INF_EOVERFLOW; break; }
                if (lane == 0) out[pos] = (uint8_t)a;
                ++pos;
                continue;
            }
            if (a == 256u) break;
            if (a != 257u) { err = INF_ECODE; break; }
            b = __shfl_sync(0xFFFFFFFFu, b, 0);
            const uint32_t len = b & 0xFFFFu, dist = b >> 16;
            if (dist > pos) { err = INF_EDIST; break; }
            if (pos + len > cap) { err = INF_EOVERFLOW; break; }
            __syncwarp();               // the bytes written so far (lane 0's literals, every lane's copies) are visible
            const uint8_t *src = out + pos - dist;
            uint8_t *dst = out + pos;
            if (dist >= len) {
                for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i];
            } else if (dist == 1) {
                const uint8_t v = src[0];
                for (uint32_t i = lane; i < len; i += 32) dst[i] = v;
            } else {
                for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i % dist];   // the period repeats
            }
            pos += len;
        }
    }
    // ---- trailer: ISIZE (RFC 1952: size of the original input modulo 2^32) and the size the caller expects
    if (lane == 0) {
        if (!err) {
            const uint64_t t = br.pos - (uint64_t)(br.cnt >> 3);   // first unread byte = CRC32
            if (t + 8 > br.end) err = INF_ETRUNC;
            else {
                const uint8_t *q = br.in + t + 4;
                const uint32_t isize = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
                if (isize != (uint32_t)pos || pos != cap) err = INF_ESIZE;
            }
        }
        status[blob] = err;
    }
}

int32_t launch_gunzip(cudaStream_t st, const uint8_t *d_gz, const uint64_t *d_gz_off, int64_t n_blob, uint8_t *d_out,
                      const uint64_t *d_out_off, int32_t *d_status) {
    if (n_blob == 0) return GAMS_CUDA_OK;
    const int64_t blocks = (n_blob + INF_WARPS - 1) / INF_WARPS;
    if (blocks >= (1ll << 31)) {
        set_error("gunzip: too many blobs for one launch");
        return GAMS_CUDA_EINVAL;
    }
    gunzip_kernel<<<(unsigned)blocks, INF_WARPS * 32, 0, st>>>(d_gz, d_gz_off, n_blob, d_out, d_out_off, d_status);
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
    DeviceGuard g(ctx->devs[dev_index].device);
    return launch_gunzip(static_cast<cudaStream_t>(stream), d_gz, d_gz_off, n_blob, d_out, d_out_off, d_status);
}

GAMS_API int32_t gams_cuda_gunzip_batch(gams_cuda_ctx *ctx, int64_t n_blob, const uint8_t *gz_concat, const uint64_t *gz_off,
                                        const uint64_t *out_off, uint8_t *out_concat, int32_t *status) {
    ARG_CHECK(ctx && n_blob >= 0 && gz_off && out_off, "gunzip_batch: bad argument");
    if (n_blob == 0) return GAMS_CUDA_OK;
    ARG_CHECK(gz_concat && status, "gunzip_batch: NULL argument");
    const uint64_t n_in = gz_off[n_blob] - gz_off[0], n_out = out_off[n_blob] - out_off[0];
    ARG_CHECK(out_concat || n_out == 0, "gunzip_batch: out_concat is NULL");
    std::vector<uint64_t> go((size_t)n_blob + 1), oo((size_t)n_blob + 1);
    for (int64_t i = 0; i <= n_blob; ++i) {
        go[i] = gz_off[i] - gz_off[0];
        oo[i] = out_off[i] - out_off[0];
        ARG_CHECK(i == 0 || (gz_off[i] >= gz_off[i - 1] && out_off[i] >= out_off[i - 1]), "gunzip_batch: offsets decrease at %lld", (long long)i);
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    cudaStream_t st = pd.stream;
    int32_t rc;
    if ((rc = pd.scratch[0].reserve(n_in + 64))) return rc;
    if ((rc = pd.scratch[1].reserve(std::max<uint64_t>(n_out, 1) + 64))) return rc;
    if ((rc = pd.scratch[2].reserve((size_t)(n_blob + 1) * 16 + (size_t)n_blob * 4 + 64))) return rc;
    uint8_t *d_gz = pd.scratch[0].as<uint8_t>(), *d_out = pd.scratch[1].as<uint8_t>();
    uint64_t *d_go = pd.scratch[2].as<uint64_t>(), *d_oo = d_go + (n_blob + 1);
    int32_t *d_status = reinterpret_cast<int32_t *>(d_oo + (n_blob + 1));
    CUDA_TRY(cudaMemcpyAsync(d_gz, gz_concat + gz_off[0], n_in, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_go, go.data(), (size_t)(n_blob + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_oo, oo.data(), (size_t)(n_blob + 1) * 8, cudaMemcpyHostToDevice, st));
    if ((rc = launch_gunzip(st, d_gz, d_go, n_blob, d_out, d_oo, d_status))) return rc;
    if (n_out) CUDA_TRY(cudaMemcpyAsync(out_concat + out_off[0], d_out, n_out, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(status, d_status, (size_t)n_blob * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n_blob; ++i)
        if (status[i] != 0) {
            set_error("gunzip_batch: blob %lld failed with inflate status %d", (long long)i, (int)status[i]);
            return GAMS_CUDA_EINVAL;
        }
    return GAMS_CUDA_OK;
}
