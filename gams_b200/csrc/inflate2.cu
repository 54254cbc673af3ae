// inflate2.cu — gunzip of a batch of `seq:{ctg}` blobs in three phases (Conn::get_seq -> decode_gz, src/libs/redis.rs:142-161).
//
// A deflate stream is bit-serial: where a symbol starts is known only after the one before it has been decoded.  The first
// kernel of this library (inflate.cu: one warp per blob, every lane decodes the same symbol, the lanes share the copies) spends
// ~100 dependent instructions per symbol on that chain, and the chain of the longest blob sets the time (128 ms per 3.1 Gbp).
// Here the chain carries only what is serial, and it is cut at every deflate block:
//   P inflate_parts_kernel    part_off[]: every blob is divided into RANGES of ~8 KB of compressed bytes (a level-1 block of
//                             sequence is ~22 KB, so a range holds at most one block header)
//   S inflate_split_kernel    one warp per range: where could a block start in it?  A deflate stream does not record that, so
//                             it is found speculatively — the first bit offset in the range that parses as a complete
//                             dynamic block header — in three stages of rising cost, each fed densely by the one before:
//                             (1) 32 offsets per lane and a dozen bit operations: BTYPE = 2, HLIT and HDIST in range (22 %
//                             pass); (2) the code-length code is complete, Kraft sum exactly 1, by table lookups of three
//                             lengths at a time (0.4 % pass); (3) the survivors of ~32 chunks are parsed in full, one per
//                             lane: both Huffman codes of the block must be complete too (none of 8.6e9 random offsets pass).
//   T inflate_tokens_kernel   one THREAD per part (the stretch from one found header to the next): Huffman decoding only, no
//                             output byte is read or written.  A symbol becomes a 32-bit token (literal byte, or length +
//                             distance); the two tables of a thread live in shared memory (6-bit literal/length index, 8-bit
//                             distance index: 640 bytes, 11 warps per SM), the bit window is three registers.  A part ends
//                             at the block header where the next part starts.
//   R inflate_resolve_kernel  one WARP per blob walks the chain of its parts — VERIFYING that every part ended exactly where
//                             the next one began, that the last one saw the final block and that the sizes agree with the
//                             trailer — 32 tokens at a time: output positions by a warp scan; then the batch's OUTPUT BYTES
//                             are dealt to the lanes (byte b of the batch to lane b mod 32): the lane finds the token that
//                             owns its byte (5 shuffle steps), loads the source byte and stores — 32 consecutive bytes per
//                             store instruction.
// A blob whose chain does not close (a false start, a token budget exceeded, any error) keeps status INF_ETOKENS and is
// decoded again by the one-pass kernel of inflate.cu, which also names the error: speculation can cost time, never
// correctness.  The token budget is half a token per output byte (+ 64): anything that compresses worse than 2 bytes per
// symbol is not a sequence.  The CRC-32 of the trailer is checked by crc32_*_kernel (below) when the caller asks for it
// (flate2's GzDecoder verifies it, redis.rs:156-161).
#include "common.cuh"
#include "inflate_common.cuh"

namespace gams {

#ifndef TK_WARPS_N
#define TK_WARPS_N 11
#endif
constexpr int TK_WARPS = TK_WARPS_N;           // warps per CTA, one CTA per SM: 352 parts in flight per SM
constexpr int TK_THREADS = TK_WARPS * 32;
constexpr int TK_LBITS = 6, TK_DBITS = 7;      // level-1 sequence data: 0.4 % of the literal/length codes and 1.8 % of the distance codes are longer
constexpr uint32_t TK_RING = 32;               // words of the compressed stream staged ahead in shared memory
// per-thread state in shared memory; the stride in 32-bit words is odd, so equal indices of different threads fall into
// different banks
struct TkFast {
    uint16_t ltab[1 << TK_LBITS];   // bits 0-3 code length + extra bits, 4-6 extra bits, 7: a length symbol, 8-15 the literal byte or the
                                    //   base length - 3.  Bits 0-3 zero: not for the branch-free body — 0: a longer code (or none),
                                    //   0x8000 | code length << 4: end of block, 0x4000: a symbol that must not occur (286, 287)
    uint16_t dtab[1 << TK_DBITS];   // bits 0-4 code length + extra bits, 5-8 extra bits, 9-13 symbol (0..29); 0: a longer code (or
                                    //   none), 0x8000: a symbol that must not occur (30, 31)
    uint8_t dsym[32];               // the distance symbols in canonical order (for the codes that are longer than the table index)
    uint32_t pad;
};
static_assert(sizeof(TkFast) % 4 == 0 && (sizeof(TkFast) / 4) % 2 == 1, "odd word stride");
// (and, behind the tables, a ring of TK_RING words per thread, 16-byte aligned for cp.async, and TK_STAGE tokens per thread)
constexpr uint32_t TK_STAGE = 16;
constexpr size_t TK_SMEM = (sizeof(TkFast) + 4 * TK_RING + 4 * TK_STAGE) * TK_THREADS;
// per-thread scratch in global memory: what a block header needs while it is parsed, and the canonical decoder of the codes
// that are longer than a table index
struct TkSlow {
    uint16_t lcnt[16], dcnt[16];    // codes per length
    uint16_t lfirst[16], dfirst[16];  // first canonical code of each length
    uint16_t loffs[16], doffs[16];    // position of that code's symbol in lsym / dsym
    uint16_t lsym[288];
    uint8_t dsym[32];
    uint8_t lens[320];
    uint8_t ctab[128];              // code-length alphabet: symbol << 3 | code length (0: no such code)
    uint8_t clen[20];
    uint8_t pad[12];
};
static_assert(sizeof(TkSlow) % 16 == 0, "TkSlow alignment");

__constant__ uint16_t TK_LBASE[32] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 0, 0, 0};
__constant__ uint8_t TK_LEXT[32] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0, 0};
__constant__ uint16_t TK_DBASE[32] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 0, 0};
__constant__ uint8_t TK_DEXT[32] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 0, 0};
__constant__ uint8_t TK_CLORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// The bit window of one thread: three consecutive aligned words of the blob and a bit offset o < 32 into the first — 64 valid
// bits from the current position in two funnel shifts, enough for a length and a distance symbol with their extra bits (48
// bits at most) without a refill in between — and one more word in a register.  What follows is staged in a ring of 32 words
// in shared memory by cp.async, 16 bytes at a time: the copy needs no register, so the symbol loop never waits for global
// memory.  The ring is topped up ONCE PER GROUP of symbols, by all lanes in the same instruction (top_up): a warp has one
// set of scoreboards, so a wait that only some lanes execute also waits for the copies the other lanes have just issued —
// with per-lane refills inside the loop (or with loads into registers, whose results the compiler's register moves at the
// loop's joins read) a sixth of the kernel's samples sat on that wait.  After a top-up at least 29 words lie ahead of the
// window, all but the newest copies have landed, and a group consumes at most 13: rotate_fast() needs no check.  Outside
// the symbol loop (block headers, stored blocks) rotate() checks and refills with a blocking wait.
// Bytes behind the blob's end read as zero (cp.async src-size), never part of a valid stream: the trailer is 8 bytes.
struct TkWin {
    const uint8_t *base;             // the blob's first byte rounded down to 16 bytes
    uint32_t n_bytes, n_words;       // from base to the blob's end
    uint32_t wi, mis;                // w0 w1 w2 p0 = words wi - 3 .. wi of base; mis: bytes between base and the blob's first byte
    uint32_t w0, w1, w2, p0, p1, o;  // p1: word wi + 1, read from the ring a rotation before it is needed
    uint32_t ring;                   // shared-memory address of the lane's ring; word k of the stream lives in slot k & 31
    uint32_t fetched, landed;        // words < fetched are in the ring or on their way, words < landed have arrived (both multiples of 4)
    __device__ __forceinline__ void set_blob(const uint8_t *in, uint32_t in_len) {
        mis = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 15u);
        base = in - mis;
        n_bytes = mis + in_len;
        n_words = (n_bytes + 3u) >> 2;
    }
    __device__ __forceinline__ void fetch4(uint32_t k4) const {     // words k4 .. k4 + 3 (k4 a multiple of 4) into their slots
        const uint32_t byte = 4u * k4;
        const uint32_t have = byte < n_bytes ? min(n_bytes - byte, 16u) : 0u;
        const uint8_t *src = base + (have ? byte : 0u);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ring + 4u * (k4 & (TK_RING - 1u))), "l"(src), "r"(have) : "memory");
    }
    __device__ __forceinline__ uint32_t slot(uint32_t k) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ring + 4u * (k & (TK_RING - 1u))) : "memory");
        return v;
    }
    __device__ __forceinline__ void fill() {     // every free slot of the ring: slots of words <= wi are free
#pragma unroll
        for (uint32_t k = 0; k < TK_RING / 4; ++k)
            if (fetched + 3u <= wi + TK_RING) fetch4(fetched), fetched += 4u;
    }
    // all lanes of the warp together (active says whether this lane has a stream): fill, and everything but the copies
    // issued just now has arrived
    __device__ __forceinline__ void top_up(bool active) {
        const uint32_t before = fetched;
        if (active) {
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)         // (a group consumes 13 words at most)
                if (fetched + 3u <= wi + TK_RING) fetch4(fetched), fetched += 4u;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        landed = before;
        if (active && landed < wi + 17u) {      // (the first group behind a block header: the ring was refilled word by word)
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            landed = fetched;
        }
    }
    __device__ __forceinline__ void seek_bit(uint64_t bit) {   // position given in bits from the blob's first byte
        const uint64_t a = 8ull * mis + bit;
        const uint32_t k = (uint32_t)(a >> 5);
        o = (uint32_t)(a & 31u);
        asm volatile("cp.async.wait_group 0;" ::: "memory");     // (nothing of an earlier position may land in the ring later)
        fetched = k & ~3u;
        wi = fetched > 0u ? fetched - 1u : 0u;                  // (every slot is free)
        fill();
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        landed = fetched;
        w0 = slot(k), w1 = slot(k + 1), w2 = slot(k + 2), p0 = slot(k + 3), p1 = slot(k + 4);
        wi = k + 3;
    }
    __device__ __forceinline__ void seek(uint32_t byte) { seek_bit(8ull * byte); }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(w0, w1, o); }      // bits 0..31 from the position
    __device__ __forceinline__ uint32_t peek_hi() const { return __funnelshift_r(w1, w2, o); }   // bits 32..63
    __device__ __forceinline__ void rotate_fast() {
        w0 = w1, w1 = w2, w2 = p0, p0 = p1;
        ++wi;
        p1 = slot(wi + 1);
    }
    __device__ __forceinline__ void rotate() {
        if (wi + 2u >= landed) {
            fill();
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            landed = fetched;
        }
        rotate_fast();
    }
    __device__ __forceinline__ void drop(uint32_t n) {   // n <= 63 (0 is a no-op)
        o += n;
        if (o >= 32u) { rotate(); o -= 32u; }
        if (o >= 32u) { rotate(); o -= 32u; }
    }
    __device__ __forceinline__ void drop_fast(uint32_t n) {   // inside a group of symbols, after top_up()
        o += n;                                               // n <= 31: a table symbol is 11 bits at most, a table distance 20
        if (o >= 32u) { rotate_fast(); o -= 32u; }
    }
    __device__ __forceinline__ uint32_t bits(uint32_t n) { const uint32_t v = peek() & ((1u << n) - 1u); drop(n); return v; }   // n < 32
    __device__ __forceinline__ uint64_t consumed_bits() const { return 32ull * (wi - 3) + o - 8ull * mis; }
    __device__ __forceinline__ bool past_end() const { return wi > n_words + 3u; }   // the position is behind the blob's last word
};

__device__ __forceinline__ uint32_t tk_lit_entry(int s, int code_len) {
    if (s < 256) return ((uint32_t)s << 8) | (uint32_t)code_len;
    if (s == 256) return 0x8000u | ((uint32_t)code_len << 4);
    if (s <= 285) {
        const uint32_t ex = TK_LEXT[s - 257];
        return ((uint32_t)(TK_LBASE[s - 257] - 3) << 8) | 128u | (ex << 4) | ((uint32_t)code_len + ex);
    }
    return 0x4000u;
}
__device__ __forceinline__ uint32_t tk_dist_entry(int s, int code_len) {
    if (s > 29) return 0x8000u;
    const uint32_t dx = TK_DEXT[s];
    return ((uint32_t)s << 9) | (dx << 5) | ((uint32_t)code_len + dx);
}

// canonical decoding of a code that is longer than the table index (PBITS): fc[k] holds, for length l = PBITS + 1 + k, the
// first canonical code (bits 0-15) and the position of its symbol in canonical order (bits 16-31); the number of codes of the
// length follows from the next length's first code, first(l + 1) = (first(l) + count(l)) << 1, and is given for length 15.
// The code of length l is the top l bits of the bit-reversed window; it belongs to this length iff it lies in
// [first, first + count).  Returns the length (0: no such code) and the symbol's position.
template <int PBITS>
__device__ __forceinline__ uint32_t tk_decode_long(uint32_t pk, const uint32_t (&fc)[15 - PBITS], uint32_t cnt15, uint32_t *pos) {
    const uint32_t c = __brev(pk);
    uint32_t len = 0, r = 0;
#pragma unroll
    for (int k = 15 - PBITS - 1; k >= 0; --k) {          // (downwards, so that the shortest length that fits stays)
        const uint32_t first = fc[k] & 0xFFFFu;
        const uint32_t cnt = k == 15 - PBITS - 1 ? cnt15 : ((fc[k == 15 - PBITS - 1 ? k : k + 1] & 0xFFFFu) >> 1) - first;
        const uint32_t idx = (c >> (32 - (PBITS + 1 + k))) - first;
        if (idx < cnt) len = PBITS + 1 + k, r = (fc[k] >> 16) + idx;
    }
    *pos = r;
    return len;
}

// Warp-cooperative build of ONE thread's tables from its code lengths lens[0 .. n): lane L (1..15) owns the codes of length L.
// false for an over-subscribed code.
template <int PBITS, bool DIST>
__device__ bool tk_build(const uint8_t *lens, int n, uint16_t *cnt, uint16_t *first, uint16_t *offs_, void *sym_, uint16_t *tab, int lane) {
    for (int i = lane; i < (1 << PBITS); i += 32) tab[i] = 0;   // not in this table
    int mine = 0;
    if (lane >= 1 && lane <= 15)
        for (int s = 0; s < n; ++s) mine += (lens[s] == lane) ? 1 : 0;
    if (lane < 16) cnt[lane] = (uint16_t)(lane ? mine : 0);
    __syncwarp();
    bool ok = true;
    if (lane >= 1 && lane <= 15) {
        int offs = 0, code = 0, left = 1;
        for (int l = 1; l <= 15; ++l) {
            left <<= 1;
            left -= cnt[l];
            if (left < 0) ok = false;
        }
        for (int l = 1; l < lane; ++l) {
            offs += cnt[l];
            code = (code + cnt[l]) << 1;
        }
        first[lane] = (uint16_t)code;
        offs_[lane] = (uint16_t)offs;
        int k = 0;
        for (int s = 0; s < n && ok; ++s) {
            if (lens[s] != lane) continue;
            if (DIST) reinterpret_cast<uint8_t *>(sym_)[offs + k] = (uint8_t)s;
            else reinterpret_cast<uint16_t *>(sym_)[offs + k] = (uint16_t)s;
            if (lane <= PBITS) {
                const uint32_t rev = __brev((uint32_t)(code + k)) >> (32 - lane);
                const uint16_t e = (uint16_t)(DIST ? tk_dist_entry(s, lane) : tk_lit_entry(s, lane));
                for (uint32_t j = rev; j < (1u << PBITS); j += (1u << lane)) tab[j] = e;
            }
            ++k;
        }
    }
    ok = __all_sync(0xFFFFFFFFu, ok);
    __syncwarp();
    return ok;
}

constexpr uint32_t TK_NONE = 0xFFFFFFFFu;
constexpr uint32_t TK_RANGE = 8192;              // compressed bytes per range, at least
struct alignas(32) PartRec {
    uint32_t start_bit;    // where the part starts (bits from the blob's first byte); TK_NONE: no such part
    uint32_t end_bit;      // where it stopped: the start of the next part if the chain closes
    uint32_t n_tok, out_len;
    int32_t status;        // bit 16: the part decoded the final block
    uint32_t isize, crc;   // the trailer (last part only)
    uint32_t tok_off;      // where its tokens are, from the blob's first token
};

// ---- phase P: ranges per blob -----------------------------------------------------------------------------------------
// The batch has room for p_cap part records; the range size is TK_RANGE, or larger when the batch would need more records.
__device__ __forceinline__ uint64_t tk_range_bytes(const uint64_t *__restrict__ gz_off, int64_t n_blob, uint64_t p_cap) {
    const uint64_t total = gz_off[n_blob] - gz_off[0];
    const uint64_t room = p_cap > (uint64_t)n_blob + 1 ? p_cap - (uint64_t)n_blob - 1 : 1;
    const uint64_t t = (total + room - 1) / room;
    return t < TK_RANGE ? TK_RANGE : t;
}
// the first token of blob b (a multiple of 8) — the budget of a blob is half a token per output byte + 64
__device__ __forceinline__ uint64_t tk_tok_base(uint64_t o0_rel, uint64_t blob) { return ((o0_rel >> 1) + 64ull * blob) & ~7ull; }
// where the tokens of the part that is range j of np and starts at bit s go, from the blob's first token: in proportion to
// the position in the compressed stream, a multiple of 8, and at least 8 tokens behind any earlier part
__device__ __forceinline__ uint32_t tk_tok_off(uint64_t budget, uint32_t np, uint64_t nbits, uint64_t s, uint32_t j) {
    return (uint32_t)((((budget - 8ull * np) * s / nbits) & ~7ull) + 8ull * j);
}

constexpr uint32_t SP_GROUP = 64;      // ranges per warp of the split kernel
__global__ void __launch_bounds__(1024)
inflate_parts_kernel(const uint64_t *__restrict__ gz_off, const uint64_t *__restrict__ out_off, int64_t n_blob, uint64_t p_cap,
                     uint32_t *__restrict__ part_off, uint32_t *__restrict__ group_off) {
    __shared__ uint32_t ws[32], ws2[32];
    __shared__ uint32_t carry, carry2;
    if (threadIdx.x == 0) carry = carry2 = 0;
    __syncthreads();
    const uint64_t range = tk_range_bytes(gz_off, n_blob, p_cap), ob = out_off[0];
    for (int64_t base = 0; base < n_blob; base += 1024) {
        const int64_t b = base + threadIdx.x;
        uint32_t v = 0;
        if (b < n_blob) {
            const uint64_t len = gz_off[b + 1] - gz_off[b];
            const uint64_t budget = tk_tok_base(out_off[b + 1] - ob, (uint64_t)b + 1) - tk_tok_base(out_off[b] - ob, (uint64_t)b);
            uint64_t np = (len + range - 1) / range;
            if (np < 1 || len >= 0x10000000ull) np = 1;        // (bit offsets are 32-bit)
            if (16ull * np > budget) np = 1;
            v = (uint32_t)np;
        }
        const uint32_t v2 = (v + SP_GROUP - 1) / SP_GROUP;
        uint32_t inc = v, inc2 = v2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d), t2 = __shfl_up_sync(0xFFFFFFFFu, inc2, d);
            if ((threadIdx.x & 31) >= d) inc += t, inc2 += t2;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc, ws2[threadIdx.x >> 5] = inc2;
        __syncthreads();
        uint32_t before = carry, before2 = carry2;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += ws[w], before2 += ws2[w];
        if (b < n_blob) part_off[b] = before + inc - v, group_off[b] = before2 + inc2 - v2;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + inc, carry2 = before2 + inc2;
        __syncthreads();
    }
    if (threadIdx.x == 0) part_off[n_blob] = carry, group_off[n_blob] = carry2;
}

// the blob of part item: the last b with part_off[b] <= item
__device__ __forceinline__ int64_t tk_blob_of(const uint32_t *__restrict__ part_off, int64_t n_blob, uint32_t item) {
    int64_t lo = 0, hi = n_blob;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(part_off + mid) <= item) lo = mid; else hi = mid;
    }
    return lo;
}

// HLIT / HDIST / HCLEN and the code lengths of a dynamic block (the 3 header bits are already consumed): lens[0 .. hlit + hdist).
// ctab / clen: scratch of 128 / 20 bytes.  false: not a valid header.
__device__ bool tk_dynamic_header(TkWin &br, uint8_t *lens, uint8_t *ctab, uint8_t *clen, int *hlit_, int *hdist_) {
    const int hlit = (int)br.bits(5) + 257, hdist = (int)br.bits(5) + 1, hclen = (int)br.bits(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    for (int i = 0; i < 19; ++i) clen[i] = 0;
    for (int i = 0; i < hclen; ++i) clen[TK_CLORDER[i]] = (uint8_t)br.bits(3);
    {   // the code-length code (lengths <= 7) as a 7-bit table
        int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, next[8], left = 1;
        for (int i = 0; i < 19; ++i) cnt[clen[i]]++;
        cnt[0] = 0;
        int code = 0;
        for (int l = 1; l <= 7; ++l) {
            left = (left << 1) - cnt[l];
            code = (code + cnt[l - 1]) << 1;
            next[l] = code;
        }
        if (left < 0) return false;
        for (int i = 0; i < 128; ++i) ctab[i] = 0;
        for (int i = 0; i < 19; ++i) {
            const int l = clen[i];
            if (!l) continue;
            const uint32_t rev = __brev((uint32_t)next[l]++) >> (32 - l);
            for (uint32_t j = rev; j < 128u; j += (1u << l)) ctab[j] = (uint8_t)((i << 3) | l);
        }
    }
    int idx = 0, prev = 0;
    const int total = hlit + hdist;
    while (idx < total) {
        const uint32_t ce = ctab[br.peek() & 127u];
        if (!(ce & 7u)) return false;
        br.drop(ce & 7u);
        const int s = (int)(ce >> 3);
        if (s < 16) {
            lens[idx++] = (uint8_t)s;
            prev = s;
        } else {
            int rep, v = 0;
            if (s == 16) {
                if (idx == 0) return false;
                v = prev;
                rep = 3 + (int)br.bits(2);
            } else if (s == 17) rep = 3 + (int)br.bits(3);
            else rep = 11 + (int)br.bits(7);
            if (idx + rep > total) return false;
            for (int k2 = 0; k2 < rep; ++k2) lens[idx + k2] = (uint8_t)v;
            idx += rep;
            prev = v;
        }
    }
    if (lens[256] == 0) return false;      // no end-of-block code
    *hlit_ = hlit;
    *hdist_ = hdist;
    return true;
}

// ---- phase S: where can a part start? ------------------------------------------------------------------------------------
constexpr int SP_WARPS = 4;
struct SpWarp {
    uint32_t words[36];         // the chunk: 32 words of bit offsets + the 3 words the last offsets reach into
    uint16_t cand[1024];        // stage 1 survivors of the chunk: word << 5 | bit
    uint32_t surv[64];          // stage 2 survivors: bit offset from the blob's first byte
    uint8_t ctab[32][128];      // stage 3: the code-length code of each lane's candidate
};
// Stage 3, one candidate per lane: the whole dynamic header at bit `bit` (BFINAL) of the blob, Kraft sums kept on the fly.
// 1: the literal/length code is complete and has an end-of-block code, and the distance code is complete (or a single code);
// 0: not a header; 2: undecided after max_steps code-length symbols (the caller asks again with the full budget).
__device__ int sp_verify(const uint32_t *__restrict__ w, uint32_t n_words, uint32_t mis, uint32_t bit, uint8_t *ctab, int max_steps) {
    // the reader: three words in registers, reloaded a word ahead
    uint64_t a = 8ull * mis + bit + 3;     // absolute bit in w[]
    uint32_t kw = (uint32_t)(a >> 5), sh = (uint32_t)(a & 31u);
    auto ldw = [&](uint32_t k) { return __ldg(w + min(k, n_words - 1u)); };
    uint32_t q0 = ldw(kw), q1 = ldw(kw + 1), q2 = ldw(kw + 2);
    auto peek = [&]() -> uint32_t { return __funnelshift_r(q0, q1, sh); };
    auto skip = [&](uint32_t n) {          // n <= 32
        sh += n;
        if (sh >= 32u) {
            sh -= 32u;
            ++kw;
            q0 = q1, q1 = q2, q2 = ldw(kw + 2);
        }
    };
    uint32_t x = peek();
    const int hlit = (int)(x & 31u) + 257, hdist = (int)((x >> 5) & 31u) + 1, hclen = (int)((x >> 10) & 15u) + 4;
    skip(14);
    if (hlit > 286 || hdist > 30) return 0;
    uint64_t cl = 0;                       // 3 bits per symbol of the code-length alphabet
    uint64_t cnt = 0;                      // 8 bits per length
    for (int i = 0; i < hclen; ++i) {
        if (i % 8 == 0) x = peek(), skip(min(24, 3 * (hclen - i)));
        const uint32_t l = (x >> (3 * (i % 8))) & 7u;
        cl |= (uint64_t)l << (3 * TK_CLORDER[i]);
        if (l) cnt += 1ull << (8 * l);
    }
    uint64_t next = 0;                     // first canonical code of each length, 8 bits per length
    {
        int code = 0, left = 1, prevc = 0;
        for (int l = 1; l <= 7; ++l) {
            const int c = (int)((cnt >> (8 * l)) & 255u);
            left = (left << 1) - c;
            code = (code + prevc) << 1;
            prevc = c;
            next |= (uint64_t)(code & 255) << (8 * l);
        }
        if (left < 0) return 0;
    }
    for (int i = 0; i < 128; i += 4) *reinterpret_cast<uint32_t *>(ctab + i) = 0u;
    for (int i = 0; i < 19; ++i) {
        const int l = (int)((cl >> (3 * i)) & 7u);
        if (!l) continue;
        const uint32_t code = (uint32_t)((next >> (8 * l)) & 255u);
        next += 1ull << (8 * l);
        const uint32_t rev = __brev(code) >> (32 - l);
        for (uint32_t j = rev; j < 128u; j += (1u << l)) ctab[j] = (uint8_t)((i << 3) | l);
    }
    const int total = hlit + hdist;
    int idx = 0, prev = 0, nd = 0;
    int kl = 0, kd = 0;                    // Kraft sums in units of 2^-15
    bool eob = false;
    for (int step = 0; idx < total; ++step) {
        if (step >= max_steps) return 2;
        x = peek();
        const uint32_t ce = ctab[x & 127u];
        if (!(ce & 7u)) return 0;
        x >>= ce & 7u;
        const int s = (int)(ce >> 3);
        int rep = 1, v = s;
        uint32_t used = ce & 7u;
        if (s == 16) {
            if (idx == 0) return 0;
            v = prev, rep = 3 + (int)(x & 3u), used += 2;
        } else if (s == 17) v = 0, rep = 3 + (int)(x & 7u), used += 3;
        else if (s == 18) v = 0, rep = 11 + (int)(x & 127u), used += 7;
        skip(used);
        if (idx + rep > total) return 0;
        if (v) {
            const int unit = 1 << (15 - v);
            const int in_l = min(rep, max(hlit - idx, 0));      // how many of the run fall into the literal/length alphabet
            kl += in_l * unit;
            kd += (rep - in_l) * unit;
            nd += rep - in_l;
            if (idx <= 256 && 256 < idx + rep) eob = true;
            if (kl > (1 << 15) || kd > (1 << 15)) return 0;
        }
        idx += rep;
        prev = v;
    }
    return (eob && kl == (1 << 15) && (kd == (1 << 15) || (nd == 1 && kd == (1 << 14)))) ? 1 : 0;
}

// One warp per GROUP of SP_GROUP consecutive ranges of a blob.  The warp walks its stretch of the stream from header to
// header: scan() finds the first header at or behind a position; two headers in a row give the size of a block, and
// encoders cut blocks at a fixed number of symbols (zlib: 16 383), so the next header is looked for from 13/16 of that size
// behind the last one on — a block whose header lies before that point is simply not found and stays part of the block
// before it (a longer part, nothing else).  Headers go into the record of the range they lie in (the first one per range).
// Scanned: a whole block once per group to learn the size, then ~1/5 of every block, instead of 4/5 of the stream.
__global__ void __launch_bounds__(SP_WARPS * 32)
inflate_split_kernel(const uint8_t *__restrict__ gz, const uint64_t *__restrict__ gz_off, int64_t n_blob,
                     const uint32_t *__restrict__ part_off, const uint32_t *__restrict__ group_off, PartRec *__restrict__ rec,
                     int64_t blob0, int64_t blob1) {      // the groups of blobs blob0 .. blob1 - 1
    __shared__ SpWarp sw_[SP_WARPS];
    __shared__ uint8_t kraft3[512];        // Kraft sum (units of 1/128) of three 3-bit code lengths
    for (int i = threadIdx.x; i < 512; i += SP_WARPS * 32) {
        uint32_t s = 0;
        for (int f = 0; f < 3; ++f) {
            const uint32_t l = (i >> (3 * f)) & 7u;
            s += l ? (128u >> l) : 0u;
        }
        kraft3[i] = (uint8_t)s;            // <= 192
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    SpWarp &S = sw_[wib];
    const uint32_t n_items = group_off[blob1];
    for (uint32_t item = group_off[blob0] + blockIdx.x * SP_WARPS + wib; item < n_items; item += gridDim.x * SP_WARPS) {
        const int64_t blob = tk_blob_of(group_off, n_blob, item);
        const uint32_t np = part_off[blob + 1] - part_off[blob];
        const uint32_t j0 = (item - group_off[blob]) * SP_GROUP, j1 = min(j0 + SP_GROUP, np);
        PartRec *R = rec + part_off[blob];
        for (uint32_t j = max(j0, 1u) + lane; j < j1; j += 32) R[j].start_bit = TK_NONE;      // part 0 starts behind the gzip header
        __syncwarp();
        if (np == 1) continue;
        const uint64_t len64 = gz_off[blob + 1] - gz_off[blob];
        const uint8_t *in = gz + gz_off[blob];
        const uint32_t in_len = (uint32_t)len64, mis = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 3u);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(in - mis);
        const uint32_t n_words = (mis + in_len + 3u) >> 2;
        const uint64_t g_bit0 = j0 == 0 ? 80ull : 8ull * ((uint64_t)in_len * j0 / np);        // (a gzip header is 10 bytes at least)
        const uint64_t g_lim = min(8ull * ((uint64_t)in_len * j1 / np), 8ull * (uint64_t)in_len > 160ull ? 8ull * (uint64_t)in_len - 160ull : 0ull);
        // ---- the first header at a bit offset in [bit0, lim), or TK_NONE
        auto scan = [&](uint64_t bit0, uint64_t lim) -> uint32_t {
            uint32_t found = TK_NONE, n_surv = 0;
            // stage 3 on the first min(n_surv, 32) survivors; the list is in rising order, so the lowest valid one is the answer
            auto verify = [&]() {
                const uint32_t n = min(n_surv, 32u);
                const bool ok = lane < n && sp_verify(w, n_words, mis, S.surv[lane], S.ctab[lane], 1 << 20) == 1;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
                if (m) found = S.surv[__ffs(m) - 1];
                __syncwarp();
                const uint32_t rest = n_surv - n;
                const uint32_t mv = lane < rest ? S.surv[n + lane] : 0u;
                __syncwarp();
                if (lane < rest) S.surv[lane] = mv;
                n_surv = rest;
                __syncwarp();
            };
            const uint64_t a0 = 8ull * mis + bit0, a1 = 8ull * mis + lim;           // absolute bits in w[]
            for (uint64_t kc = a0 >> 5; (kc << 5) < a1 && found == TK_NONE; kc += 32) {
                // ---- the chunk's words
                {
                    const uint64_t k = kc + lane;
                    S.words[lane] = __ldg(w + min(k, (uint64_t)n_words - 1u));
                    if (lane < 4) S.words[32 + lane] = __ldg(w + min(k + 32, (uint64_t)n_words - 1u));
                }
                __syncwarp();
                // ---- stage 1: bit i of m is set iff offset 32 * (kc + lane) + i can be a dynamic block header
                const uint64_t v = (uint64_t)S.words[lane] | ((uint64_t)S.words[lane + 1] << 32);
                uint32_t m = (uint32_t)(~(v >> 1) & (v >> 2) & ~((v >> 4) & (v >> 5) & (v >> 6) & (v >> 7)) & ~((v >> 9) & (v >> 10) & (v >> 11) & (v >> 12)));
                {
                    const uint64_t wa = (kc + lane) << 5;           // absolute bit of the lane's word
                    if (wa + 32 <= a0 || wa >= a1) m = 0;
                    else {
                        if (wa < a0) m &= ~0u << (uint32_t)(a0 - wa);
                        if (wa + 32 > a1) m &= ~0u >> (uint32_t)(wa + 32 - a1);
                    }
                }
                const uint32_t cnt = __popc(m);
                uint32_t inc = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                    if (lane >= (uint32_t)d) inc += t;
                }
                const uint32_t n_cand = __shfl_sync(0xFFFFFFFFu, inc, 31);
                uint32_t at = inc - cnt;
                while (m) {
                    const uint32_t i = (uint32_t)__ffs(m) - 1u;
                    m &= m - 1;
                    S.cand[at++] = (uint16_t)((lane << 5) | i);
                }
                __syncwarp();
                // ---- stage 2: the code-length code is complete
                for (uint32_t c0 = 0; c0 < n_cand && found == TK_NONE; c0 += 32) {
                    const uint32_t c = c0 + lane;
                    bool pass = false;
                    uint32_t bit = 0;
                    if (c < n_cand) {
                        const uint32_t id = S.cand[c], wl = id >> 5, i = id & 31u;
                        const uint32_t q0 = S.words[wl], q1 = S.words[wl + 1], q2 = S.words[wl + 2], q3 = S.words[wl + 3];
                        const uint32_t x0 = __funnelshift_r(q0, q1, i), x1 = __funnelshift_r(q1, q2, i), x2 = __funnelshift_r(q2, q3, i);
                        const uint32_t hclen = ((x0 >> 13) & 15u) + 4u;
                        uint64_t f = ((uint64_t)x1 << 15) | (uint64_t)(x0 >> 17) | ((uint64_t)x2 << 47);     // the 3-bit lengths from bit 17 on
                        f &= (1ull << (3u * hclen)) - 1ull;
                        const uint32_t flo = (uint32_t)f, fhi = (uint32_t)(f >> 32);
                        const uint32_t k = kraft3[flo & 511u] + kraft3[(flo >> 9) & 511u] + kraft3[(flo >> 18) & 511u] +
                                           kraft3[__funnelshift_r(flo, fhi, 27) & 511u] + kraft3[(fhi >> 4) & 511u] + kraft3[(fhi >> 13) & 511u] +
                                           kraft3[(fhi >> 22) & 511u];
                        pass = k == 128u;
                        bit = (uint32_t)((((kc + wl) << 5) + i) - 8ull * mis);
                    }
                    const unsigned pm = __ballot_sync(0xFFFFFFFFu, pass);
                    if (pass) S.surv[n_surv + __popc(pm & ((1u << lane) - 1u))] = bit;
                    n_surv += __popc(pm);
                    __syncwarp();
                    if (n_surv >= 32) verify();
                }
            }
            while (n_surv && found == TK_NONE) verify();
            return found;
        };
        // ---- from header to header
        uint64_t cur = g_bit0;
        uint32_t prev = TK_NONE;
        while (cur < g_lim) {
            const uint32_t h = scan(cur, g_lim);
            if (h == TK_NONE) break;
            const uint32_t j = (uint32_t)((uint64_t)h * np / (8ull * in_len));     // (h < 8 in_len: j < np)
            if (lane == 0 && j >= 1 && j >= j0 && j < j1 && R[j].start_bit == TK_NONE) R[j].start_bit = h;
            uint64_t skip = 64;                                                       // (a dynamic header is longer than this)
            if (prev != TK_NONE) skip = max((unsigned long long)skip, (unsigned long long)(h - prev) * 13ull / 16ull);
            prev = h;
            cur = (uint64_t)h + skip;
        }
        __syncwarp();
    }
}

enum { TS_HEADER = 0, TS_SYMBOLS = 1, TS_DONE = 2, TS_FINISHED = 3, TS_FETCH = 4 };
#ifndef TK_GROUP_N
#define TK_GROUP_N 8
#endif
constexpr int TK_GROUP = TK_GROUP_N;
#ifndef TK_UNROLL_N
#define TK_UNROLL_N 8
#endif
constexpr int TK_UNROLL = TK_UNROLL_N;
constexpr int32_t TK_FINAL = 1 << 16;

// token: bits 0-8 length (0: literal, byte in bits 9-16; 3..258: match, distance in bits 9-24).
// The 32 lanes decode 32 different streams, so the kernel is written to keep them CONVERGED: the symbol loop has one
// branch-free body (a literal / length symbol and, predicated, the distance symbol behind it) that every lane executes in
// lock step; what only some lanes need — a block header, a code longer than the table index, the trailer — runs in short
// divergent sections between votes, and table builds are done by the whole warp for one lane at a time.  (Lanes that each
// follow their own path share the warp's one instruction stream: a first version written as 32 independent loops took
// 1 400 cycles per symbol.)  Whatever does not fit — a distance before the start of the output, an output or a token
// budget exceeded, the end of the input — ends the part with a status; the resolver then hands the blob to the one-pass decoder.
__global__ void __launch_bounds__(TK_THREADS, 1)
inflate_tokens_kernel(const uint8_t *__restrict__ gz, const uint64_t *__restrict__ gz_off, int64_t n_blob,
                      const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ part_off, uint32_t *__restrict__ tokens,
                      PartRec *__restrict__ rec, TkSlow *__restrict__ slow, unsigned long long *__restrict__ next_item) {
    extern __shared__ __align__(16) uint8_t tk_smem[];
    __shared__ uint16_t s_lbase[32], s_dbase[32];   // (constant memory would serialise the lanes' different indices)
    __shared__ uint8_t s_lext[32], s_dext[32];
    if (threadIdx.x < 32) {
        s_lbase[threadIdx.x] = TK_LBASE[threadIdx.x];
        s_dbase[threadIdx.x] = TK_DBASE[threadIdx.x];
        s_lext[threadIdx.x] = TK_LEXT[threadIdx.x];
        s_dext[threadIdx.x] = TK_DEXT[threadIdx.x];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    TkFast *tabs = reinterpret_cast<TkFast *>(tk_smem) + (threadIdx.x & ~31);      // the warp's 32 tables
    TkFast &F = tabs[lane];
    TkSlow *slows = slow + ((size_t)blockIdx.x * TK_THREADS + (threadIdx.x & ~31));
    TkSlow &T = slows[lane];
    // The kernel is PERSISTENT: a lane that has finished a part takes the next one off a global counter, so the 32 lanes of
    // a warp stay busy although parts differ in length (statically, a warp would wait for its longest part).
    const uint32_t n_items = part_off[n_blob];
    const uint64_t ob = out_off[0];
    // ---- the lane's current part
    const uint8_t *in = gz;
    PartRec *R = rec;
    uint32_t next_start = TK_NONE, my_start = 0, cap = 0, in_len = 0, tcap = 0, tok_off = 0;
    uint32_t *tok = tokens;
    uint32_t r_isize = 0, r_crc = 0, r_final = 0, pos = 0, nt = 0, flushed = 0;
    // Tokens are staged in shared memory, 16 per lane, and leave as whole sectors: 8 tokens = 32 bytes, two 128-bit stores
    // per lane.  (One 4-byte store per token from 32 lanes into 32 different lines held the memory pipe for 32 cycles per
    // symbol: the kernel took 12.7 ms with those stores and 6.4 ms without any.)  Token regions start and end at multiples
    // of 8 tokens, so the last sector of a part may carry up to 7 slots of padding.
    const uint32_t stage = (uint32_t)__cvta_generic_to_shared(tk_smem + (sizeof(TkFast) + 4 * TK_RING) * TK_THREADS + 4 * TK_STAGE * threadIdx.x);
    auto put = [&](uint32_t t) {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(stage + 4u * (nt & (TK_STAGE - 1u))), "r"(t) : "memory");
        ++nt;
    };
    auto flush8 = [&]() {        // tokens flushed .. flushed + 7
        uint4 a, b;
        const uint32_t at = stage + 4u * (flushed & (TK_STAGE - 1u));
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(at) : "memory");
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(at + 16u) : "memory");
        uint4 *dst = reinterpret_cast<uint4 *>(tok + flushed);
        dst[0] = a, dst[1] = b;
        flushed += 8u;
    };
    int err = INF_OK, nlen = 288, ndist = 30;
    bool last = false;
    TkWin br;
    br.base = gz, br.n_bytes = 0, br.n_words = 0, br.mis = 0;
    br.w0 = br.w1 = br.w2 = br.p0 = br.p1 = 0, br.o = 0, br.wi = 3, br.fetched = br.landed = 0;
    br.ring = (uint32_t)__cvta_generic_to_shared(tk_smem + sizeof(TkFast) * TK_THREADS + 4 * TK_RING * threadIdx.x);
    uint32_t lfc[15 - TK_LBITS], dfc[15 - TK_DBITS], lc15 = 0, dc15 = 0;      // the codes that are longer than a table index (tk_decode_long)
#pragma unroll
    for (int k = 0; k < 15 - TK_LBITS; ++k) lfc[k] = 0;
#pragma unroll
    for (int k = 0; k < 15 - TK_DBITS; ++k) dfc[k] = 0;
    int state = TS_FETCH;
    auto fetch = [&]() {     // the next part that exists, or TS_DONE when the counter has run out
        for (;;) {
            const unsigned long long item64 = atomicAdd(next_item, 1ull);
            if (item64 >= n_items) { state = TS_DONE; return; }
            const uint32_t item = (uint32_t)item64;
            const int64_t blob = tk_blob_of(part_off, n_blob, item);
            const uint32_t p0 = part_off[blob], part = item - p0, np = part_off[blob + 1] - p0;
            R = rec + item;
            if (part > 0 && R->start_bit == TK_NONE) continue;
            next_start = TK_NONE;
            uint32_t next_j = np;
            for (uint32_t q = part + 1; q < np; ++q)
                if (R[q - part].start_bit != TK_NONE) { next_start = R[q - part].start_bit, next_j = q; break; }
            const uint64_t len64 = gz_off[blob + 1] - gz_off[blob], o0 = out_off[blob] - ob, cap64 = out_off[blob + 1] - out_off[blob];
            in = gz + gz_off[blob];
            err = INF_OK;
            if (cap64 > 0xFFFF0000ull || len64 > 0x0FFFFFFFull || len64 == 0) err = INF_ETOKENS;     // output positions and bit offsets are 32-bit
            cap = (uint32_t)cap64, in_len = (uint32_t)len64;
            br.set_blob(in, in_len);
            uint32_t p = 10;     // gzip header (RFC 1952): part 0
            if (!err && part == 0) {
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
            my_start = part == 0 ? 8u * p : R->start_bit;
            // the part's tokens: from its own place to the next part's
            const uint64_t tb = tk_tok_base(o0, (uint64_t)blob), budget = tk_tok_base(o0 + cap64, (uint64_t)blob + 1) - tb, nbits = 8ull * len64;
            tok_off = 0, tcap = 0;
            if (!err) {
                tok_off = tk_tok_off(budget, np, nbits, my_start, part);
                const uint32_t tok_end = next_start == TK_NONE ? (uint32_t)budget : tk_tok_off(budget, np, nbits, next_start, next_j);
                tcap = tok_end - tok_off - 8u;      // a group of TK_GROUP symbols may start at nt <= tcap
            }
            tok = tokens + tb + tok_off;
            r_isize = r_crc = r_final = 0;
            pos = nt = flushed = 0;
            last = false;
            if (err) br.w0 = br.w1 = br.w2 = br.p0 = br.p1 = 0, br.o = 0, br.wi = 3, br.mis = 0;
            else br.seek_bit(my_start);
            state = err ? TS_FINISHED : TS_HEADER;
            return;
        }
    };
    auto trailer = [&]() {   // CRC-32 and ISIZE (RFC 1952: size of the original input modulo 2^32); the resolver compares
        const uint64_t t = (br.consumed_bits() + 7) >> 3;
        if (t + 8 > in_len) err = INF_ETRUNC;
        else {
            const uint8_t *q = in + t;
            r_crc = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
            r_isize = (uint32_t)q[4] | ((uint32_t)q[5] << 8) | ((uint32_t)q[6] << 16) | ((uint32_t)q[7] << 24);
            r_final = 1u;
        }
        state = TS_FINISHED;
    };
    __syncwarp();
    for (;;) {
        // ---- parts that are finished are recorded and replaced
        if (__any_sync(0xFFFFFFFFu, state == TS_FINISHED || state == TS_FETCH)) {
            while (state == TS_FINISHED || state == TS_FETCH) {
                if (state == TS_FINISHED) {
                    while (flushed < nt) flush8();
                    PartRec r;
                    r.start_bit = my_start, r.end_bit = (uint32_t)br.consumed_bits(), r.n_tok = nt, r.out_len = pos;
                    r.status = err | (r_final ? TK_FINAL : 0);
                    r.isize = r_isize, r.crc = r_crc, r.tok_off = tok_off;
                    *R = r;
                }
                fetch();
            }
            __syncwarp();
        }
        // ---- block headers (per lane, divergent), then the table builds (the whole warp for one lane at a time)
        if (__any_sync(0xFFFFFFFFu, state == TS_HEADER)) {
            bool build = false;
            int type = 0;
            if (state == TS_HEADER && (br.past_end() || (next_start != TK_NONE && (uint32_t)br.consumed_bits() >= next_start))) {
                if (br.past_end()) err = INF_ETOKENS;
                state = TS_FINISHED;               // the next part takes over here (the resolver checks that it is exactly here)
            } else if (state == TS_HEADER) {
                const uint32_t hdr = br.bits(3);
                last = hdr & 1u;
                type = (int)(hdr >> 1);
                if (type == 0) {   // stored: to the byte boundary, LEN / NLEN; the bytes become literal tokens
                    br.drop((8u - (uint32_t)(br.consumed_bits() & 7u)) & 7u);
                    const uint32_t ln = br.peek();
                    br.drop(32);
                    const uint32_t len = ln & 0xFFFFu;
                    const uint32_t src = (uint32_t)(br.consumed_bits() >> 3);
                    if ((len ^ 0xFFFFu) != (ln >> 16) || (uint64_t)src + len > in_len) err = INF_ETRUNC;
                    else if ((uint64_t)pos + len > cap) err = INF_EOVERFLOW;
                    else if ((uint64_t)nt + len > tcap) err = INF_ETOKENS;   // (stored data is not what this decoder is for)
                    else {
                        for (uint32_t i = 0; i < len; ++i) {
                            put((uint32_t)in[src + i] << 9);
                            if ((int32_t)(nt - flushed) >= 8) flush8();
                        }
                        pos += len;
                        br.seek(src + len);
                    }
                    if (err) state = TS_FINISHED;   // else: the next block header, or the trailer (below)
                } else if (type == 3) {
                    err = INF_ECODE;
                    state = TS_FINISHED;
                } else if (type == 1) {
                    for (int s = 0; s < 288; ++s) T.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
                    for (int s = 0; s < 30; ++s) T.lens[288 + s] = 5;
                    nlen = 288, ndist = 30;
                    build = true;
                } else {
                    int hl = 0, hd = 0;
                    if (!tk_dynamic_header(br, T.lens, T.ctab, T.clen, &hl, &hd)) {
                        err = INF_ECODE;
                        state = TS_FINISHED;
                    } else {
                        nlen = hl, ndist = hd;
                        build = true;
                    }
                }
                if (state == TS_HEADER && !build && last) trailer();   // a stored block was the last one
            }
            __syncwarp();
            unsigned bm = __ballot_sync(0xFFFFFFFFu, build);
            while (bm) {
                const int j = __ffs(bm) - 1;
                bm &= bm - 1;
                const int nl = __shfl_sync(0xFFFFFFFFu, nlen, j), nd = __shfl_sync(0xFFFFFFFFu, ndist, j), ty = __shfl_sync(0xFFFFFFFFu, type, j);
                TkSlow &Tj = slows[j];
                TkFast &Fj = tabs[j];
                const bool ok1 = tk_build<TK_LBITS, false>(Tj.lens, nl, Tj.lcnt, Tj.lfirst, Tj.loffs, Tj.lsym, Fj.ltab, lane);
                const bool ok2 = tk_build<TK_DBITS, true>(Tj.lens + (ty == 1 ? 288 : nl), nd, Tj.dcnt, Tj.dfirst, Tj.doffs, Fj.dsym, Fj.dtab, lane);
                if (lane == j) {
                    if (ok1 && ok2) state = TS_SYMBOLS;
                    else err = INF_ECODE, state = TS_FINISHED;
                }
            }
            __syncwarp();
            if (build) {
#pragma unroll
                for (int k = 0; k < 15 - TK_LBITS; ++k) lfc[k] = (uint32_t)T.lfirst[TK_LBITS + 1 + k] | ((uint32_t)T.loffs[TK_LBITS + 1 + k] << 16);
#pragma unroll
                for (int k = 0; k < 15 - TK_DBITS; ++k) dfc[k] = (uint32_t)T.dfirst[TK_DBITS + 1 + k] | ((uint32_t)T.doffs[TK_DBITS + 1 + k] << 16);
                lc15 = T.lcnt[15], dc15 = T.dcnt[15];
            }
            continue;
        }
        if (__all_sync(0xFFFFFFFFu, state == TS_DONE)) break;
        if (__any_sync(0xFFFFFFFFu, state == TS_FINISHED)) continue;
        // ---- symbols: groups of TK_GROUP branch-free iterations that every lane executes in lock step.  A lane that meets
        // something the body does not handle (a code that is not in its tables, the end of a block) commits nothing and sits
        // out the rest of the group; the cases are sorted out between groups.
        for (;;) {
            if (state == TS_SYMBOLS && (nt > tcap || pos > cap || br.past_end())) err = INF_ETOKENS, state = TS_FINISHED;
            bool act = state == TS_SYMBOLS;
            br.top_up(act);
            const bool was = act;
#pragma unroll TK_UNROLL
            for (int it = 0; it < TK_GROUP; ++it) {
                const uint32_t lo = br.peek(), hi = br.peek_hi();
                const uint32_t e = F.ltab[lo & ((1u << TK_LBITS) - 1u)];
                const uint32_t n1 = e & 15u, ex = (e >> 4) & 7u, v = e >> 8;
                const bool is_len = (e & 128u) != 0u;
                const uint32_t len = v + 3u + ((lo >> (n1 - ex)) & ((1u << ex) - 1u));
                const uint32_t lo2 = __funnelshift_r(lo, hi, n1);             // the window behind the length symbol
                const uint32_t d = F.dtab[lo2 & ((1u << TK_DBITS) - 1u)];
                const uint32_t n2 = d & 31u, dx = (d >> 5) & 15u;
                const uint32_t dist = s_dbase[(d >> 9) & 31u] + ((lo2 >> (n2 - dx)) & ((1u << dx) - 1u));
                const bool commit = act && n1 != 0u && !(is_len && n2 == 0u);
                br.drop_fast(commit ? n1 + (is_len ? n2 : 0u) : 0u);
                if (commit) {
                    put(is_len ? ((dist << 9) | len) : (v << 9));
                    pos += is_len ? len : 1u;
                }
                act = commit;
            }
            bool eobf = false;
            if (__any_sync(0xFFFFFFFFu, was && !act)) {   // one symbol (and its distance) through the canonical decoder
                if (was && !act) {
                    uint32_t pk = br.peek();
                    const uint32_t e = F.ltab[pk & ((1u << TK_LBITS) - 1u)];
                    uint32_t sym = 300u, cl = 1;
                    if (e & 15u) {            // a symbol of the table whose distance code is not in the distance table
                        cl = (e & 15u) - ((e >> 4) & 7u);
                        sym = 1000u;          // (a length symbol: base and extra bits from the entry)
                    } else if (e & 0x8000u) sym = 256u, cl = (e >> 4) & 15u;
                    else if (e == 0u) {
                        uint32_t k;
                        cl = tk_decode_long<TK_LBITS>(pk, lfc, lc15, &k);
                        if (cl) {
                            if (k < (uint32_t)nlen) sym = T.lsym[k];
                        } else cl = 1;
                    }
                    const bool from_tab = sym == 1000u;
                    const bool is_lit = sym < 256u, is_len = from_tab || sym - 257u < 29u;
                    const uint32_t ex = from_tab ? (e >> 4) & 7u : s_lext[(sym - 257u) & 31u];
                    const uint32_t len = (from_tab ? (e >> 8) + 3u : s_lbase[(sym - 257u) & 31u]) + ((pk >> cl) & ((1u << ex) - 1u));
                    if (from_tab) sym = 257u;
                    br.drop(cl + (is_len ? ex : 0u));
                    uint32_t dist = 0;
                    if (is_len) {
                        pk = br.peek();
                        const uint32_t d = F.dtab[pk & ((1u << TK_DBITS) - 1u)];
                        uint32_t dl = 1, ds = 31u;
                        if (d & 31u) dl = (d & 31u) - ((d >> 5) & 15u), ds = d >> 9;
                        else if (d == 0u) {
                            uint32_t k2;
                            dl = tk_decode_long<TK_DBITS>(pk, dfc, dc15, &k2);
                            if (dl) {
                                if (k2 < (uint32_t)ndist) ds = F.dsym[k2];
                            } else dl = 1;
                        }
                        if (ds > 29u) err = INF_ECODE, ds = 0;
                        const uint32_t dx = s_dext[ds];
                        dist = s_dbase[ds] + ((pk >> dl) & ((1u << dx) - 1u));
                        br.drop(dl + dx);
                    }
                    if (sym == 256u) eobf = true;
                    else if (!(is_lit | is_len)) err = INF_ECODE;
                    else if (!err) {      // (nt <= tcap + TK_GROUP - 1 here, and the part's room is tcap + 8)
                        if (nt >= tcap + 8u) err = INF_ETOKENS;
                        else {
                            put(is_lit ? sym << 9 : ((dist << 9) | len));
                            pos += is_lit ? 1u : len;
                        }
                    }
                    if (err) state = TS_FINISHED;
                }
                __syncwarp();
            }
            // ---- full sectors of tokens leave (all lanes together; a group and its careful symbol add 9 tokens at most)
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if ((int32_t)(nt - flushed) >= 8) flush8();
            if (eobf && state == TS_SYMBOLS) {       // end of block
                if (!last) state = TS_HEADER;
                else if (pos > cap) err = INF_ETOKENS, state = TS_FINISHED;
                else trailer();
            }
            __syncwarp();
            if (__any_sync(0xFFFFFFFFu, state == TS_HEADER || state == TS_FINISHED) || !__any_sync(0xFFFFFFFFu, state == TS_SYMBOLS)) break;
        }
    }
}

// ---- phase R ------------------------------------------------------------------------------------------------------
constexpr int RS_WARPS = 4;
constexpr uint32_t RS_WIN = 4096;      // bytes of a blob's latest output kept in shared memory
__global__ void __launch_bounds__(RS_WARPS * 32)
inflate_resolve_kernel(const uint64_t *__restrict__ gz_off, int64_t n_blob, uint8_t *out_base, const uint64_t *__restrict__ out_off,
                       const uint32_t *__restrict__ part_off, const uint32_t *__restrict__ tokens, const PartRec *__restrict__ rec,
                       int32_t *__restrict__ status, uint32_t *__restrict__ crc_want) {
    __shared__ uint8_t rings[RS_WARPS][RS_WIN];
    const uint32_t lane = threadIdx.x & 31;
    const int64_t blob = (int64_t)blockIdx.x * RS_WARPS + (threadIdx.x >> 5);
    if (blob >= n_blob) return;
    uint8_t *ring = rings[threadIdx.x >> 5];
    const uint64_t o0 = out_off[blob], ob = out_off[0], cap = out_off[blob + 1] - o0;
    const uint32_t np = part_off[blob + 1] - part_off[blob];
    const PartRec *R = rec + part_off[blob];
    // ---- the chain: every part ended where the next one started, no part failed, the last one saw the final block, and the
    // sizes agree with the trailer and with the caller.  Part 0 alone decides real errors (it starts at a known place);
    // anything else that does not fit sends the blob to the one-pass decoder (INF_ETOKENS), which reports what is wrong.
    int st = INF_OK;
    {
        uint32_t prev_end = 0, total = 0, isize = 0, crc = 0;
        bool first = true, final_seen = false;
        for (uint32_t q = 0; q < np; ++q) {
            const PartRec r = R[q];
            if (q > 0 && r.start_bit == TK_NONE) continue;
            const int rs = r.status & 0xFFFF;
            if (rs != INF_OK) { st = (q == 0 && rs != INF_ETOKENS) ? rs : INF_ETOKENS; break; }
            if (!first && r.start_bit != prev_end) { st = INF_ETOKENS; break; }
            if (final_seen) { st = INF_ETOKENS; break; }      // a part behind the final block
            first = false;
            prev_end = r.end_bit;
            total += r.out_len;
            final_seen = (r.status & TK_FINAL) != 0;
            isize = r.isize, crc = r.crc;
        }
        if (st == INF_OK && !final_seen) st = INF_ETOKENS;
        if (st == INF_OK && (isize != total || (uint64_t)total != cap)) st = np > 1 ? INF_ETOKENS : INF_ESIZE;
        if (st == INF_OK && crc_want && lane == 0) crc_want[blob] = crc;
    }
    if (st != INF_OK) {
        if (lane == 0) status[blob] = st;
        return;
    }
    uint8_t *out = out_base + o0;
    const uint64_t t0 = tk_tok_base(o0 - ob, (uint64_t)blob);
    uint32_t P0 = 0;                                  // bytes written so far
    bool bad = false;                                 // a distance that reaches before the start of the output
    for (uint32_t q = 0; q < np; ++q) {
      if (q > 0 && R[q].start_bit == TK_NONE) continue;
      const uint32_t *tok = tokens + t0 + R[q].tok_off;
      const uint32_t nt = R[q].n_tok;
      uint32_t t_next = lane < nt ? __ldg(tok + lane) : 0u;
      for (uint32_t k0 = 0; k0 < nt; k0 += 32) {
        const uint32_t t = t_next;
        t_next = (k0 + 32 + lane < nt) ? __ldg(tok + k0 + 32 + lane) : 0u;   // the next batch's tokens are on their way
        const bool have = k0 + lane < nt;
        const uint32_t len_f = t & 511u, hi = t >> 9;
        const uint32_t n_out = !have ? 0u : (len_f == 0u ? 1u : len_f);
        uint32_t end = n_out;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, end, d);
            if (lane >= (uint32_t)d) end += v;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, end, 31);
        const uint32_t start = end - n_out;          // inside the batch
        if (have && len_f != 0u && (hi > P0 + start || hi == 0u)) bad = true;
        if ((uint64_t)P0 + total > cap) bad = true;
        if (__any_sync(0xFFFFFFFFu, bad)) break;
        // the batch's bytes, 32 at a time and in order: byte b belongs to the last token that starts at or before b — the
        // tokens mark their first bytes in a 32-bit mask (one warp-wide OR), and a lane counts the marks up to its own byte.
        // A match byte is copied from b - distance: before the round (written: the last RS_WIN bytes also sit in a ring in
        // shared memory, so the serial chain round -> round does not wait for the L2) or inside the round itself — then
        // the lane follows the chain of sources (each step moves at least one byte back) until it leaves the round or
        // ends at a literal.
        const uint32_t rounds = (total + 31u) >> 5;
        uint32_t n_before = 0;                       // tokens of the batch that start before the round
        for (uint32_t r = 0; r < rounds; ++r) {
            const uint32_t r0 = r * 32u, b = r0 + lane;
            const uint32_t s_rel = start - r0;
            const uint32_t marks = __reduce_or_sync(0xFFFFFFFFu, (have && s_rel < 32u) ? 1u << s_rel : 0u);
            const uint32_t owner = n_before + (uint32_t)__popc(marks & ((2u << lane) - 1u)) - 1u;
            n_before += (uint32_t)__popc(marks);
            const uint32_t o_tok = __shfl_sync(0xFFFFFFFFu, t, owner & 31u);
            const bool valid = b < total, lit = (o_tok & 511u) == 0u;
            // src: where the byte comes from, relative to the batch (negative: before it); literals carry their value
            const uint32_t dist = o_tok >> 9;
            int32_t src = (int32_t)b - (int32_t)dist;
            uint32_t val = dist;
            bool done = !valid || lit;
            while (__any_sync(0xFFFFFFFFu, !done && src >= (int32_t)r0)) {
                const uint32_t from = (uint32_t)(src - (int32_t)r0) & 31u;
                const int32_t s2 = __shfl_sync(0xFFFFFFFFu, src, from);
                const uint32_t v2 = __shfl_sync(0xFFFFFFFFu, val, from);
                const bool d2 = __shfl_sync(0xFFFFFFFFu, (int)done, from) != 0;
                if (!done && src >= (int32_t)r0) {
                    if (d2) val = v2, done = true;     // the source lane holds a literal (or a value it has already resolved)
                    else src = s2;                     // its source is further back
                }
            }
            if (valid && !done) {
                const uint32_t back = b - (uint32_t)src;               // how far behind this byte the source lies
                val = back <= RS_WIN - 32u ? ring[(P0 + (uint32_t)src) & (RS_WIN - 1u)] : out[(int64_t)P0 + src];
            }
            __syncwarp();
            if (valid) {
                ring[(P0 + b) & (RS_WIN - 1u)] = (uint8_t)val;
                out[P0 + b] = (uint8_t)val;
            }
            __syncwarp();
        }
        P0 += total;
      }
      if (__any_sync(0xFFFFFFFFu, bad)) break;
    }
    const bool any_bad = __any_sync(0xFFFFFFFFu, bad);   // (the one-pass decoder names the error)
    if (lane == 0) status[blob] = any_bad ? INF_ETOKENS : INF_OK;
}

// piece_off[b] = pieces of CRC_PIECE bytes before blob b (one CTA; blobs in chunks of 1024)
constexpr uint32_t CRC_PIECE = 1024;
__global__ void __launch_bounds__(1024)
crc32_piece_offsets_kernel(const uint64_t *__restrict__ out_off, int64_t n_blob, uint64_t *__restrict__ piece_off) {
    __shared__ uint64_t ws[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_blob; base += 1024) {
        const int64_t b = base + threadIdx.x;
        const uint64_t v = b < n_blob ? (out_off[b + 1] - out_off[b] + CRC_PIECE - 1) / CRC_PIECE : 0;
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
        if (b < n_blob) piece_off[b] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) piece_off[n_blob] = carry;
}

// ---- CRC-32 (IEEE 802.3, reflected 0xEDB88320: RFC 1952 section 8) of every blob's output ----------------------------
// One thread per 1024-byte piece, byte-wise through a 256-entry table in shared memory; then one thread per blob folds the
// piece values in order: crc(A || B) = shift(crc(A), |B|) ^ crc(B) with the all-ones pre/post conditioning handled as zlib's
// crc32_combine does, the shift by |B| zero bytes being a multiplication by x^(8|B|) mod P (square-and-multiply over the
// bits of |B|, carry-less arithmetic in 32-bit registers).
__device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b) {   // reflected polynomials, modulo 0xEDB88320
    uint32_t p = 0;
    for (int i = 0; i < 32; ++i) {
        if (a & 0x80000000u) p ^= b;      // bit 31 of a reflected polynomial is the coefficient of x^0
        a <<= 1;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);   // b *= x
    }
    return p;
}
__device__ uint32_t gf2_x_pow_8n(uint64_t n_bytes) {   // x^(8 n) mod P, reflected
    uint32_t r = 0x80000000u;                          // 1
    uint32_t sq = 0x00800000u;                         // x^8: coefficient of x^8 is bit 31 - 8
    while (n_bytes) {
        if (n_bytes & 1ull) r = gf2_mulmod(r, sq);
        sq = gf2_mulmod(sq, sq);
        n_bytes >>= 1;
    }
    return r;
}

__global__ void __launch_bounds__(256)
crc32_pieces_kernel(const uint8_t *__restrict__ data, const uint64_t *__restrict__ out_off, int64_t n_blob,
                    const uint64_t *__restrict__ piece_off /* n_blob + 1 */, uint32_t *__restrict__ piece_crc) {
    __shared__ uint32_t tab[4][256];     // slicing-by-4: tab[k][v] = the register after byte v and k zero bytes
    {
        uint32_t c = threadIdx.x;
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0xEDB88320u : 0u);
        tab[0][threadIdx.x] = c;
    }
    __syncthreads();
    {
        uint32_t c = tab[0][threadIdx.x];
        for (int k = 1; k < 4; ++k) {
            c = tab[0][c & 0xFFu] ^ (c >> 8);
            tab[k][threadIdx.x] = c;
        }
    }
    __syncthreads();
    const uint64_t n_piece = piece_off[n_blob];
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_piece; p += (uint64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = n_blob;     // the blob of piece p
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (piece_off[mid] <= p) lo = mid; else hi = mid;
        }
        const uint64_t b0 = out_off[lo] + (p - piece_off[lo]) * CRC_PIECE, b1 = min(b0 + CRC_PIECE, out_off[lo + 1]);
        uint32_t c = 0;                  // raw register value (no pre / post inversion): pieces combine linearly
        uint64_t i = b0;
        for (; i < b1 && (reinterpret_cast<uintptr_t>(data + i) & 15u); ++i) c = tab[0][(c ^ data[i]) & 0xFFu] ^ (c >> 8);
        for (; i + 16 <= b1; i += 16) {
            const uint4 q = *reinterpret_cast<const uint4 *>(data + i);
            c ^= q.x;
            c = tab[3][c & 0xFFu] ^ tab[2][(c >> 8) & 0xFFu] ^ tab[1][(c >> 16) & 0xFFu] ^ tab[0][c >> 24];
            c ^= q.y;
            c = tab[3][c & 0xFFu] ^ tab[2][(c >> 8) & 0xFFu] ^ tab[1][(c >> 16) & 0xFFu] ^ tab[0][c >> 24];
            c ^= q.z;
            c = tab[3][c & 0xFFu] ^ tab[2][(c >> 8) & 0xFFu] ^ tab[1][(c >> 16) & 0xFFu] ^ tab[0][c >> 24];
            c ^= q.w;
            c = tab[3][c & 0xFFu] ^ tab[2][(c >> 8) & 0xFFu] ^ tab[1][(c >> 16) & 0xFFu] ^ tab[0][c >> 24];
        }
        for (; i < b1; ++i) c = tab[0][(c ^ data[i]) & 0xFFu] ^ (c >> 8);
        piece_crc[p] = c;
    }
}

__global__ void __launch_bounds__(128)
crc32_fold_kernel(const uint64_t *__restrict__ out_off, int64_t n_blob, const uint64_t *__restrict__ piece_off,
                  const uint32_t *__restrict__ piece_crc, const uint32_t *__restrict__ crc_want, int32_t *__restrict__ status) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blob || status[b] != INF_OK) return;
    const uint64_t len = out_off[b + 1] - out_off[b], np = piece_off[b + 1] - piece_off[b];
    const uint32_t xfull = gf2_x_pow_8n(CRC_PIECE);
    // register after the whole message when it starts from all ones: the all-ones state advanced over len bytes, plus the
    // pieces (each started from zero) advanced over what follows them
    uint32_t c = 0;
    for (uint64_t k = 0; k < np; ++k) {
        const uint64_t plen = (k + 1 < np) ? CRC_PIECE : len - k * CRC_PIECE;
        c = gf2_mulmod(c, plen == CRC_PIECE ? xfull : gf2_x_pow_8n(plen)) ^ piece_crc[piece_off[b] + k];
    }
    c ^= gf2_mulmod(0xFFFFFFFFu, gf2_x_pow_8n(len));
    c ^= 0xFFFFFFFFu;
    if (c != crc_want[b]) status[b] = INF_ECRC;
}

// Layout of the decoder's scratch (pd.scratch[7]): tokens | part records | part_off | per-lane header scratch | crc_want |
// piece_off | piece_crc | counter.  Every offset grows with n_blob and total_out, so a buffer reserved for a whole batch also
// holds any piece of it.
struct Gz2Layout {
    uint64_t p_cap, n_lanes, n_piece_max;
    size_t o_rec, o_poffs, o_slow, o_crc, o_poff, o_pcrc, bytes;
};
static int32_t gz2_layout(PerDevice &pd, int64_t n_blob, uint64_t total_out, int decoder, bool check_crc, Gz2Layout &L, int &n_sm_out) {
    static int n_sm = 0;
    if (!n_sm) CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, pd.device));
    n_sm_out = n_sm;
    // part records: sequence compresses ~3:1, so ranges of TK_RANGE compressed bytes are ~total_out / (3 TK_RANGE); room for a
    // ratio of 2, and the range grows when the batch compresses worse (tk_range_bytes)
    L.p_cap = std::min<uint64_t>(total_out / (2ull * TK_RANGE) + 2ull * (uint64_t)n_blob + 64, 0x7FFFFFF0ull);
    L.n_lanes = (uint64_t)n_sm * TK_THREADS;
    const uint64_t n_tokens = decoder == 0 ? (total_out >> 1) + 64ull * (uint64_t)(n_blob + 1) + 64 : 0;
    L.n_piece_max = total_out / CRC_PIECE + (uint64_t)n_blob + 1;
    L.o_rec = ((size_t)n_tokens * 4 + 31) & ~(size_t)31;
    L.o_poffs = L.o_rec + (decoder == 0 ? (size_t)L.p_cap * sizeof(PartRec) : 0);
    L.o_slow = (L.o_poffs + (decoder == 0 ? (size_t)(n_blob + 1) * 8 : 0) + 15) & ~(size_t)15;
    L.o_crc = L.o_slow + (decoder == 0 ? (size_t)L.n_lanes * sizeof(TkSlow) : 0);
    L.o_poff = (L.o_crc + (size_t)n_blob * 4 + 15) & ~(size_t)15;
    L.o_pcrc = L.o_poff + (size_t)(n_blob + 1) * 8;
    L.bytes = L.o_pcrc + (check_crc ? (size_t)L.n_piece_max * 4 : 0) + 64;
    return GAMS_CUDA_OK;
}

int32_t gunzip2_scratch_bytes(PerDevice &pd, int64_t n_blob, uint64_t total_out, int decoder, bool check_crc, size_t *bytes) {
    Gz2Layout L;
    int n_sm;
    int32_t rc = gz2_layout(pd, n_blob, total_out, decoder, check_crc, L, n_sm);
    if (rc) return rc;
    *bytes = (L.bytes + 255) & ~(size_t)255;
    return GAMS_CUDA_OK;
}

int32_t launch_gunzip2(PerDevice &pd, cudaStream_t st, const uint8_t *d_gz, const uint64_t *d_gz_off, int64_t n_blob, uint8_t *d_out,
                       const uint64_t *d_out_off, int32_t *d_status, uint64_t total_out, int decoder, bool check_crc, int n_arrive,
                       const int64_t *arrive_end, const cudaEvent_t *arrive_ev, cudaEvent_t decoded_ev) {
    if (n_blob == 0) return GAMS_CUDA_OK;
    if (n_blob >= (1ll << 30)) {
        set_error("gunzip: too many blobs for one launch");
        return GAMS_CUDA_EINVAL;
    }
    Gz2Layout L;
    int n_sm;
    int32_t rc;
    if ((rc = gz2_layout(pd, n_blob, total_out, decoder, check_crc, L, n_sm))) return rc;
    const uint64_t p_cap = L.p_cap, n_piece_max = L.n_piece_max;
    const size_t o_rec = L.o_rec, o_poffs = L.o_poffs, o_slow = L.o_slow, o_crc = L.o_crc, o_poff = L.o_poff, o_pcrc = L.o_pcrc, bytes = L.bytes;
    if ((rc = pd.scratch[7].reserve(bytes))) return rc;
    uint8_t *base = pd.scratch[7].as<uint8_t>();
    uint32_t *d_tok = reinterpret_cast<uint32_t *>(base), *d_crc = reinterpret_cast<uint32_t *>(base + o_crc),
             *d_pcrc = reinterpret_cast<uint32_t *>(base + o_pcrc), *d_parts = reinterpret_cast<uint32_t *>(base + o_poffs), *d_groups = d_parts + (n_blob + 1);
    PartRec *d_rec = reinterpret_cast<PartRec *>(base + o_rec);
    TkSlow *d_slow = reinterpret_cast<TkSlow *>(base + o_slow);
    uint64_t *d_poff = reinterpret_cast<uint64_t *>(base + o_poff);
    unsigned long long *d_counter = reinterpret_cast<unsigned long long *>(base + ((bytes - 16) & ~(size_t)7));
    if (decoder == 0) {
        const size_t smem = TK_SMEM;
        CUDA_TRY(cudaFuncSetAttribute(inflate_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        inflate_parts_kernel<<<1, 1024, 0, st>>>(d_gz_off, d_out_off, n_blob, p_cap, d_parts, d_groups);
        // the header search needs nothing but the compressed bytes: it runs on each stretch of blobs as soon as that stretch
        // has arrived (arrive_ev[i] fires when blobs < arrive_end[i] are on the device), while the rest is still on the link
        int64_t b0 = 0;
        for (int i = 0; i <= n_arrive; ++i) {
            const int64_t b1 = i < n_arrive ? std::min<int64_t>(arrive_end[i], n_blob) : n_blob;
            if (i < n_arrive) CUDA_TRY(cudaStreamWaitEvent(st, arrive_ev[i], 0));
            if (b1 <= b0) continue;
            const uint64_t share = (p_cap / SP_GROUP) * (uint64_t)(b1 - b0) / (uint64_t)n_blob + (uint64_t)(b1 - b0);
            const unsigned s_blocks = (unsigned)std::min<uint64_t>((share + SP_WARPS) / SP_WARPS, (uint64_t)n_sm * 16);
            inflate_split_kernel<<<s_blocks, SP_WARPS * 32, 0, st>>>(d_gz, d_gz_off, n_blob, d_parts, d_groups, d_rec, b0, b1);
            b0 = b1;
        }
        trace_stage("gunzip: header search", st, true);
        CUDA_TRY(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
        const unsigned t_blocks = (unsigned)std::min<uint64_t>((p_cap + TK_THREADS - 1) / TK_THREADS, (uint64_t)n_sm);
        inflate_tokens_kernel<<<t_blocks, TK_THREADS, smem, st>>>(d_gz, d_gz_off, n_blob, d_out_off, d_parts, d_tok, d_rec, d_slow, d_counter);
        trace_stage("gunzip: tokens", st, true);
        inflate_resolve_kernel<<<(unsigned)((n_blob + RS_WARPS - 1) / RS_WARPS), RS_WARPS * 32, 0, st>>>(d_gz_off, n_blob, d_out, d_out_off, d_parts,
                                                                                                     d_tok, d_rec, d_status, check_crc ? d_crc : nullptr);
        trace_stage("gunzip: bytes", st, true);
        // blobs that did not fit their token budget, or whose chain of parts did not close: the one-pass kernel
        if ((rc = launch_gunzip_only(st, d_gz, d_gz_off, n_blob, d_out, d_out_off, d_status, INF_ETOKENS, check_crc ? d_crc : nullptr))) return rc;
    } else {
        if (n_arrive) CUDA_TRY(cudaStreamWaitEvent(st, arrive_ev[n_arrive - 1], 0));
        if ((rc = launch_gunzip_only(st, d_gz, d_gz_off, n_blob, d_out, d_out_off, d_status, -1, check_crc ? d_crc : nullptr))) return rc;
    }
    if (decoded_ev) CUDA_TRY(cudaEventRecord(decoded_ev, st));     // the bytes are final; what follows only reads them
    if (check_crc) {
        crc32_piece_offsets_kernel<<<1, 1024, 0, st>>>(d_out_off, n_blob, d_poff);
        const unsigned pb = (unsigned)std::min<uint64_t>((n_piece_max + 255) / 256, 148ull * 16);
        crc32_pieces_kernel<<<pb, 256, 0, st>>>(d_out, d_out_off, n_blob, d_poff, d_pcrc);
        crc32_fold_kernel<<<(unsigned)((n_blob + 127) / 128), 128, 0, st>>>(d_out_off, n_blob, d_poff, d_pcrc, d_crc, d_status);
    }
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
