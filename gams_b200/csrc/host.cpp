// host.cpp — host-side mirrors of the text/format glue around the hot path (no CUDA in here).
//
//  * gams_cuda_wave_text: the emit part of wave::proc_ctg, src/cmd_gams/wave.rs:157-212, including
//    merge_ints (wave.rs:217-252: connected components of same-sign windows that intersect, with the
//    reference's size/intersect >= coverage test) and Rust's `{}` formatting of f32.
//  * gams_cuda_lapper_blob_view: zero-copy slices of starts[]/stops[] inside a bincode 1.3
//    serialisation of rust_lapper::Lapper<u32,String> (src/libs/redis.rs:299-300, 316-317).
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/gams_cuda.h"

#define GAMS_API extern "C" __attribute__((visibility("default")))

namespace gams {
void set_error(const char *fmt, ...);

// Rust `{}` for f32: shortest digits that round-trip, positional notation, "NaN"/"inf".
static void fmt_f32(std::string &out, float x) {
    if (std::isnan(x)) { out += "NaN"; return; }
    if (std::isinf(x)) { out += x > 0 ? "inf" : "-inf"; return; }
    char buf[128];
    auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::fixed);
    out.append(buf, r.ptr);
}

static void fmt_runlist(std::string &out, int64_t a, int64_t b) {
    out += std::to_string(a);
    if (a != b) {
        out += '-';
        out += std::to_string(b);
    }
}

struct Dsu {
    std::vector<int> p;
    explicit Dsu(int n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int find(int x) {
        while (p[x] != x) x = p[x] = p[p[x]];
        return x;
    }
    void unite(int a, int b) {
        a = find(a);
        b = find(b);
        if (a != b) p[b] = a;
    }
};

// merge_ints on windows given in start order; fills merged[k] (text) for windows that own an edge.
static void merge_ints(const std::vector<int64_t> &idx, int64_t chr_start, int32_t size, int32_t step, float coverage,
                       std::unordered_map<int64_t, std::string> &merged) {
    const int n = (int)idx.size();
    if (n < 2) return;
    Dsu dsu(n);
    std::vector<char> in_graph(n, 0);
    const float fsz = (float)size;
    for (int i = 0; i < n; ++i) {
        const int64_t si = chr_start + idx[i] * step, ei = si + size - 1;
        for (int j = i + 1; j < n; ++j) {
            const int64_t sj = chr_start + idx[j] * step;
            if (sj > ei) break;  // sorted by start: nothing later intersects window i
            const float isz = (float)(int32_t)(ei - sj + 1);  // intersect.size() as f32
            if (fsz / isz >= coverage) {                       // cov_i == cov_j: all windows share one size
                in_graph[i] = in_graph[j] = 1;
                dsu.unite(i, j);
            }
        }
    }
    std::unordered_map<int, std::pair<int64_t, int64_t>> span;  // root -> union of members (always one run)
    for (int i = 0; i < n; ++i) {
        if (!in_graph[i]) continue;
        const int64_t s = chr_start + idx[i] * step, e = s + size - 1;
        auto it = span.find(dsu.find(i));
        if (it == span.end()) span.emplace(dsu.find(i), std::make_pair(s, e));
        else {
            it->second.first = std::min(it->second.first, s);
            it->second.second = std::max(it->second.second, e);
        }
    }
    for (int i = 0; i < n; ++i) {
        if (!in_graph[i]) continue;
        const auto &sp = span[dsu.find(i)];
        std::string t;
        fmt_runlist(t, sp.first, sp.second);
        merged.emplace(idx[i], std::move(t));
    }
}
}  // namespace gams

using namespace gams;

GAMS_API int32_t gams_cuda_wave_text(const char *chr_id, int32_t chr_start, int32_t size, int32_t step, int64_t n,
                                     const float *gc, const int8_t *signal, float coverage, int32_t is_signal,
                                     char **text, size_t *text_len) {
    if (!chr_id || !text || n < 0 || size <= 0 || step <= 0 || (n > 0 && (!gc || !signal))) {
        set_error("wave_text: bad argument");
        return GAMS_CUDA_EINVAL;
    }
    std::string out;
    auto line = [&](bool plus, const std::string &rl, int64_t k) {
        out += chr_id;
        if (plus) out += "(+)";
        out += ':';
        out += rl;
        out += '\t';
        fmt_f32(out, gc[k]);
        out += '\t';
        out += std::to_string((int)signal[k]);
        out += '\n';
    };
    if (is_signal) {  // wave.rs:158-168
        for (int64_t k = 0; k < n; ++k) {
            std::string rl;
            fmt_runlist(rl, (int64_t)chr_start + k * step, (int64_t)chr_start + k * step + size - 1);
            line(false, rl, k);
        }
    } else {  // wave.rs:170-211
        std::vector<int64_t> crest, trough;
        for (int64_t k = 0; k < n; ++k) {
            if (signal[k] == 1) crest.push_back(k);
            else if (signal[k] == -1) trough.push_back(k);
        }
        std::unordered_map<int64_t, std::string> merged;
        merge_ints(crest, chr_start, size, step, coverage, merged);
        merge_ints(trough, chr_start, size, step, coverage, merged);
        std::unordered_set<std::string> seen;
        for (int64_t k = 0; k < n; ++k) {
            if (signal[k] == 0) continue;
            auto it = merged.find(k);
            if (it != merged.end()) {
                if (seen.insert(it->second).second) line(true, it->second, k);
            } else {
                std::string rl;
                fmt_runlist(rl, (int64_t)chr_start + k * step, (int64_t)chr_start + k * step + size - 1);
                line(false, rl, k);
            }
        }
    }
    char *buf = (char *)malloc(out.size() + 1);
    if (!buf) {
        set_error("wave_text: out of memory");
        return GAMS_CUDA_ENOMEM;
    }
    memcpy(buf, out.data(), out.size());
    buf[out.size()] = 0;
    *text = buf;
    if (text_len) *text_len = out.size();
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_wave_peaks_text(const char *chr_id, int32_t chr_start, int32_t size, int32_t step, int64_t n_pk,
                                           const uint32_t *pk_first, const uint32_t *pk_last, const float *pk_gc,
                                           const int8_t *pk_signal, char **text, size_t *text_len) {
    if (!chr_id || !text || n_pk < 0 || size <= 0 || step <= 0 || (n_pk > 0 && (!pk_first || !pk_last || !pk_gc || !pk_signal))) {
        set_error("wave_peaks_text: bad argument");
        return GAMS_CUDA_EINVAL;
    }
    std::string out;
    for (int64_t i = 0; i < n_pk; ++i) {
        out += chr_id;
        if (pk_last[i] != pk_first[i]) out += "(+)";   // a merged peak (wave.rs:196-203)
        out += ':';
        fmt_runlist(out, (int64_t)chr_start + (int64_t)pk_first[i] * step, (int64_t)chr_start + (int64_t)pk_last[i] * step + size - 1);
        out += '\t';
        fmt_f32(out, pk_gc[i]);
        out += '\t';
        out += std::to_string((int)pk_signal[i]);
        out += '\n';
    }
    char *buf = (char *)malloc(out.size() + 1);
    if (!buf) {
        set_error("wave_peaks_text: out of memory");
        return GAMS_CUDA_ENOMEM;
    }
    memcpy(buf, out.data(), out.size());
    buf[out.size()] = 0;
    *text = buf;
    if (text_len) *text_len = out.size();
    return GAMS_CUDA_OK;
}

GAMS_API void gams_cuda_free(void *p) { free(p); }

// bincode 1.3 default config: little-endian, fixed-width ints, u64 length prefixes.
// rust-lapper 1.1.0 field order: intervals: Vec<Interval{start:u32, stop:u32, val:String}>,
// starts: Vec<u32>, stops: Vec<u32>, max_len: u32, cov: Option<u32>, overlaps_merged: bool.
GAMS_API int32_t gams_cuda_lapper_blob_view(const uint8_t *blob, size_t blob_len, uint64_t *n_iv,
                                            const uint8_t **starts, const uint8_t **stops, uint32_t *max_len) {
    if (!blob || !n_iv || !starts || !stops) {
        set_error("lapper_blob_view: NULL argument");
        return GAMS_CUDA_EINVAL;
    }
    size_t pos = 0;
    auto need = [&](size_t k) { return pos + k <= blob_len && pos + k >= pos; };
    auto rd64 = [&](uint64_t &v) {
        if (!need(8)) return false;
        memcpy(&v, blob + pos, 8);
        pos += 8;
        return true;
    };
    uint64_t n = 0, n2 = 0, n3 = 0;
    if (!rd64(n) || n > blob_len / 16) goto bad;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t sl = 0;
        if (!need(8)) goto bad;
        pos += 8;  // start, stop
        if (!rd64(sl) || !need((size_t)sl)) goto bad;
        pos += (size_t)sl;
    }
    if (!rd64(n2) || n2 != n || !need((size_t)n * 4)) goto bad;
    *starts = blob + pos;
    pos += (size_t)n * 4;
    if (!rd64(n3) || n3 != n || !need((size_t)n * 4)) goto bad;
    *stops = blob + pos;
    pos += (size_t)n * 4;
    if (!need(4)) goto bad;
    if (max_len) memcpy(max_len, blob + pos, 4);
    pos += 4;
    if (!need(1)) goto bad;
    if (blob[pos] == 1) pos += 5;  // Some(cov)
    else if (blob[pos] == 0) pos += 1;
    else goto bad;
    if (!need(1) || pos + 1 != blob_len) goto bad;
    *n_iv = n;
    return GAMS_CUDA_OK;
bad:
    set_error("lapper_blob_view: not a bincode Lapper<u32,String> blob (offset %zu of %zu)", pos, blob_len);
    return GAMS_CUDA_EPARSE;
}
