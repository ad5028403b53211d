// Host-side fast path for `{prefix}.variant.json` (reference writer/reader:
// graphkir/hisat2.py:847-866).  At 200k read pairs the file is ~200 MB, most of it the raw SAM text
// of every pair; json.load + dataclass construction + per-object packing then cost seconds per
// sample while the GPU types it in a fraction of a millisecond.  This scanner walks the JSON text
// once and extracts only what the typing path needs from every element of "reads":
//     backbone (interned), multiple, and the four variant-id lists lpv / lnv / rpv / rnv (interned),
// as CSR arrays, skipping l_sam / r_sam without copying them.  The "variants" array is returned as a
// byte span for the caller to parse (it is small).  Pure host code; no CUDA calls.
#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gk_common.cuh"

namespace {

// String interning: open addressing over FNV-1a hashes; the ids of a sample repeat ~10^2 times each.
struct StringTable {
    std::vector<std::string> items;
    std::vector<int32_t> slots = std::vector<int32_t>(1024, -1);
    std::vector<uint64_t> hashes;

    static uint64_t hash(const char* s, size_t n) {
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)s[i]) * 1099511628211ull;
        return h;
    }
    void grow() {
        std::vector<int32_t> bigger(slots.size() * 2, -1);
        const size_t mask = bigger.size() - 1;
        for (size_t id = 0; id < items.size(); ++id) {
            size_t i = hashes[id] & mask;
            while (bigger[i] >= 0) i = (i + 1) & mask;
            bigger[i] = (int32_t)id;
        }
        slots.swap(bigger);
    }
    int32_t intern(const char* s, size_t n) {
        const uint64_t h = hash(s, n);
        size_t mask = slots.size() - 1;
        size_t i = h & mask;
        while (slots[i] >= 0) {
            const int32_t id = slots[i];
            if (hashes[id] == h && items[id].size() == n && memcmp(items[id].data(), s, n) == 0) return id;
            i = (i + 1) & mask;
        }
        const int32_t id = (int32_t)items.size();
        items.emplace_back(s, n);
        hashes.push_back(h);
        slots[i] = id;
        if (items.size() * 2 > slots.size()) grow();
        return id;
    }
    int32_t intern(const std::string& s) { return intern(s.data(), s.size()); }
};

struct JsonScan {
    std::vector<int32_t> backbone, multiple;
    std::vector<int64_t> off[4];
    std::vector<int32_t> idx[4];
    StringTable ids, genes;
    int64_t variants_begin = -1, variants_end = -1;
};

struct Parser {
    const char* p;
    const char* end;
    const char* begin;
    const char* error = nullptr;

    bool fail(const char* msg) {
        if (!error) error = msg;
        return false;
    }
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    }
    bool expect(char c) {
        ws();
        if (p < end && *p == c) {
            ++p;
            return true;
        }
        return fail("unexpected character");
    }
    bool skip_string() {                      // p at the opening quote
        ++p;
        while (p < end) {                     // next quote (libc scans 16+ bytes per step) ...
            const char* q = static_cast<const char*>(memchr(p, '"', (size_t)(end - p)));
            if (q == nullptr) break;
            const char* b = q;                // ... that is not escaped: an even run of backslashes before it
            while (b > p && b[-1] == '\\') --b;
            p = q + 1;
            if (((q - b) & 1) == 0) return true;
        }
        p = end;
        return fail("unterminated string");
    }
    static void put_utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) {
            out.push_back((char)cp);
        } else if (cp < 0x800) {
            out.push_back((char)(0xC0 | (cp >> 6)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        } else if (cp < 0x10000) {
            out.push_back((char)(0xE0 | (cp >> 12)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            out.push_back((char)(0xF0 | (cp >> 18)));
            out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    bool hex4(unsigned& v) {
        if (end - p < 4) return fail("truncated \\u escape");
        v = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else return fail("bad \\u escape");
        }
        return true;
    }
    bool string(std::string& out) {           // p at the opening quote
        out.clear();
        ++p;
        const char* run = p;
        while (p < end) {
            const char c = *p;
            if (c == '"') {
                out.append(run, p);
                ++p;
                return true;
            }
            if (c != '\\') {
                ++p;
                continue;
            }
            out.append(run, p);
            if (++p >= end) break;
            const char e = *p++;
            switch (e) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    unsigned cp;
                    if (!hex4(cp)) return false;
                    if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        p += 2;
                        unsigned lo;
                        if (!hex4(lo)) return false;
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    put_utf8(out, cp);
                    break;
                }
                default: return fail("bad escape");
            }
            run = p;
        }
        return fail("unterminated string");
    }
    bool skip_value() {
        ws();
        if (p >= end) return fail("unexpected end");
        const char c = *p;
        if (c == '"') return skip_string();
        if (c == '{' || c == '[') {
            const char close = c == '{' ? '}' : ']';
            ++p;
            ws();
            if (p < end && *p == close) {
                ++p;
                return true;
            }
            while (true) {
                if (c == '{') {
                    ws();
                    if (p >= end || *p != '"') return fail("expected a key");
                    if (!skip_string() || !expect(':')) return false;
                }
                if (!skip_value()) return false;
                ws();
                if (p < end && *p == ',') {
                    ++p;
                    continue;
                }
                return expect(close);
            }
        }
        // number, true, false, null
        while (p < end && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') ++p;
        return true;
    }
    bool integer(int32_t& out) {
        ws();
        char* stop = nullptr;
        errno = 0;
        const long v = strtol(p, &stop, 10);
        if (stop == p || errno) return fail("expected an integer");
        p = stop;
        out = (int32_t)v;
        return true;
    }
    bool id_list(JsonScan& scan, int which, std::string& tmp) {
        if (!expect('[')) return false;
        ws();
        if (p < end && *p == ']') {
            ++p;
            return true;
        }
        while (true) {
            ws();
            if (p >= end || *p != '"') return fail("variant ids must be strings");
            const char* q = static_cast<const char*>(memchr(p + 1, '"', (size_t)(end - p - 1)));
            if (q != nullptr && memchr(p + 1, '\\', (size_t)(q - p - 1)) == nullptr) {    // plain: no copy
                scan.idx[which].push_back(scan.ids.intern(p + 1, (size_t)(q - p - 1)));
                p = q + 1;
            } else {
                if (!string(tmp)) return false;
                scan.idx[which].push_back(scan.ids.intern(tmp));
            }
            ws();
            if (p < end && *p == ',') {
                ++p;
                continue;
            }
            return expect(']');
        }
    }
    bool read(JsonScan& scan, std::string& key, std::string& tmp) {
        if (!expect('{')) return false;
        int32_t backbone = scan.genes.intern("");      // PairRead defaults (hisat2.py:41-52)
        int32_t multiple = 1;
        ws();
        if (p < end && *p == '}') {
            ++p;
        } else {
            while (true) {
                ws();
                if (p >= end || *p != '"') return fail("expected a key");
                if (!string(key) || !expect(':')) return false;
                ws();
                if (key == "backbone") {
                    if (p >= end || *p != '"') return fail("backbone must be a string");
                    if (!string(tmp)) return false;
                    backbone = scan.genes.intern(tmp);
                } else if (key == "multiple") {
                    if (!integer(multiple)) return false;
                } else if (key == "lpv") {
                    if (!id_list(scan, 0, tmp)) return false;
                } else if (key == "lnv") {
                    if (!id_list(scan, 1, tmp)) return false;
                } else if (key == "rpv") {
                    if (!id_list(scan, 2, tmp)) return false;
                } else if (key == "rnv") {
                    if (!id_list(scan, 3, tmp)) return false;
                } else if (!skip_value()) {
                    return false;
                }
                ws();
                if (p < end && *p == ',') {
                    ++p;
                    continue;
                }
                if (!expect('}')) return false;
                break;
            }
        }
        scan.backbone.push_back(backbone);
        scan.multiple.push_back(multiple);
        for (int w = 0; w < 4; ++w) scan.off[w].push_back((int64_t)scan.idx[w].size());
        return true;
    }
    bool document(JsonScan& scan) {
        std::string key, tmp;
        for (int w = 0; w < 4; ++w) scan.off[w].push_back(0);
        if (!expect('{')) return false;
        ws();
        if (p < end && *p == '}') return true;
        while (true) {
            ws();
            if (p >= end || *p != '"') return fail("expected a key");
            if (!string(key) || !expect(':')) return false;
            ws();
            if (key == "reads") {
                if (!expect('[')) return false;
                ws();
                if (p < end && *p == ']') {
                    ++p;
                } else {
                    while (true) {
                        if (!read(scan, key, tmp)) return false;
                        ws();
                        if (p < end && *p == ',') {
                            ++p;
                            continue;
                        }
                        if (!expect(']')) return false;
                        break;
                    }
                }
            } else if (key == "variants") {
                scan.variants_begin = p - begin;
                if (!skip_value()) return false;
                scan.variants_end = p - begin;
            } else if (!skip_value()) {
                return false;
            }
            ws();
            if (p < end && *p == ',') {
                ++p;
                continue;
            }
            return expect('}');
        }
    }
};

void table_sizes(const StringTable& t, int64_t* n, int64_t* bytes) {
    *n = (int64_t)t.items.size();
    int64_t b = 0;
    for (const auto& s : t.items) b += (int64_t)s.size();
    *bytes = b;
}

void table_fill(const StringTable& t, int64_t* off, char* bytes) {
    int64_t o = 0;
    int64_t i = 0;
    for (const auto& s : t.items) {
        off[i++] = o;
        memcpy(bytes + o, s.data(), s.size());
        o += (int64_t)s.size();
    }
    off[i] = o;
}

}  // namespace

// Scan a .variant.json held in memory.  Returns an opaque handle (nullptr on error, message in
// gk_last_error()); sizes[0] = reads, sizes[1..4] = total ids of lpv / lnv / rpv / rnv,
// sizes[5], sizes[6] = id strings and their bytes, sizes[7], sizes[8] = backbone strings and their
// bytes, sizes[9], sizes[10] = byte span of the "variants" value (-1 when absent).
extern "C" void* gk_json_scan(const char* buf, int64_t len, int64_t* sizes) {
    JsonScan* scan = new JsonScan();
    Parser ps{buf, buf + len, buf};
    if (!ps.document(*scan)) {
        gk_set_error("gk_json_scan: %s at byte %lld", ps.error ? ps.error : "parse error", (long long)(ps.p - buf));
        delete scan;
        return nullptr;
    }
    sizes[0] = (int64_t)scan->backbone.size();
    for (int w = 0; w < 4; ++w) sizes[1 + w] = (int64_t)scan->idx[w].size();
    table_sizes(scan->ids, &sizes[5], &sizes[6]);
    table_sizes(scan->genes, &sizes[7], &sizes[8]);
    sizes[9] = scan->variants_begin;
    sizes[10] = scan->variants_end;
    return scan;
}

// Copy the result of gk_json_scan into caller-allocated arrays: backbone / multiple [reads],
// off[w] [reads + 1] and idx[w] for the four lists, string tables as offsets [n + 1] + bytes.
extern "C" int gk_json_fill(void* handle, int32_t* backbone, int32_t* multiple, int64_t* const* off,
                            int32_t* const* idx, int64_t* id_off, char* id_bytes, int64_t* gene_off,
                            char* gene_bytes) {
    GK_REQUIRE(handle != nullptr, "gk_json_fill: null handle%s", "");
    const JsonScan* scan = static_cast<const JsonScan*>(handle);
    const size_t n = scan->backbone.size();
    if (n) {
        memcpy(backbone, scan->backbone.data(), n * sizeof(int32_t));
        memcpy(multiple, scan->multiple.data(), n * sizeof(int32_t));
    }
    for (int w = 0; w < 4; ++w) {
        memcpy(off[w], scan->off[w].data(), scan->off[w].size() * sizeof(int64_t));
        if (!scan->idx[w].empty()) memcpy(idx[w], scan->idx[w].data(), scan->idx[w].size() * sizeof(int32_t));
    }
    table_fill(scan->ids, id_off, id_bytes);
    table_fill(scan->genes, gene_off, gene_bytes);
    return 0;
}

extern "C" void gk_json_free(void* handle) { delete static_cast<JsonScan*>(handle); }

// Observation entries of the likelihood kernel from CSR lists (host; replaces two argsorts over all
// observations of a gene in packing.py).  For every read: its observations (list w has polarity
// polarity[w]: 1 = positive, 0 = negative) get an occurrence rank among identical (polarity,
// variant) pairs - duplicates must stay separate so that multiplicities are exact - and are merged
// per (rank, 32-variant word) into one entry (word, positive bits, negative bits); entries of a read
// are ordered by (rank, word).  ent_off has n_reads + 1 entries, the entry arrays room for one entry
// per observation; k_obs[r] = observations of read r.  Returns the number of entries, -1 on error.
extern "C" int64_t gk_pack_entries(int64_t n_reads, const int64_t* const* off, const int32_t* const* idx,
                                   const int32_t* polarity, int32_t* ent_off, int32_t* ent_word,
                                   uint32_t* ent_pos, uint32_t* ent_neg, int32_t* k_obs) {
    struct Obs {
        int32_t vid, pol, rank;
    };
    std::vector<Obs> obs;
    int64_t n_ent = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        obs.clear();
        for (int w = 0; w < 4; ++w)
            for (int64_t i = off[w][r]; i < off[w][r + 1]; ++i) obs.push_back({idx[w][i], polarity[w], 0});
        k_obs[r] = (int32_t)obs.size();
        ent_off[r] = (int32_t)n_ent;
        if (obs.empty()) continue;
        std::sort(obs.begin(), obs.end(), [](const Obs& a, const Obs& b) {
            return a.pol != b.pol ? a.pol < b.pol : a.vid < b.vid;
        });
        for (size_t i = 1; i < obs.size(); ++i)
            if (obs[i].pol == obs[i - 1].pol && obs[i].vid == obs[i - 1].vid) obs[i].rank = obs[i - 1].rank + 1;
        for (const Obs& o : obs)
            if (o.rank >= 256) {
                gk_set_error("gk_pack_entries: an observation is repeated more than 255 times in read pair %lld",
                             (long long)r);
                return -1;
            }
        std::sort(obs.begin(), obs.end(), [](const Obs& a, const Obs& b) {
            return a.rank != b.rank ? a.rank < b.rank : (a.vid >> 5) < (b.vid >> 5);
        });
        for (size_t i = 0; i < obs.size(); ++i) {
            const int32_t word = obs[i].vid >> 5;
            if (i == 0 || obs[i].rank != obs[i - 1].rank || word != (obs[i - 1].vid >> 5)) {
                ent_word[n_ent] = word;
                ent_pos[n_ent] = 0u;
                ent_neg[n_ent] = 0u;
                ++n_ent;
            }
            const uint32_t bit = 1u << (obs[i].vid & 31);
            if (obs[i].pol) ent_pos[n_ent - 1] |= bit;
            else ent_neg[n_ent - 1] |= bit;
        }
    }
    ent_off[n_reads] = (int32_t)n_ent;
    return n_ent;
}
