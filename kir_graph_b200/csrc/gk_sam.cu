// Host-side walk of one SAM record (reference: graphkir/hisat2.py:279-538,
// `recordToRawVariant` + `readZs` + `readMd`): CIGAR x MD x Zs -> match / single / insertion /
// deletion segments with 0-based backbone positions.  Pure host code; kir_graph_b200/hisat2.py keeps
// a Python statement of the same walk (`_RecordWalker`) that the tests compare this one with.
//
// State, as in the reference walk (hisat2.py:342-355):
//   pos     reference position of the current CIGAR operation
//   read_i  read offset of the current CIGAR operation
//   md_i    next MD token;   md_len  reference bases of the MD match run already read but not yet
//           consumed by CIGAR (it carries across an insertion, whose bases MD does not list)
//   zs_i    next Zs entry;   zs_pos  read offset up to which Zs gaps have been consumed
#include <cstdint>
#include <cstring>
#include <vector>

#include "gk_common.cuh"

namespace {

enum { kOk = 0, kSplicing = -3, kBadOp = -4, kAssert = -5, kIndex = -6, kValue = -7, kSpace = -8 };

struct Span {
    int32_t off, len;
};

struct MdTok {
    bool is_int;
    int64_t value;      // number, or the character
};

struct ZsEntry {
    int64_t gap;
    char kind;          // first character of the kind field; kind_len tells whether it is one character
    int32_t kind_len;
    Span id;
};

struct Walker {
    const char* line;
    std::vector<Span>& cols;          // scratch vectors of the calling thread (no allocation per record)
    std::vector<MdTok>& md;
    std::vector<ZsEntry>& zs;
    int64_t pos = 0, read_i = 0, md_len = 0, zs_pos = 0;
    size_t md_i = 0, zs_i = 0;
    Span seq{0, 0};
    int32_t* seg;
    int max_seg;
    int n_seg = 0;
    int status = kOk;

    Walker(std::vector<Span>& c, std::vector<MdTok>& m, std::vector<ZsEntry>& z) : cols(c), md(m), zs(z) {
        cols.clear();
        md.clear();
        zs.clear();
    }

    bool emit(int typ, int64_t p, int64_t length, Span val, Span id) {
        if (n_seg >= max_seg) {
            status = kSpace;
            return false;
        }
        int32_t* s = seg + 7 * n_seg++;
        s[0] = typ;
        s[1] = (int32_t)p;
        s[2] = (int32_t)length;
        s[3] = val.off;
        s[4] = val.len;
        s[5] = id.off;
        s[6] = id.len;
        return true;
    }
    // id of the Zs entry sitting exactly at the current read offset, else "unknown" (:357-371)
    Span known_id(char kind) {
        if (zs_i < zs.size()) {
            const ZsEntry& z = zs[zs_i];
            if (z.kind_len == 1 && z.kind == kind && read_i + md_len == zs_pos + z.gap) {
                zs_pos += z.gap + (kind == 'S' ? 1 : 0);
                ++zs_i;
                return z.id;
            }
        }
        return Span{0, -1};
    }
    void skip_zero() {
        if (md_i < md.size() && md[md_i].is_int && md[md_i].value == 0) ++md_i;
    }
    static bool is_base(const MdTok& t) {
        return !t.is_int && (t.value == 'A' || t.value == 'C' || t.value == 'G' || t.value == 'T');
    }
    // an M operation: match runs split by mismatches (:373-446)
    bool match(int64_t length) {
        int64_t done = 0;
        while (true) {
            if (md_len <= done && md_i < md.size() && md[md_i].is_int) {
                md_len += md[md_i].value;
                ++md_i;
            }
            if (md_len >= length) {
                md_len -= length;
                return emit(0, pos + done, length - done, Span{0, -2}, Span{0, -2});
            }
            const int64_t at = read_i + md_len;
            if (at < 0 || at >= seq.len) return fail(kIndex);
            const char base = line[seq.off + at];
            if (md_i >= md.size()) return fail(kIndex);
            if (md[md_i].is_int && md[md_i].value == 0) ++md_i;
            if (md_i >= md.size()) return fail(kIndex);
            if (!is_base(md[md_i]) || (char)md[md_i].value == base) return fail(kAssert);
            ++md_i;
            if (md_len > done && !emit(0, pos + done, md_len - done, Span{0, -2}, Span{0, -2})) return false;
            if (!emit(1, pos + md_len, 1, Span{(int32_t)(seq.off + at), 1}, known_id('S'))) return false;
            md_len += 1;
            done = md_len;
            if (md_len == length) {
                md_len = 0;
                return true;
            }
        }
    }
    bool fail(int code) {
        status = code;
        return false;
    }
};

bool parse_int(const char* s, int32_t n, int64_t& out) {      // int(): optional sign, digits, surrounding blanks
    int32_t i = 0;
    while (i < n && (s[i] == ' ')) ++i;
    bool neg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
    if (i >= n || s[i] < '0' || s[i] > '9') return false;
    int64_t v = 0;
    while (i < n && s[i] >= '0' && s[i] <= '9') v = v * 10 + (s[i++] - '0');
    while (i < n && s[i] == ' ') ++i;
    if (i != n) return false;
    out = neg ? -v : v;
    return true;
}

}  // namespace

// seg: int32 [max_seg][7] = typ (0 match, 1 single, 2 insertion, 3 deletion), pos, length,
//      val span (offset, length in `line`; length -2 = no value, -1 = the value is `length`),
//      id span (length -2 = no id, -1 = "unknown").
// meta: [0], [1] head / tail soft clip, [2], [3] span of the backbone name (column 3).
// Returns the number of segments, or < 0: -3 splicing (N), -4 unsupported CIGAR operation,
// -5 inconsistent record (the reference's asserts), -6 index out of range, -7 malformed number /
// Zs item, -8 more than max_seg segments.
extern "C" int gk_sam_walk(const char* line, int64_t len, int32_t* seg, int max_seg, int32_t* meta) {
    // line.strip()
    int64_t b = 0, e = len;
    auto blank = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; };
    while (b < e && blank(line[b])) ++b;
    while (e > b && blank(line[e - 1])) --e;
    static thread_local std::vector<Span> t_cols;
    static thread_local std::vector<MdTok> t_md;
    static thread_local std::vector<ZsEntry> t_zs;
    Walker w(t_cols, t_md, t_zs);
    w.line = line;
    w.seg = seg;
    w.max_seg = max_seg;
    for (int64_t i = b, start = b; i <= e; ++i) {
        if (i == e || line[i] == '\t') {
            w.cols.push_back(Span{(int32_t)start, (int32_t)(i - start)});
            start = i + 1;
        }
    }
    if (w.cols.size() < 4) return kIndex;            // same order of failures as the reference's field accesses
    meta[2] = w.cols[2].off;
    meta[3] = w.cols[2].len;
    int64_t pos1;
    if (!parse_int(line + w.cols[3].off, w.cols[3].len, pos1)) return kValue;
    if (w.cols.size() < 10) return kIndex;
    w.pos = pos1 - 1;
    w.seq = w.cols[9];
    // tags: the first column that starts with "Zs" / "MD" (readZs :518-527, readMd :530-538)
    bool has_zs = false, has_md = false;
    for (size_t c = 11; c < w.cols.size(); ++c) {
        const char* s = line + w.cols[c].off;
        const int32_t n = w.cols[c].len;
        if (!has_zs && n >= 2 && s[0] == 'Z' && s[1] == 's') {
            has_zs = true;
            int32_t i = n < 5 ? n : 5;
            while (true) {                                  // items separated by ',', fields by '|'
                int32_t j = i;
                while (j < n && s[j] != ',') ++j;
                // item.split("|") -> (int(f[0]), f[1], f[2]): the gap is parsed first (ValueError), a missing
                // second or third field is an IndexError, fields beyond the third are ignored
                int32_t bars[3] = {j, j, j}, nb = 0;
                for (int32_t k = i; k < j; ++k)
                    if (s[k] == '|') {
                        if (nb < 3) bars[nb] = k;
                        ++nb;
                    }
                ZsEntry z;
                if (!parse_int(s + i, bars[0] - i, z.gap)) return kValue;
                if (nb < 2) return kIndex;
                z.kind_len = bars[1] - bars[0] - 1;
                z.kind = z.kind_len > 0 ? s[bars[0] + 1] : '\0';
                z.id = Span{(int32_t)(w.cols[c].off + bars[1] + 1), (int32_t)(bars[2] - bars[1] - 1)};
                w.zs.push_back(z);
                if (j >= n) break;
                i = j + 1;
            }
        }
        if (!has_md && n >= 2 && s[0] == 'M' && s[1] == 'D') {
            has_md = true;
            for (int32_t i = n < 5 ? n : 5; i < n;) {
                if (s[i] >= '0' && s[i] <= '9') {
                    int64_t v = 0;
                    while (i < n && s[i] >= '0' && s[i] <= '9') v = v * 10 + (s[i++] - '0');
                    w.md.push_back(MdTok{true, v});
                } else {
                    w.md.push_back(MdTok{false, (unsigned char)s[i++]});
                }
            }
        }
    }
    meta[0] = meta[1] = 0;
    // CIGAR: every (digits)(word character) pair, like re.findall(r"(\d+)(\w)")
    const char* cg = line + w.cols[5].off;
    const int32_t cn = w.cols[5].len;
    int op_index = 0;
    for (int32_t i = 0; i < cn;) {
        if (cg[i] < '0' || cg[i] > '9') {
            ++i;
            continue;
        }
        int64_t length = 0;
        int32_t j = i;
        while (j < cn && cg[j] >= '0' && cg[j] <= '9') length = length * 10 + (cg[j++] - '0');
        auto is_word = [](char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_'; };
        if (j >= cn || !is_word(cg[j])) {
            // (\d+)(\w) backtracks: the last digit of a longer run is taken as the operation, which no
            // branch of the walk knows; a single digit matches nothing
            if (j - i >= 2) return kBadOp;
            i = j;
            continue;
        }
        const char op = cg[j];
        i = j + 1;
        w.skip_zero();
        if (op == 'M') {
            if (!w.match(length)) return w.status;
        } else if (op == 'I') {
            int64_t lo = w.read_i < w.seq.len ? w.read_i : w.seq.len;
            int64_t hi = w.read_i + length < w.seq.len ? w.read_i + length : w.seq.len;
            if (lo < 0) lo = 0;
            if (hi < lo) hi = lo;
            if (!w.emit(2, w.pos, length, Span{(int32_t)(w.seq.off + lo), (int32_t)(hi - lo)}, w.known_id('I')))
                return w.status;
        } else if (op == 'D') {
            if (w.md_i >= w.md.size()) return kIndex;
            if (w.md[w.md_i].is_int || w.md[w.md_i].value != '^') return kAssert;
            ++w.md_i;
            while (w.md_i < w.md.size() && Walker::is_base(w.md[w.md_i])) ++w.md_i;
            if (!w.emit(3, w.pos, length, Span{0, -1}, w.known_id('D'))) return w.status;
        } else if (op == 'S') {
            meta[op_index == 0 ? 0 : 1] = (int32_t)length;
            w.zs_pos += length;
        } else if (op == 'N') {
            return kSplicing;
        } else {
            return kBadOp;
        }
        if (op == 'M' || op == 'N' || op == 'D') w.pos += length;
        if (op == 'M' || op == 'I' || op == 'S') w.read_i += length;
        ++op_index;
    }
    w.skip_zero();
    if (w.zs_i != w.zs.size() || w.md_i != w.md.size() || w.read_i != w.seq.len) return kAssert;
    return w.n_seg;
}

// ---------------------------------------------------------------------------------------
// Batch extraction: name-sorted SAM text -> per read pair the positive / negative variant lists
// as CSR arrays, without a Python object per record (reference: graphkir/hisat2.py, readPair
// :228-276, filterRead :541-578, findVariantId :581-606, recordToVariants :657-689,
// getVariantsBoundary :692-713, getPNFromVariantList :716-800, extractVariant :803-844; the
// pileup-based error correction is off on the CLI path, main.py:149, and not covered).
// kir_graph_b200/hisat2.py keeps the Python statement of every step; tests compare the two.
#include <string>
#include <unordered_map>
#include <algorithm>

namespace {

struct XVar {                       // a table or novel variant
    int32_t ref, pos, typ;          // typ: 0 insertion, 1 single, 2 deletion (the order of Variant.__lt__), 3 match
    int32_t val_int;                // deletion length
    std::string val;                // base / inserted sequence
    int32_t length;
};

struct XKey {
    int32_t ref, pos, typ, val_int;
    std::string val;
    bool operator==(const XKey& o) const {
        return ref == o.ref && pos == o.pos && typ == o.typ && val_int == o.val_int && val == o.val;
    }
};

struct XKeyHash {
    size_t operator()(const XKey& k) const {
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
        mix((uint32_t)k.ref);
        mix((uint32_t)k.pos);
        mix((uint32_t)k.typ);
        mix((uint32_t)k.val_int);
        for (char c : k.val) mix((unsigned char)c);
        return (size_t)h;
    }
};

struct ReadVar {                    // one segment of a record after findVariantId
    int32_t pos, typ, val_int, length;
    std::string val;
    int32_t index;                  // variant index (table or n_table + novel), -1 for a match segment
    bool novel;
};

struct Extract {
    std::vector<XVar> table;        // sorted index variants, then the novel ones in creation order
    int32_t n_table = 0;
    std::unordered_map<XKey, int32_t, XKeyHash> by_key;
    std::vector<std::string> refs;
    std::unordered_map<std::string, int32_t> ref_index;
    std::vector<int32_t> multiple, backbone;
    std::vector<int64_t> span;      // per pair: offset, length of the left record and of the right record
    std::string json;               // gk_sam_extract_json: the "reads" array, owned by the handle
    std::vector<int64_t> off[4];    // lpv, lnv, rpv, rnv
    std::vector<int32_t> idx[4];
    int64_t n_strange = 0;
    int status = 0;
    int64_t bad_line = -1;

    int32_t ref_id(const std::string& name) {
        auto it = ref_index.find(name);
        if (it != ref_index.end()) return it->second;
        const int32_t id = (int32_t)refs.size();
        refs.push_back(name);
        ref_index.emplace(name, id);
        return id;
    }
};

// Variant.__lt__ on (ref, pos, typ rank, val) for variants of one reference
bool var_less(int32_t pos_a, int32_t typ_a, int32_t vi_a, const std::string& vs_a, int32_t pos_b, int32_t typ_b,
              int32_t vi_b, const std::string& vs_b) {
    if (pos_a != pos_b) return pos_a < pos_b;
    if (typ_a != typ_b) return typ_a < typ_b;
    if (typ_a == 2) return vi_a < vi_b;
    return vs_a < vs_b;
}

struct Line {
    const char* s;
    int64_t n;
};

void split_tabs(const char* s, int64_t n, std::vector<Span>& cols, size_t limit) {
    cols.clear();
    int64_t start = 0;
    for (int64_t i = 0; i <= n && cols.size() < limit; ++i)
        if (i == n || s[i] == '\t') {
            cols.push_back(Span{(int32_t)start, (int32_t)(i - start)});
            start = i + 1;
        }
}

// filterRead (:541-578): flag & 2 and an NM tag <= num_editdist; -1 on a malformed number
int filter_read(const Line& ln, int num_editdist, std::vector<Span>& cols) {
    int64_t b = 0, e = ln.n;
    auto blank = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; };
    while (b < e && blank(ln.s[b])) ++b;
    while (e > b && blank(ln.s[e - 1])) --e;
    split_tabs(ln.s + b, e - b, cols, (size_t)-1);
    if (cols.size() < 2) return -2;
    int64_t flag;
    if (!parse_int(ln.s + b + cols[1].off, cols[1].len, flag)) return -1;
    if ((flag & 2) == 0) return 0;
    bool has_nm = false;
    int64_t nm = 0;
    for (size_t c = 11; c < cols.size(); ++c) {
        const char* s = ln.s + b + cols[c].off;
        if (cols[c].len >= 2 && s[0] == 'N' && s[1] == 'M') {
            const int32_t skip = cols[c].len < 5 ? cols[c].len : 5;
            if (!parse_int(s + skip, cols[c].len - skip, nm)) return -1;
            has_nm = true;
        }
    }
    return has_nm && nm <= num_editdist ? 1 : 0;
}

// NH tag of a record, 1 when absent: re.search(r"NH:i:(\d+)")
int32_t get_nh(const Line& ln) {
    for (int64_t i = 0; i + 5 < ln.n; ++i)
        if (ln.s[i] == 'N' && ln.s[i + 1] == 'H' && ln.s[i + 2] == ':' && ln.s[i + 3] == 'i' && ln.s[i + 4] == ':' &&
            ln.s[i + 5] >= '0' && ln.s[i + 5] <= '9') {
            int64_t v = 0;
            for (int64_t j = i + 5; j < ln.n && ln.s[j] >= '0' && ln.s[j] <= '9'; ++j) v = v * 10 + (ln.s[j] - '0');
            return (int32_t)v;
        }
    return 1;
}

// recordToVariants (:657-689): walk, findVariantId in walk order, sort; empty for a soft-clipped record
bool record_variants(Extract& ex, const Line& ln, std::vector<ReadVar>& out, int32_t& novel_id,
                     std::vector<int32_t>& seg) {
    out.clear();
    int32_t meta[4];
    int n = gk_sam_walk(ln.s, ln.n, seg.data(), (int)(seg.size() / 7), meta);
    if (n == kSpace) {
        seg.resize((size_t)(ln.n + 1) * 7);
        n = gk_sam_walk(ln.s, ln.n, seg.data(), (int)(seg.size() / 7), meta);
    }
    if (n < 0) {
        ex.status = n;
        return false;
    }
    if (meta[0] + meta[1] > 0) return true;
    const int32_t ref = ex.ref_id(std::string(ln.s + meta[2], (size_t)meta[3]));
    for (int i = 0; i < n; ++i) {
        const int32_t* s = seg.data() + 7 * i;
        ReadVar rv;
        rv.pos = s[1];
        rv.length = s[2];
        rv.val_int = 0;
        rv.index = -1;
        rv.novel = false;
        switch (s[0]) {                                  // walker types -> order of Variant.__lt__
            case 0: rv.typ = 3; break;
            case 1: rv.typ = 1; rv.val.assign(ln.s + s[3], (size_t)s[4]); break;
            case 2: rv.typ = 0; rv.val.assign(ln.s + s[3], (size_t)s[4]); break;
            default: rv.typ = 2; rv.val_int = s[2]; break;
        }
        if (rv.typ != 3) {                               // findVariantId (:581-606)
            XKey key{ref, rv.pos, rv.typ, rv.val_int, rv.val};
            auto it = ex.by_key.find(key);
            if (it != ex.by_key.end()) {
                rv.index = it->second;
                rv.novel = rv.index >= ex.n_table;
                rv.length = ex.table[rv.index].length;   // the map's object replaces the segment (:594-595)
            } else {
                rv.index = (int32_t)ex.table.size();
                rv.novel = true;
                ex.table.push_back(XVar{ref, rv.pos, rv.typ, rv.val_int, rv.val, rv.length});
                ex.by_key.emplace(std::move(key), rv.index);
                ++novel_id;
            }
        }
        out.push_back(std::move(rv));
    }
    std::stable_sort(out.begin(), out.end(), [](const ReadVar& a, const ReadVar& b) {
        return var_less(a.pos, a.typ, a.val_int, a.val, b.pos, b.typ, b.val_int, b.val);
    });
    return true;
}

// getPNFromVariantList (:716-800) with exon_only = False, discard_novel_index = True
void positives_negatives(Extract& ex, int32_t ref, const std::vector<ReadVar>& rv, int which_pos, int which_neg,
                         const std::vector<int32_t>& ref_begin) {
    if (rv.empty()) return;
    for (const ReadVar& v : rv)
        if ((v.typ == 0 || v.typ == 2) && v.novel) return;          // a novel indel is taken as a mapping error
    // window of the sorted table: bisect_left with the probes (first.pos, single, "A") and
    // (last.pos + last.length, single, "T"); the table is sorted by (ref, pos, typ, val), so the
    // search runs over the slice of this reference
    const int32_t lo_i = ref < (int32_t)ref_begin.size() - 1 ? ref_begin[ref] : ex.n_table;
    const int32_t hi_i = ref < (int32_t)ref_begin.size() - 1 ? ref_begin[ref + 1] : ex.n_table;
    auto lower = [&](int32_t pos, const std::string& base) {
        int32_t a = lo_i, b = hi_i;
        while (a < b) {
            const int32_t m = a + (b - a) / 2;
            const XVar& t = ex.table[m];
            if (var_less(t.pos, t.typ, t.val_int, t.val, pos, 1, 0, base)) a = m + 1;
            else b = m;
        }
        return a;
    };
    const int32_t read_end = rv.back().pos + rv.back().length;
    const int32_t left = lower(rv.front().pos, "A"), right = lower(read_end, "T");
    std::vector<XKey> excluded;
    for (const ReadVar& v : rv) {
        if (v.typ != 2 && v.val == "N")                              // a masked base matches every base
            for (const char* base : {"A", "T", "C", "G"}) excluded.push_back(XKey{ref, v.pos, v.typ, v.val_int, base});
        if (v.typ != 3) {
            excluded.push_back(XKey{ref, v.pos, v.typ, v.val_int, v.val});
            ex.idx[which_pos].push_back(v.index);
        }
    }
    for (int32_t i = left; i < right; ++i) {
        const XVar& t = ex.table[i];
        bool skip = false;
        for (const XKey& k : excluded)
            if (k.pos == t.pos && k.typ == t.typ && k.val_int == t.val_int && k.val == t.val) {
                skip = true;
                break;
            }
        if (skip) continue;
        if (t.typ == 2 && t.pos + t.val_int + 10 >= read_end) continue;   // ambiguous near the read end
        ex.idx[which_neg].push_back(i);
    }
}

}  // namespace

// See include/gk_typing.h.
extern "C" void* gk_sam_extract(const char* sam, int64_t sam_len, int32_t n_var, const int32_t* v_ref,
                                const int32_t* v_pos, const int32_t* v_typ, const int32_t* v_val_int,
                                const int32_t* v_length, const int64_t* v_val_off, const char* v_val_bytes,
                                int32_t n_ref, const int64_t* ref_off, const char* ref_bytes, int32_t novel_id,
                                int32_t num_editdist, int64_t* sizes) {
    Extract* ex = new Extract();
    for (int32_t i = 0; i < n_ref; ++i) ex->ref_id(std::string(ref_bytes + ref_off[i], (size_t)(ref_off[i + 1] - ref_off[i])));
    ex->n_table = n_var;
    std::vector<int32_t> ref_begin(ex->refs.size() + 1, n_var);      // table slice per reference (refs ascending)
    for (int32_t i = 0; i < n_var; ++i) {
        XVar v{v_ref[i], v_pos[i], v_typ[i], v_val_int[i],
               std::string(v_val_bytes + v_val_off[i], (size_t)(v_val_off[i + 1] - v_val_off[i])), v_length[i]};
        ex->by_key[XKey{v.ref, v.pos, v.typ, v.val_int, v.val}] = i;   // later duplicates win, as dict() does
        ex->table.push_back(std::move(v));
    }
    for (int32_t i = n_var - 1; i >= 0; --i) ref_begin[v_ref[i]] = i;
    for (int32_t r = (int32_t)ex->refs.size() - 1; r >= 0; --r)
        if (ref_begin[r] == n_var && r + 1 <= (int32_t)ex->refs.size()) ref_begin[r] = ref_begin[r + 1];
    for (int w = 0; w < 4; ++w) ex->off[w].push_back(0);

    struct Pending {
        Line line;
        int64_t flag;
    };
    std::unordered_map<std::string, Pending> pending;               // key: name \t ref \t pos \t (flag & 256)
    std::vector<Span> cols, fcols;
    std::vector<ReadVar> lv, rv;
    std::vector<int32_t> seg(7 * 64);
    std::string key, mate_key;
    int64_t p = 0, line_no = 0;
    while (p < sam_len && ex->status == 0) {
        const char* nl = static_cast<const char*>(memchr(sam + p, '\n', (size_t)(sam_len - p)));
        const int64_t e = nl ? nl - sam : sam_len;
        Line ln{sam + p, e - p};
        p = e + 1;
        ++line_no;
        if (ln.n == 0 || ln.s[0] == '@' || (ln.n >= 15 && memcmp(ln.s, "[bam_sort_core]", 15) == 0)) continue;
        split_tabs(ln.s, ln.n, cols, 8);
        if (cols.size() < 8) {
            ex->status = kValue;
            ex->bad_line = line_no;
            break;
        }
        if (!(cols[6].len == 1 && ln.s[cols[6].off] == '=')) continue;
        int64_t flag;
        if (!parse_int(ln.s + cols[1].off, cols[1].len, flag)) {
            ex->status = kValue;
            ex->bad_line = line_no;
            break;
        }
        auto make_key = [&](std::string& out, const Span& pos) {
            out.assign(ln.s + cols[0].off, (size_t)cols[0].len);
            out.push_back('\t');
            out.append(ln.s + cols[2].off, (size_t)cols[2].len);
            out.push_back('\t');
            out.append(ln.s + pos.off, (size_t)pos.len);
            out.push_back('\t');
            out.push_back((flag & 256) ? '1' : '0');
        };
        make_key(mate_key, cols[7]);
        auto it = pending.find(mate_key);
        if (it == pending.end()) {
            make_key(key, cols[3]);
            pending[key] = Pending{ln, flag};
            continue;
        }
        if (((it->second.flag | flag) & 192) != 192) {
            ++ex->n_strange;
            continue;
        }
        const Line left = ln, right = it->second.line;              // (current record, earlier mate), as readPair yields
        pending.erase(it);
        const int fl = filter_read(left, num_editdist, fcols);
        const int fr = fl == 1 ? filter_read(right, num_editdist, fcols) : 0;
        if (fl < 0 || fr < 0) {
            ex->status = fl == -2 || fr == -2 ? kIndex : kValue;
            ex->bad_line = line_no;
            break;
        }
        if (fl != 1 || fr != 1) continue;
        if (!record_variants(*ex, left, lv, novel_id, seg) || !record_variants(*ex, right, rv, novel_id, seg)) {
            ex->bad_line = line_no;
            break;
        }
        const int32_t ref = ex->ref_id(std::string(ln.s + cols[2].off, (size_t)cols[2].len));
        // the reference table may not know this name: ref_begin covers the names given by the caller
        positives_negatives(*ex, ref, lv, 0, 1, ref_begin);
        positives_negatives(*ex, ref, rv, 2, 3, ref_begin);
        ex->multiple.push_back(get_nh(left));
        ex->backbone.push_back(ref);
        for (const Line& l : {left, right}) {
            ex->span.push_back((int64_t)(l.s - sam));
            ex->span.push_back(l.n);
        }
        for (int w = 0; w < 4; ++w) ex->off[w].push_back((int64_t)ex->idx[w].size());
    }
    sizes[0] = (int64_t)ex->multiple.size();
    for (int w = 0; w < 4; ++w) sizes[1 + w] = (int64_t)ex->idx[w].size();
    sizes[5] = (int64_t)ex->table.size() - n_var;
    int64_t vb = 0;
    for (size_t i = (size_t)n_var; i < ex->table.size(); ++i) vb += (int64_t)ex->table[i].val.size();
    sizes[6] = vb;
    sizes[7] = (int64_t)ex->refs.size();
    int64_t rb = 0;
    for (const auto& r : ex->refs) rb += (int64_t)r.size();
    sizes[8] = rb;
    sizes[9] = ex->status;
    sizes[10] = ex->n_strange;
    sizes[11] = ex->bad_line;
    return ex;
}

extern "C" int gk_sam_extract_fill(void* handle, int32_t* multiple, int32_t* backbone, int64_t* span, int64_t* const* off,
                                   int32_t* const* idx, int32_t* nv_ref, int32_t* nv_pos, int32_t* nv_typ,
                                   int32_t* nv_val_int, int32_t* nv_length, int64_t* nv_val_off, char* nv_val_bytes,
                                   int64_t* ref_off, char* ref_bytes) {
    GK_REQUIRE(handle != nullptr, "gk_sam_extract_fill: null handle%s", "");
    const Extract* ex = static_cast<const Extract*>(handle);
    const size_t n = ex->multiple.size();
    if (n) {
        memcpy(multiple, ex->multiple.data(), n * sizeof(int32_t));
        memcpy(backbone, ex->backbone.data(), n * sizeof(int32_t));
        memcpy(span, ex->span.data(), 4 * n * sizeof(int64_t));
    }
    for (int w = 0; w < 4; ++w) {
        memcpy(off[w], ex->off[w].data(), ex->off[w].size() * sizeof(int64_t));
        if (!ex->idx[w].empty()) memcpy(idx[w], ex->idx[w].data(), ex->idx[w].size() * sizeof(int32_t));
    }
    int64_t o = 0;
    size_t j = 0;
    for (size_t i = (size_t)ex->n_table; i < ex->table.size(); ++i, ++j) {
        const XVar& v = ex->table[i];
        nv_ref[j] = v.ref;
        nv_pos[j] = v.pos;
        nv_typ[j] = v.typ;
        nv_val_int[j] = v.val_int;
        nv_length[j] = v.length;
        nv_val_off[j] = o;
        memcpy(nv_val_bytes + o, v.val.data(), v.val.size());
        o += (int64_t)v.val.size();
    }
    nv_val_off[j] = o;
    o = 0;
    j = 0;
    for (const auto& r : ex->refs) {
        ref_off[j++] = o;
        memcpy(ref_bytes + o, r.data(), r.size());
        o += (int64_t)r.size();
    }
    ref_off[j] = o;
    return 0;
}

namespace {

// json.dumps(str) with ensure_ascii=True: the escapes of Python's encoder (py_encode_basestring_ascii)
bool json_string(std::string& out, const char* s, int64_t n) {
    static const char hex[] = "0123456789abcdef";
    auto u16 = [&](uint32_t c) {
        out += "\\u";
        out.push_back(hex[(c >> 12) & 15]);
        out.push_back(hex[(c >> 8) & 15]);
        out.push_back(hex[(c >> 4) & 15]);
        out.push_back(hex[c & 15]);
    };
    out.push_back('"');
    for (int64_t i = 0; i < n;) {
        int64_t j = i;                                      // run of characters the encoder copies as they are
        while (j < n && s[j] >= ' ' && s[j] <= '~' && s[j] != '"' && s[j] != '\\') ++j;
        if (j > i) {
            out.append(s + i, (size_t)(j - i));
            i = j;
            continue;
        }
        const unsigned char c = (unsigned char)s[i];
        if (c < 0x80) {
            switch (c) {
                case '"': out += "\\\""; break;
                case '\\': out += "\\\\"; break;
                case '\n': out += "\\n"; break;
                case '\r': out += "\\r"; break;
                case '\t': out += "\\t"; break;
                case '\b': out += "\\b"; break;
                case '\f': out += "\\f"; break;
                default:
                    if (c < 0x20 || c == 0x7f) u16(c);              // the encoder keeps ' ' .. '~' only
                    else out.push_back((char)c);
            }
            ++i;
            continue;
        }
        // UTF-8 -> code point (what bytes.decode("utf-8") accepts), then \uXXXX / a surrogate pair
        int len = c >= 0xf0 ? 4 : c >= 0xe0 ? 3 : c >= 0xc2 ? 2 : 0;
        if (len == 0 || c > 0xf4 || i + len > n) return false;
        uint32_t cp = len == 2 ? c & 0x1f : len == 3 ? c & 0x0f : c & 0x07;
        for (int k = 1; k < len; ++k) {
            const unsigned char t = (unsigned char)s[i + k];
            if ((t & 0xc0) != 0x80) return false;
            cp = (cp << 6) | (t & 0x3f);
        }
        if ((len == 3 && (cp < 0x800 || (cp >= 0xd800 && cp <= 0xdfff))) || (len == 4 && (cp < 0x10000 || cp > 0x10ffff)))
            return false;
        if (cp >= 0x10000) {
            cp -= 0x10000;
            u16(0xd800 + (cp >> 10));
            u16(0xdc00 + (cp & 0x3ff));
        } else {
            u16(cp);
        }
        i += len;
    }
    out.push_back('"');
    return true;
}

}  // namespace

// See include/gk_typing.h.
extern "C" int gk_sam_extract_json(void* handle, const char* sam, const int64_t* id_off, const char* id_bytes,
                                   int32_t novel_id, const char** out, int64_t* out_len) {
    GK_REQUIRE(handle != nullptr && out != nullptr && out_len != nullptr, "gk_sam_extract_json: null argument%s", "");
    Extract* ex = static_cast<Extract*>(handle);
    std::string& js = ex->json;
    js.clear();
    js.reserve(ex->span.size() / 4 * 1024);
    static const char* const names[4] = {"\"lpv\": [", "\"lnv\": [", "\"rpv\": [", "\"rnv\": ["};
    char num[32];
    const size_t n = ex->multiple.size();
    for (size_t r = 0; r < n; ++r) {
        js += r ? ", {\"l_sam\": " : "{\"l_sam\": ";
        if (!json_string(js, sam + ex->span[4 * r], ex->span[4 * r + 1])) {
            gk_set_error("gk_sam_extract_json: record of pair %lld is not UTF-8", (long long)r);
            return -1;
        }
        js += ", \"r_sam\": ";
        if (!json_string(js, sam + ex->span[4 * r + 2], ex->span[4 * r + 3])) {
            gk_set_error("gk_sam_extract_json: record of pair %lld is not UTF-8", (long long)r);
            return -1;
        }
        js += ", \"multiple\": ";
        js.append(num, (size_t)snprintf(num, sizeof num, "%d", ex->multiple[r]));
        js += ", \"backbone\": ";
        const std::string& ref = ex->refs[(size_t)ex->backbone[r]];
        if (!json_string(js, ref.data(), (int64_t)ref.size())) {
            gk_set_error("gk_sam_extract_json: backbone name of pair %lld is not UTF-8", (long long)r);
            return -1;
        }
        for (int w = 0; w < 4; ++w) {
            js += ", ";
            js += names[w];
            for (int64_t j = ex->off[w][r]; j < ex->off[w][r + 1]; ++j) {
                if (j > ex->off[w][r]) js += ", ";
                const int32_t v = ex->idx[w][(size_t)j];
                if (v < ex->n_table) {
                    if (!json_string(js, id_bytes + id_off[v], id_off[v + 1] - id_off[v])) {
                        gk_set_error("gk_sam_extract_json: id of variant %d is not UTF-8", v);
                        return -1;
                    }
                } else {
                    js.append(num, (size_t)snprintf(num, sizeof num, "\"nv%d\"", novel_id + (v - ex->n_table)));
                }
            }
            js.push_back(']');
        }
        js.push_back('}');
    }
    *out = js.data();
    *out_len = (int64_t)js.size();
    return 0;
}

extern "C" void gk_sam_extract_free(void* handle) { delete static_cast<Extract*>(handle); }
