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
    std::vector<Span> cols;
    std::vector<MdTok> md;
    std::vector<ZsEntry> zs;
    int64_t pos = 0, read_i = 0, md_len = 0, zs_pos = 0;
    size_t md_i = 0, zs_i = 0;
    Span seq{0, 0};
    int32_t* seg;
    int max_seg;
    int n_seg = 0;
    int status = kOk;

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
    Walker w;
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
                int32_t bars[2], nb = 0;
                for (int32_t k = i; k < j; ++k)
                    if (s[k] == '|') {
                        if (nb < 2) bars[nb] = k;
                        ++nb;
                    }
                if (nb != 2) return kValue;
                ZsEntry z;
                if (!parse_int(s + i, bars[0] - i, z.gap)) return kValue;
                z.kind_len = bars[1] - bars[0] - 1;
                z.kind = z.kind_len > 0 ? s[bars[0] + 1] : '\0';
                z.id = Span{(int32_t)(w.cols[c].off + bars[1] + 1), (int32_t)(j - bars[1] - 1)};
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
