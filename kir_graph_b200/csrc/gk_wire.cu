// Wire format of the read observations (host -> device), and its expansion on the device.
//
// What crosses PCIe per read pair decides the end-to-end throughput of a cohort on 8 GPUs (round 1:
// 36 B per pair = 694 MB per 96-sample pass; two GPUs share a PCIe switch uplink, so a GPU gets
// ~27 GB/s and the copies, not the kernels, set the pace).  The observation entries the likelihood
// kernel reads - (word, positive bits, negative bits), graphkir/typing_mulit_allele.py:340-381 in
// packed form - are therefore not shipped; the host sends what graphkir/hisat2.py's
// getPNFromVariantList (:716-800) actually decided per mate:
//     negatives = every variant of the window [lo, lo + n) of the (position-sorted) variant table
//                 that is not a positive of the mate and not excluded,
// i.e. per mate {lo, n, bitmap of the positives inside the window, the few excluded variants
// ("holes"), positives outside the window (novel variants)}, and the device rebuilds the entries.
// Variants that errorCorrection (:302-338) removed from the negative lists of the whole gene are not
// holes: a per-gene mask `neg_keep` (variants that occur as a negative anywhere in the gene) is
// applied to the window on the device.  14 B per read pair on the cfg3 / cfg5 workloads.
//
// Per read   hdr (uint16): bits 0-7 number of entries the expansion emits, bits 8-15 record length
//            in 2-byte units, 0 = raw record.
// Record     mate L then mate R, each (uint16 units):
//              [0] lo   [1] n | n_out << 8 | n_hole << 12
//              ceil(n / 16) units: bit i = variant lo + i is a positive of the mate
//              ceil(n_hole / 2) units: two 8-bit window offsets each
//              n_out units: variant index of a positive outside the window
// Raw record (anything the above cannot express: a variant twice in a list, positive and negative in
//            one mate, windows over 255 variants, more than 15 holes or outside positives, variant
//            indices beyond 65535): the read's canonical entries, 5 units each: word, pos lo, pos hi,
//            neg lo, neg hi.
// Expansion  the window words of the two mates merged in ascending order (words with no bit are
//            skipped); bits that both mates observe go to a second entry of the same word, which is how
//            the likelihood counts a variant seen by both mates twice (:363-368); then one entry per
//            outside positive.
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "gk_common.cuh"

namespace {

struct MateRec {
    int lo = 0, n = 0;
    std::vector<int> pos_in, holes, pos_out;
    bool ok = true;
};

inline bool keep_bit(const uint32_t* keep, int v) { return (keep[v >> 5] >> (v & 31)) & 1u; }

// Build the record of one mate from its positive / negative variant lists.
MateRec mate_record(const int32_t* pos, int64_t n_pos, const int32_t* neg, int64_t n_neg, const uint32_t* neg_keep,
                    std::vector<int>& scratch) {
    MateRec m;
    scratch.clear();
    if (n_neg > 0) {
        scratch.assign(neg, neg + n_neg);
        std::sort(scratch.begin(), scratch.end());
        for (size_t i = 1; i < scratch.size(); ++i)
            if (scratch[i] == scratch[i - 1]) m.ok = false;                  // a negative twice
        m.lo = scratch.front();
        m.n = scratch.back() - scratch.front() + 1;
        if (m.n > 255 || scratch.back() > 65535) m.ok = false;
    }
    std::vector<int> p(pos, pos + n_pos);
    std::sort(p.begin(), p.end());
    for (size_t i = 1; i < p.size(); ++i)
        if (p[i] == p[i - 1]) m.ok = false;                                  // a positive twice
    if (!m.ok) return m;
    // the window is the hull of everything the mate observed; positives that would stretch it beyond
    // 255 variants (novel variants sit at the end of the table) stay outside and are listed one by one
    if (!p.empty()) {
        const int lo_all = n_neg > 0 ? std::min(m.lo, p.front()) : p.front();
        const int hi_all = n_neg > 0 ? std::max(m.lo + m.n - 1, p.back()) : p.back();
        if (hi_all - lo_all + 1 <= 255 && hi_all <= 65535) {
            m.lo = lo_all;
            m.n = hi_all - lo_all + 1;
        }
    }
    for (int v : p) {
        if (v > 65535) { m.ok = false; return m; }
        if (v >= m.lo && v < m.lo + m.n) {
            if (n_neg > 0 && std::binary_search(scratch.begin(), scratch.end(), v)) { m.ok = false; return m; }
            m.pos_in.push_back(v);
        } else {
            m.pos_out.push_back(v);
        }
    }
    // window variants that are neither positive nor negative for this mate although the gene keeps
    // them as negatives elsewhere: excluded by the extraction (N-masked base, deletion near the read
    // end: hisat2.py:785-797) - and any negative the gene mask would not rebuild
    size_t ni = 0, pi = 0;
    for (int v = m.lo; v < m.lo + m.n; ++v) {
        while (ni < scratch.size() && scratch[ni] < v) ++ni;
        while (pi < m.pos_in.size() && m.pos_in[pi] < v) ++pi;
        const bool is_neg = ni < scratch.size() && scratch[ni] == v;
        const bool is_pos = pi < m.pos_in.size() && m.pos_in[pi] == v;
        const bool kept = keep_bit(neg_keep, v);
        if (is_neg && !kept) { m.ok = false; return m; }                    // inconsistent mask: raw
        if (!is_neg && !is_pos && kept) m.holes.push_back(v - m.lo);
    }
    if (m.holes.size() > 15 || m.pos_out.size() > 15) m.ok = false;
    return m;
}

inline int mate_units(const MateRec& m) { return 2 + (m.n + 15) / 16 + ((int)m.holes.size() + 1) / 2 + (int)m.pos_out.size(); }

struct WordBits {
    int w;
    uint32_t p, n;
};

// (word, positive bits, negative bits) of a mate's window, every word of the window in ascending order
std::vector<WordBits> mate_words(const MateRec& m, const uint32_t* neg_keep) {
    std::vector<WordBits> out;
    if (m.n == 0) return out;
    for (int w = m.lo >> 5; w <= (m.lo + m.n - 1) >> 5; ++w) {
        uint32_t bits = 0u;
        const int v0 = std::max(m.lo, 32 * w), v1 = std::min(m.lo + m.n, 32 * w + 32);
        for (int v = v0; v < v1; ++v) bits |= 1u << (v & 31);
        uint32_t posb = 0u, holeb = 0u;
        for (int v : m.pos_in) if ((v >> 5) == w) posb |= 1u << (v & 31);
        for (int h : m.holes) if (((m.lo + h) >> 5) == w) holeb |= 1u << ((m.lo + h) & 31);
        out.push_back({w, posb, bits & neg_keep[w] & ~posb & ~holeb});
    }
    return out;
}

// Entries the expansion emits for a read pair (the rule of expand_pair below): the window words of the
// two mates are merged word by word - bits both mates observe go to a second entry, so that they count
// twice - and every outside positive is an entry of its own.
int pair_entries(const MateRec& l, const MateRec& r, const uint32_t* neg_keep) {
    const std::vector<WordBits> a = mate_words(l, neg_keep), b = mate_words(r, neg_keep);
    size_t i = 0, j = 0;
    int count = 0;
    while (i < a.size() || j < b.size()) {
        const int w = std::min(i < a.size() ? a[i].w : INT32_MAX, j < b.size() ? b[j].w : INT32_MAX);
        uint32_t lp = 0u, ln = 0u, rp = 0u, rn = 0u;
        if (i < a.size() && a[i].w == w) { lp = a[i].p; ln = a[i].n; ++i; }
        if (j < b.size() && b[j].w == w) { rp = b[j].p; rn = b[j].n; ++j; }
        const uint32_t ov = (lp | ln) & (rp | rn);
        count += ((lp | ln | ((rp | rn) & ~ov)) != 0u) + (ov != 0u);
    }
    return count + (int)l.pos_out.size() + (int)r.pos_out.size();
}

void write_mate(const MateRec& m, uint16_t* out) {
    out[0] = (uint16_t)m.lo;
    out[1] = (uint16_t)(m.n | ((int)m.pos_out.size() << 8) | ((int)m.holes.size() << 12));
    int p = 2;
    const int nb = (m.n + 15) / 16;
    for (int i = 0; i < nb; ++i) out[p + i] = 0;
    for (int v : m.pos_in) out[p + ((v - m.lo) >> 4)] |= (uint16_t)(1u << ((v - m.lo) & 15));
    p += nb;
    for (size_t i = 0; i < m.holes.size(); i += 2)
        out[p++] = (uint16_t)(m.holes[i] | ((i + 1 < m.holes.size() ? m.holes[i + 1] : 0) << 8));
    for (int v : m.pos_out) out[p++] = (uint16_t)v;
}

}  // namespace

// off[w] / idx[w]: the CSR lists in the order lpv, rpv, lnv, rnv (kir_graph_b200.synthetic.LIST_NAMES).
extern "C" int64_t gk_wire_encode(int64_t n_reads, const int64_t* const* off, const int32_t* const* idx,
                                  const uint32_t* neg_keep, const int32_t* ent_off, const int32_t* ent_word,
                                  const uint32_t* ent_pos, const uint32_t* ent_neg, uint16_t* hdr, uint16_t* stream,
                                  int64_t capacity, int64_t* n_entries_out) {
    std::vector<int> scratch;
    int64_t units = 0, entries = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        MateRec mate[2];
        for (int s = 0; s < 2; ++s)               // s = 0: lpv / lnv, s = 1: rpv / rnv
            mate[s] = mate_record(idx[s] + off[s][r], off[s][r + 1] - off[s][r], idx[2 + s] + off[2 + s][r],
                                  off[2 + s][r + 1] - off[2 + s][r], neg_keep, scratch);
        int len = 0, n_ent = 0;
        bool raw = !(mate[0].ok && mate[1].ok);
        if (!raw) {
            len = mate_units(mate[0]) + mate_units(mate[1]);
            n_ent = pair_entries(mate[0], mate[1], neg_keep);
            raw = len > 255 || n_ent > 255;
        }
        if (raw) {
            n_ent = ent_off[r + 1] - ent_off[r];
            if (n_ent > 255) {
                gk_set_error("gk_wire_encode: read pair %lld has %d observation entries (limit 255)", (long long)r, n_ent);
                return -1;
            }
            len = 5 * n_ent;
        }
        if (stream != nullptr) {
            if (units + len > capacity) {
                gk_set_error("gk_wire_encode: stream capacity %lld exceeded", (long long)capacity);
                return -1;
            }
            uint16_t* out = stream + units;
            if (raw) {
                for (int e = 0; e < n_ent; ++e) {
                    const int64_t g = ent_off[r] + e;
                    if (ent_word[g] > 65535) {
                        gk_set_error("gk_wire_encode: variant word %d beyond the 16-bit wire field", ent_word[g]);
                        return -1;
                    }
                    out[5 * e] = (uint16_t)ent_word[g];
                    out[5 * e + 1] = (uint16_t)(ent_pos[g] & 0xffffu);
                    out[5 * e + 2] = (uint16_t)(ent_pos[g] >> 16);
                    out[5 * e + 3] = (uint16_t)(ent_neg[g] & 0xffffu);
                    out[5 * e + 4] = (uint16_t)(ent_neg[g] >> 16);
                }
            } else {
                write_mate(mate[0], out);
                write_mate(mate[1], out + mate_units(mate[0]));
            }
            hdr[r] = (uint16_t)(n_ent | ((raw ? 0 : len) << 8));
        }
        units += len;
        entries += n_ent;
    }
    if (n_entries_out != nullptr) *n_entries_out = entries;
    return units;
}

// ---------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------
namespace {

constexpr int kExpandThreads = GK_LIK_READS;      // one thread per read of the tile
constexpr int kStageUnits = 4096;                 // 8 KB of the tile's records staged in shared memory

// 32 bits of a little-endian bit string stored in uint16 units, starting at bit `start` (may be
// negative: bits before the string are zero); bits at or beyond `nbits` are zero.
__device__ __forceinline__ uint32_t bitmap_window(const uint16_t* bm, int nbits, int start) {
    uint32_t out = 0u;
    const int n_units = (nbits + 15) >> 4;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int u = (start >> 4) + k;           // arithmetic shift: floor for negative starts
        if (u < 0 || u >= n_units) continue;
        const uint32_t val = bm[u];
        const int shift = 16 * u - start;         // position of unit u's bit 0 in the output
        if (shift >= 32 || shift <= -16) continue;
        out |= shift >= 0 ? val << shift : val >> (-shift);
    }
    return out;
}

struct EntrySink {
    uint4* entries;              // {word * row stride of mem in bytes, pos, neg, 1 << 8 (r & 3)}
    int64_t at;
    uint32_t stride_bytes, mult;
    __device__ __forceinline__ void put(int w, uint32_t p, uint32_t n) {
        entries[at++] = make_uint4((uint32_t)w * stride_bytes, p, n, mult);
    }
};

// One mate's record, decoded lazily word by word.
struct MateView {
    const uint16_t* bm;
    const uint16_t* holes;
    const uint16_t* outs;
    int lo, n, n_out, n_hole, w_next, w_last, units;

    __device__ __forceinline__ void open(const uint16_t* rec) {
        lo = rec[0];
        const int x = rec[1];
        n = x & 255;
        n_out = (x >> 8) & 15;
        n_hole = x >> 12;
        bm = rec + 2;
        holes = bm + ((n + 15) >> 4);
        outs = holes + ((n_hole + 1) >> 1);
        units = 2 + ((n + 15) >> 4) + ((n_hole + 1) >> 1) + n_out;
        w_next = n > 0 ? lo >> 5 : INT32_MAX;
        w_last = n > 0 ? (lo + n - 1) >> 5 : -1;
    }
    // positive / negative bits of window word w_next, then advance
    __device__ __forceinline__ void take(const uint32_t* __restrict__ keep, uint32_t& posb, uint32_t& negb) {
        const int w = w_next;
        const int start = 32 * w - lo;            // window offset of bit 0 of word w
        const int b0 = start < 0 ? -start : 0;    // bits of [lo, lo + n) inside word w: [b0, b1)
        const int b1 = (n - start) < 32 ? (n - start) : 32;
        const uint32_t win = (b1 >= 32 ? 0xffffffffu : ((1u << b1) - 1u)) & ~((1u << b0) - 1u);
        posb = bitmap_window(bm, n, start) & win;
        uint32_t holeb = 0u;
        for (int h = 0; h < n_hole; ++h) {
            const int rel = (int)((holes[h >> 1] >> (8 * (h & 1))) & 255) - start;
            if (rel >= 0 && rel < 32) holeb |= 1u << rel;
        }
        negb = win & __ldg(keep + w) & ~posb & ~holeb;
        w_next = w < w_last ? w + 1 : INT32_MAX;
    }
};

// Entries of a read pair from its two mate records: window words merged in ascending order - what both
// mates observe goes to a second entry so that it counts twice - then one entry per outside positive.
__device__ void expand_pair(const uint16_t* rec, const uint32_t* __restrict__ keep, EntrySink& sink) {
    MateView l, r;
    l.open(rec);
    r.open(rec + l.units);
    while (l.w_next != INT32_MAX || r.w_next != INT32_MAX) {
        const int w = l.w_next < r.w_next ? l.w_next : r.w_next;
        uint32_t lp = 0u, ln = 0u, rp = 0u, rn = 0u;
        if (l.w_next == w) l.take(keep, lp, ln);
        if (r.w_next == w) r.take(keep, rp, rn);
        const uint32_t ov = (lp | ln) & (rp | rn);
        const uint32_t p = lp | (rp & ~ov), n = ln | (rn & ~ov);
        if (p | n) sink.put(w, p, n);
        if (ov) sink.put(w, rp & ov, rn & ov);
    }
    for (int i = 0; i < l.n_out; ++i) sink.put(l.outs[i] >> 5, 1u << (l.outs[i] & 31), 0u);
    for (int i = 0; i < r.n_out; ++i) sink.put(r.outs[i] >> 5, 1u << (r.outs[i] & 31), 0u);
}

// Fast path of expand_pair for the usual read pair: both windows of at most 32 variants, no outside
// positives.  A mate then touches word w0 and at most w0 + 1, and its bits are two 64-bit shifts; the
// pair's words are merged without loops.  Emits exactly the entries of expand_pair, in the same order.
struct MateSmall {
    int w0;                      // first word of the window (undefined when n == 0)
    uint32_t p0, n0, p1, n1;     // positive / negative bits of words w0 and w0 + 1
    int n, units;
};

__device__ __forceinline__ bool mate_is_small(const uint16_t* rec) {
    const int x = rec[1];
    return (x & 255) <= 32 && ((x >> 8) & 15) == 0;
}

__device__ __forceinline__ MateSmall open_small(const uint16_t* rec, const uint32_t* __restrict__ keep) {
    MateSmall m;
    const int lo = rec[0];
    const int x = rec[1];
    m.n = x & 255;
    const int n_hole = x >> 12;
    const int nb = (m.n + 15) >> 4;
    m.units = 2 + nb + ((n_hole + 1) >> 1);
    m.w0 = lo >> 5;
    m.p0 = m.n0 = m.p1 = m.n1 = 0u;
    if (m.n == 0) return m;
    const int s = lo & 31;
    const uint32_t bitmap = (uint32_t)rec[2] | (nb > 1 ? (uint32_t)rec[3] << 16 : 0u);
    const unsigned long long win = (m.n == 32 ? 0xffffffffull : ((1ull << m.n) - 1ull)) << s;
    const unsigned long long pos = ((unsigned long long)bitmap << s) & win;
    unsigned long long hole = 0ull;
    const uint16_t* holes = rec + 2 + nb;
    for (int h = 0; h < n_hole; ++h) hole |= 1ull << (((holes[h >> 1] >> (8 * (h & 1))) & 255) + s);
    const unsigned long long keep64 = (unsigned long long)__ldg(keep + m.w0) |
                                      ((win >> 32) ? (unsigned long long)__ldg(keep + m.w0 + 1) << 32 : 0ull);
    const unsigned long long neg = win & keep64 & ~pos & ~hole;
    m.p0 = (uint32_t)pos;
    m.n0 = (uint32_t)neg;
    m.p1 = (uint32_t)(pos >> 32);
    m.n1 = (uint32_t)(neg >> 32);
    return m;
}

__device__ __forceinline__ void expand_pair_small(const uint16_t* rec, const uint32_t* __restrict__ keep,
                                                  EntrySink& sink) {
    const MateSmall l = open_small(rec, keep);
    const MateSmall r = open_small(rec + l.units, keep);
    // the words to visit, ascending: three consecutive ones when the windows are within a word of each
    // other (or a mate is empty), else the two words of the lower mate followed by those of the upper
    const bool l_on = l.n > 0, r_on = r.n > 0;
    const int lw = l_on ? l.w0 : r.w0, rw = r_on ? r.w0 : l.w0;
    const int a = lw < rw ? lw : rw, b = lw < rw ? rw : lw;
    const bool close = b - a <= 1;
    int words[4] = {a, a + 1, close ? a + 2 : b, close ? -1 : b + 1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int w = words[k];
        if (w < 0) continue;
        uint32_t lp = 0u, ln = 0u, rp = 0u, rn = 0u;
        if (l_on && w == l.w0) { lp = l.p0; ln = l.n0; }
        if (l_on && w == l.w0 + 1) { lp = l.p1; ln = l.n1; }
        if (r_on && w == r.w0) { rp = r.p0; rn = r.n0; }
        if (r_on && w == r.w0 + 1) { rp = r.p1; rn = r.n1; }
        const uint32_t ov = (lp | ln) & (rp | rn);
        const uint32_t p = lp | (rp & ~ov), n = ln | (rn & ~ov);
        if (p | n) sink.put(w, p, n);
        if (ov) sink.put(w, rp & ov, rn & ov);
    }
}

__global__ void __launch_bounds__(kExpandThreads, 12)
gk_expand_reads_kernel(const GkMatrix* __restrict__ matrices, const GkExpandItem* __restrict__ items,
                       const uint16_t* __restrict__ hdr_pool, const uint16_t* __restrict__ stream,
                       const uint32_t* __restrict__ keep_pool, int32_t* __restrict__ entoff_pool,
                       uint4* __restrict__ entries) {
    __shared__ __align__(16) uint16_t s_rec[kStageUnits];
    __shared__ int s_warp_len[kExpandThreads / 32], s_warp_ent[kExpandThreads / 32];
    const GkExpandItem item = items[blockIdx.x];
    const GkMatrix M = matrices[item.matrix];
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int r = item.r0 + t;
    const bool on = r < M.n_reads;
    int n_ent = 0, len = 0;
    bool raw = false;
    if (on) {
        const int h = hdr_pool[item.hdr_base + r];
        n_ent = h & 255;
        len = h >> 8;
        raw = len == 0;
        if (raw) len = 5 * n_ent;
    }
    // exclusive scans of the record lengths and entry counts over the tile
    int inc_len = len, inc_ent = n_ent;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, inc_len, o);
        const int b = __shfl_up_sync(0xffffffffu, inc_ent, o);
        if (lane >= o) {
            inc_len += a;
            inc_ent += b;
        }
    }
    if (lane == 31) {
        s_warp_len[warp] = inc_len;
        s_warp_ent[warp] = inc_ent;
    }
    __syncthreads();
    int base_len = 0, base_ent = 0, total_len = 0;
#pragma unroll
    for (int w = 0; w < kExpandThreads / 32; ++w) {
        if (w < warp) {
            base_len += s_warp_len[w];
            base_ent += s_warp_ent[w];
        }
        total_len += s_warp_len[w];
    }
    const int my_rec = base_len + inc_len - len;                   // units from the tile's first record
    const int64_t my_ent = (int64_t)item.ent_off + base_ent + inc_ent - n_ent;
    int32_t* eoff = entoff_pool + M.entoff_off;
    if (on) {
        eoff[r] = (int32_t)my_ent;
        if (r == M.n_reads - 1) eoff[r + 1] = (int32_t)(my_ent + n_ent);
    } else if (M.n_reads == 0 && t == 0 && item.r0 == 0) {
        eoff[0] = (int32_t)item.ent_off;
    }
    // stage the tile's records (coalesced) when they fit
    const uint16_t* g_rec = stream + item.stream_off;
    const bool staged = total_len <= kStageUnits;
    if (staged) {
        for (int i = t; i < total_len; i += kExpandThreads) s_rec[i] = g_rec[i];
    }
    __syncthreads();
    if (!on || n_ent == 0) return;
    const uint16_t* rec = (staged ? s_rec : g_rec) + my_rec;
    EntrySink sink{entries, my_ent, (uint32_t)(M.n_ablk * M.a_tile) * 4u, 1u << (8 * (r & 3))};
    if (raw) {
        for (int e = 0; e < n_ent; ++e)
            sink.put(rec[5 * e], (uint32_t)rec[5 * e + 1] | ((uint32_t)rec[5 * e + 2] << 16),
                     (uint32_t)rec[5 * e + 3] | ((uint32_t)rec[5 * e + 4] << 16));
        return;
    }
    const uint32_t* keep = keep_pool + item.keep_off;
    if (mate_is_small(rec)) {
        const int x = rec[1];
        const uint16_t* rec_r = rec + 2 + (((x & 255) + 15) >> 4) + (((x >> 12) + 1) >> 1);
        if (mate_is_small(rec_r)) {
            expand_pair_small(rec, keep, sink);
            return;
        }
    }
    expand_pair(rec, keep, sink);
}

}  // namespace

extern "C" int gk_expand_reads(const GkMatrix* matrices, const GkExpandItem* items, int n_items,
                               const uint16_t* hdr_pool, const uint16_t* stream, const uint32_t* keep_pool,
                               int32_t* entoff_pool, void* entries, void* stream_handle) {
    if (n_items <= 0) return 0;
    GK_REQUIRE(((uintptr_t)entries & 15) == 0, "gk_expand_reads: the entry pool must be 16-byte aligned");
    gk_expand_reads_kernel<<<n_items, kExpandThreads, 0, (cudaStream_t)stream_handle>>>(
        matrices, items, hdr_pool, stream, keep_pool, entoff_pool, reinterpret_cast<uint4*>(entries));
    GK_CHECK_LAUNCH("gk_expand_reads");
    return 0;
}
