// Host-side work-item tables of the search kernels (pure host code, no CUDA calls).
//
// Per copy-number step engine.SearchGroup hands gk_score, gk_rescore_count and gk_write_p a table of work
// items: tiles of (kept sets x candidate alleles) or (alive sets) or (k-blocks), each cut into chunks of
// reads.  The tables are a few 10^4 rows and depend on the read counts of the batch, so a cohort that is
// typed once (bench.py: e2e_cold) builds them for every pass; as vectorised NumPy that is 11 ms per
// 96-sample pass on one host thread - as long as the whole GPU pass.  These routines produce the same
// tables, row for row (tests/test_plan_native.py compares them with the NumPy statements in engine.py,
// which remain the definition and the fallback for the cases not covered here: the FP32 scoring path,
// restricted candidate lists, candidate-column sharding).
#include <algorithm>
#include <cstdint>
#include <vector>

#include "gk_common.cuh"

namespace {

constexpr int kMaxGroups = 16;      // row groups (of 8 kept sets) in a remainder below 128 rows
constexpr int kMaxPieces = 3;       // warp-split pieces a remainder is cut into

struct Tile {
    int32_t search;   // index into the caller's arrays
    int32_t k_blk, a_blk, shape;
    int64_t rows, cols;
};

// Row pieces of `k` kept sets (engine.SearchGroup._row_pieces): kind 0 = under a full-width (128-allele)
// column tile, 1 = under a warp-split column tile, 2 = the remainder below 32 rows beside a full-width tile.
// cut[g] = up to kMaxPieces pieces (G', log2 WK) covering a remainder of g row groups, G' = 0 ends the list.
template <class Emit>
inline void row_pieces(int64_t k, int kind, const int32_t* cut, Emit emit) {
    const int64_t n_full = k / 128, rem = k % 128;
    if (kind == 0) {
        for (int64_t i = 0; i < n_full; ++i) emit(128 * i, (int32_t)(4 + 4), (int64_t)128);
        if (rem / 32) emit(128 * n_full, (int32_t)(4 + rem / 32), 32 * (rem / 32));
        return;
    }
    if (kind == 2) {
        if (rem % 32) {
            const int32_t gp = (int32_t)((rem % 32 + 7) / 8);
            emit(128 * n_full + 32 * (rem / 32), gp | GK_SHAPE_WARP_SPLIT, (int64_t)8 * gp);
        }
        return;
    }
    for (int64_t i = 0; i < n_full; ++i) emit(128 * i, (int32_t)(4 | (2 << 4) | GK_SHAPE_WARP_SPLIT), (int64_t)128);
    int64_t at = 128 * n_full;
    const int64_t g = (rem + 7) / 8;
    if (g < 1 || g > kMaxGroups) return;
    const int32_t* pieces = cut + g * kMaxPieces * 2;
    for (int p = 0; p < kMaxPieces && pieces[2 * p] > 0; ++p) {
        const int32_t gp = pieces[2 * p], wk = pieces[2 * p + 1];
        const int64_t rows = (int64_t)(8 * gp) << wk;
        emit(at, gp | (wk << 4) | GK_SHAPE_WARP_SPLIT, rows);
        at += rows;
    }
}

// Every (row piece x column tile) of the packed scoring path, in the order engine.SearchGroup._packed_tiles
// enumerates them: the 128-column blocks of all searches (full-width tiles), their two 64-column halves
// (row remainders), then the first and the second warp-split column tile after the full blocks.
template <class Emit>
inline void packed_tiles(int n, const int64_t* A, const int64_t* kept, const int32_t* cut, Emit emit) {
    auto pieces_of = [&](int j, int kind, int64_t a_blk, int64_t ct) {
        row_pieces(kept[j], kind, cut, [&](int64_t start, int32_t code, int64_t rows) {
            const bool split = (code & GK_SHAPE_WARP_SPLIT) != 0;
            const int32_t shape = split ? (int32_t)(code | ((int32_t)ct << 8) | ((int32_t)((start % GK_KB) / 8) << 20)) : code;
            emit(Tile{j, (int32_t)(start / GK_KB), (int32_t)a_blk, shape, rows, kind == 0 ? (int64_t)128 : 8 * ct});
        });
    };
    for (int pass = 0; pass < 3; ++pass)
        for (int j = 0; j < n; ++j) {
            const int64_t n128 = ((A[j] + 7) / 8) / 16;
            for (int64_t b = 0; b < n128; ++b) {
                if (pass == 0) pieces_of(j, 0, 4 * b, 16);
                else pieces_of(j, 2, 4 * b + (pass == 2 ? 2 : 0), 8);
            }
        }
    for (int pass = 0; pass < 2; ++pass)
        for (int j = 0; j < n; ++j) {
            const int64_t a8 = (A[j] + 7) / 8, n128 = a8 / 16, rem8 = a8 % 16;
            const int64_t first = rem8 <= 8 ? rem8 : (rem8 <= 12 ? 4 : 8);
            if (pass == 0 && rem8 > 0) pieces_of(j, 1, 4 * n128, first);
            if (pass == 1 && rem8 > 8) pieces_of(j, 1, 4 * n128 + first / 4, rem8 - first);
        }
}

}  // namespace

// Tiles per search of the packed scoring path (the caller picks the read chunk from them).
// A, kept: per live search; cut: int32 [kMaxGroups + 1][kMaxPieces][2].  Returns the total.
extern "C" int64_t gk_plan_score_tiles(int n, const int64_t* A, const int64_t* kept, const int32_t* cut,
                                       int64_t* tiles_per_search) {
    for (int j = 0; j < n; ++j) tiles_per_search[j] = 0;
    int64_t total = 0;
    packed_tiles(n, A, kept, cut, [&](const Tile& t) {
        ++tiles_per_search[t.search];
        ++total;
    });
    return total;
}

// Work items of gk_score (packed path): every tile x every chunk of `chunk` reads of its search, largest
// item first (stable).  search_id[j] = the value written to GkScoreItem.search; r16[j] = reads rounded up to
// GK_RT.  Returns the number of items, or -(needed) when `cap` is too small.
extern "C" int64_t gk_plan_score_items(int n, const int32_t* search_id, const int64_t* A, const int64_t* kept,
                                       const int64_t* r16, int64_t chunk, const int32_t* cut, GkScoreItem* out,
                                       int64_t cap) {
    std::vector<Tile> tiles;
    packed_tiles(n, A, kept, cut, [&](const Tile& t) { tiles.push_back(t); });
    int64_t total = 0;
    for (const Tile& t : tiles) total += std::max<int64_t>(1, (r16[t.search] + chunk - 1) / chunk);
    if (total > cap) return -total;
    std::vector<GkScoreItem> items((size_t)total);
    std::vector<int64_t> work((size_t)total);
    int64_t at = 0;
    for (const Tile& t : tiles) {
        const int64_t n_ch = std::max<int64_t>(1, (r16[t.search] + chunk - 1) / chunk);
        for (int64_t c = 0; c < n_ch; ++c, ++at) {
            const int64_t r0 = c * chunk, r1 = std::min((c + 1) * chunk, r16[t.search]);
            items[at] = GkScoreItem{search_id[t.search], t.k_blk, t.a_blk, (int32_t)r0, (int32_t)r1, t.shape};
            work[at] = (r1 - r0) * t.rows * t.cols;
        }
    }
    std::vector<int64_t> order((size_t)total);
    for (int64_t i = 0; i < total; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return work[a] > work[b]; });
    for (int64_t i = 0; i < total; ++i) out[i] = items[order[i]];
    return total;
}

// Work items that are a grid per search: `count[j]` leading indices (scaled by `scale`) x chunks of `chunk`
// reads over `extent[j]` - gk_rescore_count (8 alive sets x reads) and gk_write_p (k-block x reads).
// Rows are {search_id[j], index * scale, r0, r1}, searches in order, the read chunk varying fastest.
// Returns the number of rows, or -(needed) when `cap` is too small.
extern "C" int64_t gk_plan_grid_items(int n, const int32_t* search_id, const int64_t* count, int32_t scale,
                                      const int64_t* extent, int64_t chunk, int32_t* out, int64_t cap) {
    int64_t total = 0;
    for (int j = 0; j < n; ++j) total += count[j] * std::max<int64_t>(1, (extent[j] + chunk - 1) / chunk);
    if (total > cap) return -total;
    int32_t* row = out;
    for (int j = 0; j < n; ++j) {
        const int64_t n_ch = std::max<int64_t>(1, (extent[j] + chunk - 1) / chunk);
        for (int64_t i = 0; i < count[j]; ++i)
            for (int64_t c = 0; c < n_ch; ++c, row += 4) {
                row[0] = search_id[j];
                row[1] = (int32_t)(i * scale);
                row[2] = (int32_t)(c * chunk);
                row[3] = (int32_t)std::min((c + 1) * chunk, extent[j]);
            }
    }
    return total;
}
