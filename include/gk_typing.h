/*
 * gk_typing.h -- C ABI of the B200 allele-typing core (libgk_typing.so).
 *
 * The reference (linnil1/KIR_graph) has no FFI layer: its typing core is NumPy
 * code inside graphkir/typing_mulit_allele.py and graphkir/typing_em.py.  Each
 * entry point below names the reference expression it replaces; a reference
 * maintainer binds them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the parameter name starts with h_;
 *   - every launcher takes the CUDA stream (cudaStream_t passed as void*) and
 *     is asynchronous on it; return value 0 = ok, negative = error, message via
 *     gk_last_error() (thread local);
 *   - "matrix"  = the likelihood data of one (sample, gene): mismatch counts
 *     m[r, a] in two layouts plus per-allele column sums;
 *     "search"  = the state of one greedy top-N search over a matrix (several
 *     searches may share a matrix: exon-first runs one per tied exon result);
 *   - all per-problem arrays live in pooled device buffers; the descriptor
 *     tables hold element offsets into those pools.
 *
 * Integer formulation: log10 P(read r | allele a) = (K_r - m) log10(.999) +
 * m log10(.001) with m[r, a] the number of the read pair's K_r variant
 * observations that disagree with allele a, so every score below is an exact
 * integer count of mismatches; conversion to log10 units happens on the host.
 */
#ifndef GK_TYPING_H
#define GK_TYPING_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GK_MAX_CN 8        /* alleles per set (copy number) supported by the search kernels */
#define GK_KB 64           /* kept-set block width of P                                      */
#define GK_RT 32           /* read rows per row block of L and P = one shared-memory stage of
                              the scoring kernel                                             */
#define GK_LIK_READS 128   /* read rows per CTA of the likelihood kernel                     */

/* Likelihood data of one gene problem.  Offsets are in elements of the pool type. */
typedef struct GkMatrix {
    int64_t mem_off;     /* uint32 pool: mem[w * (n_ablk * a_tile) + a], bit b = allele a carries variant 32w+b;
                            rows are padded with zeros to whole allele blocks (aligned vector loads); a
                            multiple of 4 */
    int64_t entoff_off;  /* int32 pool : n_reads+1 entry offsets (absolute indices into the entry pool)  */
    int64_t L_off;       /* 4-byte pool, row-blocked: L[r_blk][a_blk][GK_RT][a_tile] with r_blk = r / GK_RT,
                            a_blk = a / a_tile, i.e. element
                            ((r_blk * n_ablk + a_blk) * GK_RT + r % GK_RT) * a_tile + a % a_tile
                            = m[r, a] as float32, or as the 16-bit pair (m, m) in packed mode.  The
                            GK_RT rows of adjacent allele blocks are contiguous, so a scoring stage is
                            ONE bulk copy                                                               */
    int64_t LT_off;      /* uint8 pool : LT[a * r_pad + r] = m[r, a]                                     */
    int64_t col_off;     /* uint64 pool: colsum[a] = sum_r m[r, a]                                       */
    int32_t n_reads;
    int32_t n_alleles;
    int32_t n_words;
    int32_t r_pad;       /* multiple of 128; rows >= n_reads are zero                                    */
    int32_t a_tile;      /* allele block width of L (32)                                                 */
    int32_t n_ablk;      /* ceil(n_alleles / a_tile)                                                     */
    int32_t n_reads_total; /* reads of the whole problem when its reads are sharded over ranks (the
                            denominator of the read fractions); = n_reads otherwise                      */
    int32_t m_max;       /* upper bound of every m[r, a] of this matrix (the largest K_r; m <= K_r), or 0 when
                            unknown.  Below 128 the tie-counting and P kernels take their cheaper byte
                            predicates (gk_rescore.cu); it never changes a result                         */
} GkMatrix;

/* State of one search.  Strides are fixed by (top_n, GK_MAX_CN). */
typedef struct GkSearch {
    int64_t P_off;       /* float (uint16 in packed mode) pool, row-blocked like L:
                            P[r_blk][k_blk][GK_RT][GK_KB], element
                            ((r_blk * n_kblk + k / GK_KB) * GK_RT + r % GK_RT) * GK_KB + k % GK_KB
                            = min over the members of kept set k of m[r, id]                             */
    int64_t S_off;       /* uint32 pool: S[k * s_stride + a]                                             */
    int64_t cand_off;    /* int32 pool : candidate allele ids of the current step                        */
    int64_t flag_off;    /* uint32 pool: per flat candidate k * n_cand + j its min-sum score, or 0xffffffff
                            when an earlier candidate yields the same allele set                          */
    int64_t alive_off;   /* int32 pool : flat candidates that can still reach the final top_n            */
    int64_t cnt_off;     /* uint32 pool: cnt[(f * n + t) * n + (q-1)] tie-split counts per alive set     */
    int32_t matrix;      /* index into the GkMatrix table                                                */
    int32_t n_cand;
    int32_t s_stride;    /* n_ablk * a_tile of the matrix                                                */
    int32_t alive_cap;
    int32_t n_kblk;      /* kept-set blocks allocated in P (ceil(top_n / GK_KB))                         */
    int32_t pad;
} GkSearch;

/* Work items (built by the host per launch). */
typedef struct GkLikItem { int32_t matrix, a_blk, r0, flags; } GkLikItem;            /* up to 4 a-blocks from a_blk; r0 multiple of GK_LIK_READS;
    flags bit 0 (GK_LIK_COLSUM_ONLY): accumulate colsum only, write neither L nor LT (a problem that is
    typed with one step - CN 1 or the homozygous shortcut - never reads them) */
#define GK_LIK_COLSUM_ONLY 1
typedef struct GkScoreItem { int32_t search, k_blk, a_blk, r0, r1, shape; } GkScoreItem; /* [r0, r1) multiple of GK_RT;
    FP32 path: shape = row mode | column mode << 8; modes: 0 = 128 wide, 1 = 64, 2 = 16, 3 = 32, 4 = 48
    (from k_blk / a_blk).  Packed path, full-width tile: rows 5..8 = 32, 64, 96, 128 kept sets, column mode 0
    (128 alleles).  Packed path, warp-split tile (small genes, ragged right edge):
    shape = G' | log2(WK) << 4 | TA' << 8 | (row offset / 8) << 20 | GK_SHAPE_WARP_SPLIT = WK * 8 G' kept sets
    (G' = 1..4, WK = 1, 2, 4) x 8 TA' alleles (TA' = 1..8), starting `row offset` (0, 8, .. 56) rows into
    k-block k_blk; offset + rows <= 2 GK_KB */
#define GK_SHAPE_WARP_SPLIT (1 << 16)
typedef struct GkExpandItem { int32_t matrix, r0, hdr_base, keep_off; uint32_t stream_off, ent_off; } GkExpandItem;
    /* one tile of GK_LIK_READS reads of the wire format (gk_expand_reads): hdr_base = index of the matrix's
       first read in the header pool, keep_off = offset of the gene's neg_keep words, stream_off = first
       record of the tile (uint16 units), ent_off = index of its first entry in the entry pools */
typedef struct GkCountItem { int32_t search, f0, r0, r1; } GkCountItem;              /* 8 alive sets from f0; r multiple of 16; */
typedef struct GkPItem { int32_t search, k_blk, r0, r1; } GkPItem;                   /* one k-block x reads [r0, r1), multiples of 128 */

/* Per-search step outputs (device arrays indexed [search]). */
typedef struct GkStepInfo {
    int32_t n_kept;      /* sets kept after this step (<= top_n)                                         */
    int32_t n_unique;    /* distinct candidate sets (N_uniq of typing_mulit_allele.py:561-567)           */
    int32_t n_alive;     /* sets rescored                                                                */
    int32_t cut;         /* M = max(top_n, n_unique / 5)                                                 */
    uint32_t bar;        /* score of the top_n-th unique candidate                                       */
    int32_t tie_flags;   /* bit0: the tie group of the top_n-th score straddles the M cut, bit1: a tie group
                            straddles the final top_n cut,
                            bit2: rank 0 and rank 1 share the score, bit3: for a rank selectBest looks at,
                            a member's fraction is below 1/(2n) counting only the reads it wins alone and
                            reaches it counting its tied reads in full (the reference's float fractions
                            fall anywhere in between, typing_mulit_allele.py:81-96, :575-580)            */
    int32_t best_rank;   /* first rank whose every member fraction >= 1/(2n), else 0
                            (TypingResult.selectBest, typing_mulit_allele.py:63-103), exact integers    */
    int32_t pad1;
} GkStepInfo;

const char* gk_last_error(void);
int gk_abi_version(void);
int gk_sizeof(const char* struct_name);     /* sizeof(GkMatrix) etc., for binding self-checks */

/* (a) likelihood build: replaces AlleleTyping.reads2AlleleProb + np.log10
 *     (graphkir/typing_mulit_allele.py:340-381, :263), fed by hisat2.py's per-read
 *     positive/negative variant ids.  Writes L, LT and accumulates colsum
 *     (colsum must be zeroed by the caller). */
int gk_likelihood(const GkMatrix* matrices, const GkLikItem* items, int n_items,
                  const uint32_t* mem_pool, const int32_t* entoff_pool,
                  const void* entries /* 16 bytes per observation entry: uint32 {word * row stride of mem in
                     bytes, positive bits, negative bits, 1 << 8 (r & 3)} with r the read's index in its
                     matrix; 16-byte aligned */,
                  float* L_pool, uint8_t* LT_pool, unsigned long long* col_pool, int half_mode, void* stream);

/* Wire format of the read observations (host -> device) and its expansion; layout in csrc/gk_wire.cu.
 *     What graphkir/hisat2.py's getPNFromVariantList (:716-800) decides per mate - the window of the
 *     variant table whose variants are negatives unless positive or excluded - is shipped as
 *     {lo, n, bitmap of positives, excluded offsets, outside positives} (14 B per read pair on cfg3)
 *     instead of the 36 B of observation entries, and gk_expand_reads rebuilds the entry offsets and
 *     the entries (the inputs of gk_likelihood) on the device.
 * gk_wire_encode (host): off / idx = CSR lists in the order lpv, rpv, lnv, rnv; neg_keep = per word
 *     of the gene the variants that occur as a negative in any read; ent_* = the canonical entries
 *     (gk_pack_entries), copied into the raw records of reads the window form cannot express.
 *     hdr gets one uint16 per read, stream the records.  With stream == NULL only the sizes are
 *     computed.  Returns the stream length in uint16 units (or -1), *n_entries_out = entries the
 *     expansion will emit. */
int64_t gk_wire_encode(int64_t n_reads, const int64_t* const* off, const int32_t* const* idx,
                       const uint32_t* neg_keep, const int32_t* ent_off, const int32_t* ent_word,
                       const uint32_t* ent_pos, const uint32_t* ent_neg, uint16_t* hdr, uint16_t* stream,
                       int64_t capacity, int64_t* n_entries_out);
int gk_expand_reads(const GkMatrix* matrices, const GkExpandItem* items, int n_items, const uint16_t* hdr_pool,
                    const uint16_t* stream_pool, const uint32_t* keep_pool, int32_t* entoff_pool,
                    void* entries /* as gk_likelihood reads them */, void* stream);

/* CN = 1 step: replaces log_probs[:, idx].sum(0) + argsort()[::-1][:top_n]
 *     (typing_mulit_allele.py:512-532).  One CTA per search.  Order: (colsum, position). */
int gk_first_step(const GkMatrix* matrices, const GkSearch* searches, int n_search, int top_n,
                  const unsigned long long* col_pool, const int32_t* cand_pool,
                  int32_t* ids_out, uint32_t* score_out, uint32_t* cnt_out, int32_t* flat_out,
                  GkStepInfo* info, int32_t* kept_count, void* stream);

/* (b) max-then-sum candidate scoring: replaces
 *     np.maximum(log_probs[:, idx], prev.T[:, :, None]).sum(axis=1)   (:540-542).
 *     FP32 path (half_mode = 0): accumulates D[k, a] += sum_{r in item} |L[r, a] - P[r, k]| into
 *     S_pool (zeroed by the caller); the min-sum score is (colsum[a] + score_prev[k] - D[k, a]) / 2,
 *     formed by gk_select / gk_rank (score_prev = score_out of the previous step = sum_r P[r, k]).
 *     Packed path (half_mode = 1: L as 16-bit pairs, P as uint16): accumulates the min-sum itself
 *     with VIMNMX.U16x2 + IMAD on 16-bit lanes that are added to S_pool every flush_stages stages
 *     of GK_RT reads: flush_stages * GK_RT * (largest m of the batch) must be <= 65535.  Pass
 *     s_is_minsum = 1 to gk_select / gk_rank. */
int gk_score(const GkMatrix* matrices, const GkSearch* searches, const GkScoreItem* items, int n_items,
             const float* L_pool, const void* P_pool, uint32_t* S_pool, int half_mode, int flush_stages,
             const int32_t* kept_count /* optional: skip tiles whose first row is >= kept_count[search] */,
             void* stream);

/* (c) segmented selection, part 1: canonical-key dedup (uniqueAllele, :456-476, :551-563),
 *     N_uniq, the cut max(top_n, N_uniq // 5) (:567) and the list of candidates that can
 *     still reach the final top_n.  Dedup runs in slices of 8192 (few searches: 2048) candidates (many CTAs per search),
 *     the cut and the compaction in one CTA per search. */
int gk_select(const GkMatrix* matrices, const GkSearch* searches, int n_search, int top_n, int n_prev,
              int max_alleles, int max_cand, const int32_t* kept_count, const int32_t* ids_prev,
              const int32_t* cand_pool, const uint32_t* S_pool, const unsigned long long* col_pool,
              const uint32_t* score_prev, uint32_t* val_pool, int32_t* alive_pool, GkStepInfo* info,
              int s_is_minsum /* S holds the min-sum itself (packed scoring path) */, void* stream);

/*     rescoring of the alive sets: replaces log_probs[:, ids].max(2) / np.equal / belong_norm
 *     (:569-580) with integer tie-split counts (cnt zeroed by the caller). */
int gk_rescore_count(const GkMatrix* matrices, const GkSearch* searches, const GkCountItem* items,
                     int n_items, int top_n, int n_set, const GkStepInfo* info,
                     const int32_t* ids_prev, const int32_t* cand_pool, const int32_t* alive_pool,
                     const uint8_t* LT_pool, uint32_t* cnt_pool, void* stream);

/*     part 2: 3-key ranking (rankScore / sortByScoreAndEveness, :156-171, :202-214) on exact
 *     integers (score, sum of member column sums, unevenness, flat index); keeps top_n. */
int gk_rank(const GkMatrix* matrices, const GkSearch* searches, int n_search, int top_n, int n_set,
            const int32_t* ids_prev, const int32_t* cand_pool, const int32_t* alive_pool,
            const uint32_t* S_pool, const uint32_t* cnt_pool, const unsigned long long* col_pool,
            const uint32_t* score_prev, unsigned long long* key_pool /* 3 words per alive slot */, int32_t* ids_out, uint32_t* score_out, uint32_t* cnt_out, int32_t* flat_out,
            GkStepInfo* info, int32_t* kept_count_out, int s_is_minsum, void* stream);

/*     P for the next step: P[r, k] = min over members of m[r, id]  (allele_prob, :569). */
int gk_write_p(const GkMatrix* matrices, const GkSearch* searches, const GkPItem* items, int n_items,
               int top_n, int n_set, const int32_t* kept_count, const int32_t* ids,
               const uint8_t* LT_pool, void* P_pool, int half_mode, void* stream);

/* ---- EM path (graphkir/typing_em.py) ------------------------------------------------- */

/* One gene of the EM solver; offsets index the pools passed to gk_em_squarem. */
typedef struct GkEmProblem {
    int64_t row_off;     /* uint32 pool: rows[u * n_awords + w], distinct compatibility rows (allele bitsets)   */
    int64_t wgt_off;     /* uint32 pool: number of read pairs with row u                                        */
    int64_t len_off;     /* double pool: allele length normalisation (ones when seq_len is not given)           */
    int64_t out_off;     /* double pool: prob[n_alleles] then 4 * n_alleles + n_rows doubles of scratch         */
    int32_t n_rows, n_alleles, n_awords, pad;
} GkEmProblem;

/* Replaces getCandidateAllelePerRead + getMostFreqAllele (typing_em.py:68-104): compatible alleles
 * per read pair as bitsets.  membT[v * n_awords + w] holds the alleles of variant v; off_x/idx_x are
 * the CSR lists of positive / negative variants of the left / right mate.  compat has room for
 * 2 * n_reads rows; rows [0, n_reads) are the result. */
int gk_em_compat(const uint32_t* membT, int n_awords, int n_alleles, const int32_t* off_lp,
                 const int32_t* idx_lp, const int32_t* off_ln, const int32_t* idx_ln,
                 const int32_t* off_rp, const int32_t* idx_rp, const int32_t* off_rn,
                 const int32_t* idx_rn, int n_reads, uint32_t* compat, void* stream);

/* Replaces hisatEMnp (typing_em.py:107-188): SQUAREM EM, one CTA per problem, float64. */
int gk_em_squarem(const GkEmProblem* problems, int n_problems, const uint32_t* row_pool,
                  const uint32_t* wgt_pool, const double* len_pool, double* out_pool,
                  int32_t* iters_out, int iter_max, double diff_threshold, void* stream);

/* Read grouping by called alleles: replaces
 *     np.equal(probs[:, ids], probs[:, ids].max(axis=1)[:, None])      (graphkir/novel_discover.py:62-64)
 * on the device-resident likelihood of matrix `matrix`: pattern[r] bit t = allele ids[t] attains the
 * smallest mismatch count of read r among the n_ids (1..32) alleles.  pattern has n_reads entries. */
int gk_group_reads(const GkMatrix* matrices, int matrix, int n_reads, const int32_t* ids, int n_ids,
                   const uint8_t* LT_pool, uint32_t* pattern, void* stream);

/* CN model: replaces the loop of CNgroup.fit over the candidate bases and
 * calcCNGroupProb (graphkir/cn_model.py:124-204): likelihood[b] = sum_i log(max_n pdf_n(x[i]; bases[b]) * space
 * + 1e-9) * density[i], with pdf_n the normal density of copy number n (n = 0 .. max_cn - 1; means n * base,
 * deviations from base_dev, y0_dev, dev_decay, dev_decay_neg and start_base as in the reference).  float64;
 * x, density [bin_num], bases, likelihood [n_base] are device arrays; prob_out (optional, device,
 * [n_base][max_cn][bin_num]) receives the CN-group probabilities themselves. */
int gk_cn_fit(const double* x, const double* density, const double* bases, int n_base, int bin_num, int max_cn,
              int start_base, double base_dev, double y0_dev, double dev_decay, double dev_decay_neg, double space,
              double* likelihood, double* prob_out, void* stream);

/* Host-side fast path for `{prefix}.variant.json` (reference reader
 * graphkir/hisat2.py:847-866 + kir_typing.py:92-97).  No CUDA involved.  gk_json_scan walks the JSON
 * text once and keeps, for every element of "reads", backbone (interned), multiple and the four
 * variant-id lists as CSR arrays; the raw SAM text is skipped.  It returns an opaque handle (NULL on
 * error, message in gk_last_error()) and fills
 *   sizes[0]      reads
 *   sizes[1..4]   total ids of lpv, lnv, rpv, rnv
 *   sizes[5], [6] distinct id strings, their bytes     sizes[7], [8] distinct backbone strings, their bytes
 *   sizes[9], [10] byte span of the "variants" value (for the caller's JSON parser; -1 if absent)
 * gk_json_fill copies into caller-allocated arrays (off[w] has reads + 1 entries; the string tables
 * are offsets [n + 1] into a byte blob); gk_json_free releases the handle. */
void* gk_json_scan(const char* buf, int64_t len, int64_t* sizes);
int gk_json_fill(void* handle, int32_t* backbone, int32_t* multiple, int64_t* const* off,
                 int32_t* const* idx, int64_t* id_off, char* id_bytes, int64_t* gene_off, char* gene_bytes);
void gk_json_free(void* handle);

/* Host: observation entries of the likelihood kernel from the CSR lists of a gene (see GkMatrix:
 * entoff / ent_word / ent_pos / ent_neg).  off[w] / idx[w], w < 4: CSR lists of variant indices per
 * read, polarity[w] = 1 for positive lists.  Duplicates of an observation go to separate entries
 * (occurrence rank) so that multiplicities are exact; entries of a read are ordered by (rank, word).
 * The entry arrays need room for one entry per observation.  Returns the number of entries or -1. */
int64_t gk_pack_entries(int64_t n_reads, const int64_t* const* off, const int32_t* const* idx,
                        const int32_t* polarity, int32_t* ent_off, int32_t* ent_word, uint32_t* ent_pos,
                        uint32_t* ent_neg, int32_t* k_obs);

/* Host: walk of one SAM record, CIGAR x MD x Zs -> match / single / insertion / deletion segments
 * (replaces recordToRawVariant + readZs + readMd, graphkir/hisat2.py:279-538).
 * seg: int32 [max_seg][7] = typ (0 match, 1 single, 2 insertion, 3 deletion), 0-based backbone pos,
 * length, value span (offset, length in `line`; length -2 = none, -1 = the value is `length`), id
 * span (length -2 = none, -1 = "unknown").  meta[0], meta[1] = head / tail soft clip, meta[2], meta[3]
 * = span of the backbone name.  Returns the number of segments, or -3 splicing (N), -4 unsupported
 * CIGAR operation, -5 inconsistent record (the reference's asserts), -6 index out of range,
 * -7 malformed number or Zs item, -8 more than max_seg segments. */
int gk_sam_walk(const char* line, int64_t len, int32_t* seg, int max_seg, int32_t* meta);

/* Host: name-sorted SAM text -> per read pair the positive / negative variant lists as CSR arrays,
 * no Python object per record (replaces the loop of extractVariantFromBam
 * with error_correction = False: readPair graphkir/hisat2.py:228-276, filterRead :541-578,
 * recordToVariants :657-689, findVariantId :581-606, getVariantsBoundary :692-713,
 * getPNFromVariantList :716-800, extractVariant :803-844).
 * Variant table (sorted as the index is, Variant.__lt__ msa2hisat.py:48-53): v_ref = index into the
 * name table (ref_off / ref_bytes, n_ref names, ids ascending along the table), v_typ = 0 insertion,
 * 1 single, 2 deletion, v_val_int = deletion length, v_val_off [n_var + 1] / v_val_bytes = base or
 * inserted sequence.  novel_id = Variant.novel_id at entry; num_editdist = filterRead's bound.
 * gk_sam_extract returns a handle (never null) and fills sizes[12]:
 *   [0] pairs kept, [1..4] ids in lpv / lnv / rpv / rnv, [5] novel variants, [6] bytes of their
 *   values, [7] backbone names (the caller's, then new ones), [8] bytes of the names,
 *   [9] status (0, or a gk_sam_walk code for the record at line [11]), [10] pairs whose flags are
 *   not first + second mate ("strange case"), [11] 1-based line of the failing record or -1.
 * Indices in idx[w] < n_var name table variants, the others novel variant (index - n_var), whose id
 * is "nv<novel_id + index - n_var>".  span: int64 [pairs][4] = offset, length of the left and of the
 * right record in `sam`.  gk_sam_extract_fill copies into caller-allocated arrays (off[w] has
 * pairs + 1 entries); gk_sam_extract_free releases the handle. */
void* gk_sam_extract(const char* sam, int64_t sam_len, int32_t n_var, const int32_t* v_ref,
                     const int32_t* v_pos, const int32_t* v_typ, const int32_t* v_val_int,
                     const int32_t* v_length, const int64_t* v_val_off, const char* v_val_bytes,
                     int32_t n_ref, const int64_t* ref_off, const char* ref_bytes, int32_t novel_id,
                     int32_t num_editdist, int64_t* sizes);
int gk_sam_extract_fill(void* handle, int32_t* multiple, int32_t* backbone, int64_t* span,
                        int64_t* const* off, int32_t* const* idx, int32_t* nv_ref, int32_t* nv_pos,
                        int32_t* nv_typ, int32_t* nv_val_int, int32_t* nv_length, int64_t* nv_val_off,
                        char* nv_val_bytes, int64_t* ref_off, char* ref_bytes);
void gk_sam_extract_free(void* handle);
/* The "reads" array of the reference's {prefix}.json (writeReadsAndVariantsData, hisat2.py:847-857)
 * for the pairs of a gk_sam_extract handle, byte for byte what json.dump writes for
 * [asdict(PairRead), ...] without the enclosing brackets: {"l_sam": ..., "r_sam": ..., "multiple": ...,
 * "backbone": ..., "lpv": [...], "lnv": [...], "rpv": [...], "rnv": [...]} joined by ", ", strings
 * escaped as ensure_ascii does.  sam = the text given to gk_sam_extract; id_off [n_var + 1] /
 * id_bytes = ids of the table variants; novel variants are written as "nv<novel_id + k>".  *out points
 * into the handle (valid until the next call on it or gk_sam_extract_free).  -1 if a string is not
 * UTF-8. */
int gk_sam_extract_json(void* handle, const char* sam, const int64_t* id_off, const char* id_bytes,
                        int32_t novel_id, const char** out, int64_t* out_len);

/* Host: work-item tables of the search kernels (csrc/gk_plan.cu; no CUDA calls).  They restate what
 * kir_graph_b200/engine.py builds with NumPy per copy-number step (the definition, and the fallback for the
 * FP32 scoring path, restricted candidate lists and candidate-column sharding), row for row, for hosts that
 * type a new cohort every pass.
 * gk_plan_score_tiles / gk_plan_score_items: packed scoring path.  A / kept / r16 (reads rounded up to GK_RT)
 *     per live search; cut = int32 [17][3][2]: for a remainder of g = 1..16 row groups of 8 kept sets up to
 *     three warp-split pieces (G', log2 WK), G' = 0 ends the list (engine.SearchGroup._W_CUT).  _tiles counts
 *     the (row piece x column tile) pairs per search (the caller picks the read chunk from them); _items
 *     writes every tile x every chunk of `chunk` reads as GkScoreItem with .search = search_id[j], largest
 *     item first (stable).
 * gk_plan_grid_items: rows {search_id[j], i * scale, r0, r1} for i < count[j] and the chunks [r0, r1) of
 *     `chunk` reads over extent[j], the chunk varying fastest: GkCountItem (count = ceil(alive / 8), scale 8,
 *     extent = reads rounded up to 16) and GkPItem (count = k-blocks, scale 1, extent = r_pad).
 * _items / _grid_items return the number of rows, or -(rows needed) when `cap` is too small. */
int64_t gk_plan_score_tiles(int n, const int64_t* A, const int64_t* kept, const int32_t* cut,
                            int64_t* tiles_per_search);
int64_t gk_plan_score_items(int n, const int32_t* search_id, const int64_t* A, const int64_t* kept,
                            const int64_t* r16, int64_t chunk, const int32_t* cut, GkScoreItem* out, int64_t cap);
int64_t gk_plan_grid_items(int n, const int32_t* search_id, const int64_t* count, int32_t scale,
                           const int64_t* extent, int64_t chunk, int32_t* out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* GK_TYPING_H */
