// Rescoring kernels: stream the allele-major byte matrix LT[a][r] along reads.
//
//   gk_rescore_count   for every alive candidate set: per member t and tie size q, the
//                      number of reads where member t attains the set's minimum mismatch
//                      count together with q-1 other members.  Replaces
//                      log_probs[:, ids].max(2), np.equal(...), belong / belong.sum(2)
//                      (reference: typing_mulit_allele.py:569, :575-580) with exact integer
//                      counts; fraction[t] = sum_q cnt[t][q] / q / R is formed later.
//   gk_write_p         P[r, k] = min over the members of kept set k of m[r, id]
//                      (allele_prob for the next step, :569) in the row-blocked layout (uint16, or
//                      float for the FP32 path) the scoring kernel's TMA stages expect; columns
//                      k >= n_kept are zero.
//
// Per (read, set) they read n bytes (16-byte vector loads, coalesced along reads) and write 0 resp.
// 2 bytes; the tie counting is ALU-pipe-bound, the P writer HBM-bound.
#include "gk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// Byte-SIMD tie counting, four reads per 32-bit word.  sm_100a has no native byte-wise min /
// compare (nvcc expands __vminu4 / __vcmpeq4 into ~9 ALU instructions each, and the ALU pipe is the
// bound of this kernel: 97 % busy in ncu), so the per-byte predicates are built from carry-free
// word arithmetic and the N*N counters are accumulated with DP4A, which issues on the otherwise
// idle FMA pipe:
//     ge80(a, b)  0x80 in the bytes where a >= b        (5 ALU instructions)
//     zero80(x)   0x80 in the bytes of x that are zero  (3)
//     mn          running byte-wise minimum of the members (select through a 0xff mask)
//     eq[t]       0x01 where member t attains mn; q = sum_t eq[t] = tie size of the read
//     sel[qq]     0x01 where q == qq + 1 (bit logic on the binary digits of q)
//     cnt[t][qq] += dp4a(eq[t], sel[qq])
// Pad reads (r >= n_reads, all-zero rows: every member ties at 0) are not masked in the loop but
// subtracted from cnt[t][N-1] afterwards.
__device__ __forceinline__ uint32_t ge80(uint32_t a, uint32_t b) {
    const uint32_t t = (a | 0x80808080u) - (b & 0x7f7f7f7fu);       // bit 7: low 7 bits of a >= those of b
    return ((a & ~b) | (~(a ^ b) & t)) & 0x80808080u;
}

__device__ __forceinline__ uint32_t zero80(uint32_t x) {
    const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;             // bit 7: low 7 bits non-zero
    return ~(t | x) & 0x80808080u;
}

// Counts below 128 (m_max of the matrix, known on the host: m <= K_r) make the predicates cheap, and the
// adds can leave the ALU pipe - the bound of both kernels of this file - for the FMA pipe:
//     (a | 0x80..) - b        bit 7 of a byte = [a >= b], no borrow crosses a byte     LOP3 + IMAD
//     x + 0x7f..              bit 7 = [x != 0]                                         IMAD
//     prmt with sign replication turns bit 7 into a 0xff byte mask                     PRMT
// acc = x * one + acc with `one` read from constant memory is an IMAD (ptxas cannot fold it into an IADD3).
__constant__ uint32_t gk_one = 1u;
__constant__ uint32_t gk_minus_one = 0xffffffffu;

__device__ __forceinline__ uint32_t fma_add(uint32_t a, uint32_t b) {        // a + b on the FMA pipe
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(gk_one), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t fma_sub(uint32_t a, uint32_t b) {        // a - b on the FMA pipe
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(gk_minus_one), "r"(a));
    return d;
}
// bytes of a and b below 128: bit 7 of every byte of the result = [a >= b]
__device__ __forceinline__ uint32_t ge7(uint32_t a, uint32_t b) { return fma_sub(a | 0x80808080u, b); }
// bytes of x below 128: bit 7 of every byte of the result = [x != 0]
__device__ __forceinline__ uint32_t nz7(uint32_t x) { return fma_add(x, 0x7f7f7f7fu); }
// 0xff in the bytes whose bit 7 is set, 0x00 elsewhere
__device__ __forceinline__ uint32_t sign_mask(uint32_t x) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0u), "r"(0xba98u));
    return d;
}
// byte-wise minimum of a and b, both below 128 per byte: 3 ALU-pipe instructions + 1 IMAD
__device__ __forceinline__ uint32_t min7(uint32_t a, uint32_t b) {
    const uint32_t keep = sign_mask(ge7(b, a));                      // 0xff where a stays (b >= a)
    return (a & keep) | (b & ~keep);
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t c) {
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
}

// Two members: no minimum needed.  a < b: member 0 alone; a > b: member 1 alone; a == b: both.
__device__ __forceinline__ void count_pair(const uint8_t* row_a, const uint8_t* row_b, int r0, int r1, int n_reads,
                                           uint32_t* __restrict__ out) {
    const int lane = gk_lane();
    uint32_t n_ge = 0u, n_eq = 0u, n_all = 0u;                       // n_ge, n_eq in units of 1/128 read
    for (int r = r0 + lane * 16; r < r1; r += 32 * 16) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(row_a + r));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(row_b + r));
        const uint32_t av[4] = {a.x, a.y, a.z, a.w};
        const uint32_t bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            n_ge = __dp4a(ge80(av[w], bv[w]), 0x01010101u, n_ge);
            n_eq = __dp4a(zero80(av[w] ^ bv[w]), 0x01010101u, n_eq);
        }
        n_all += 16u;
    }
    n_ge = warp_sum(n_ge) >> 7;
    n_eq = warp_sum(n_eq) >> 7;
    n_all = warp_sum(n_all);
    if (lane == 0) {
        const int pad_lo = r0 > n_reads ? r0 : n_reads;
        const uint32_t pad = r1 > pad_lo ? (uint32_t)(r1 - pad_lo) : 0u;
        const uint32_t lt = n_all - n_ge, gt = n_ge - n_eq, eq = n_eq - pad;
        if (lt) atomicAdd(out + 0, lt);          // cnt[0][q=1]
        if (eq) atomicAdd(out + 1, eq);          // cnt[0][q=2]
        if (gt) atomicAdd(out + 2, gt);          // cnt[1][q=1]
        if (eq) atomicAdd(out + 3, eq);          // cnt[1][q=2]
    }
}

// The same for counts below 128: 4 ALU-pipe and 4 FMA-pipe instructions per word pair (7 and 2 above).
__device__ __forceinline__ void count_pair_small(const uint8_t* row_a, const uint8_t* row_b, int r0, int r1,
                                                 int n_reads, uint32_t* __restrict__ out) {
    const int lane = gk_lane();
    uint32_t n_ge = 0u, n_ne = 0u, n_all = 0u;                       // n_ge, n_ne in units of 1/128 read
#pragma unroll 2
    for (int r = r0 + lane * 16; r < r1; r += 32 * 16) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(row_a + r));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(row_b + r));
        const uint32_t av[4] = {a.x, a.y, a.z, a.w};
        const uint32_t bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            n_ge = __dp4a(ge7(av[w], bv[w]) & 0x80808080u, 0x01010101u, n_ge);
            n_ne = __dp4a(nz7(av[w] ^ bv[w]) & 0x80808080u, 0x01010101u, n_ne);
        }
        n_all += 16u;
    }
    n_ge = warp_sum(n_ge) >> 7;
    n_ne = warp_sum(n_ne) >> 7;
    n_all = warp_sum(n_all);
    if (lane == 0) {
        const int pad_lo = r0 > n_reads ? r0 : n_reads;
        const uint32_t pad = r1 > pad_lo ? (uint32_t)(r1 - pad_lo) : 0u;
        const uint32_t n_eq = n_all - n_ne;
        const uint32_t lt = n_all - n_ge, gt = n_ge - n_eq, eq = n_eq - pad;
        if (lt) atomicAdd(out + 0, lt);          // cnt[0][q=1]
        if (eq) atomicAdd(out + 1, eq);          // cnt[0][q=2]
        if (gt) atomicAdd(out + 2, gt);          // cnt[1][q=1]
        if (eq) atomicAdd(out + 3, eq);          // cnt[1][q=2]
    }
}

// Three or four members with counts below 128: count the reads of every tie PATTERN (the non-empty
// subsets S of the members: exactly the members of S attain the minimum) and fold the patterns into
// cnt[t][|S|] at the end.  A pattern flag is one LOP3 of the per-member flags (bit 7 of a byte = member t
// attains the minimum), so no tie size, binary digits or selectors are formed: N = 3 needs 17 ALU-pipe
// instructions per word where the digit form needs about 36.
template <int N>
__device__ __forceinline__ void count_set_small(const uint8_t* const (&rows)[N], int r0, int r1, int n_reads,
                                                uint32_t* __restrict__ out) {
    static_assert(N == 3 || N == 4, "pattern counting is written for three and four members");
    constexpr int NP = (1 << N) - 1;
    constexpr uint32_t H = 0x80808080u;
    const int lane = gk_lane();
    uint32_t pc[NP];                                                 // pc[S - 1], in units of 1/128 read
#pragma unroll
    for (int i = 0; i < NP; ++i) pc[i] = 0u;
    for (int r = r0 + lane * 16; r < r1; r += 32 * 16) {
        uint4 x[N];
#pragma unroll
        for (int t = 0; t < N; ++t) x[t] = __ldg(reinterpret_cast<const uint4*>(rows[t] + r));
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t v[N];
#pragma unroll
            for (int t = 0; t < N; ++t) v[t] = w == 0 ? x[t].x : w == 1 ? x[t].y : w == 2 ? x[t].z : x[t].w;
            uint32_t mn = v[0];
#pragma unroll
            for (int t = 1; t < N; ++t) mn = min7(mn, v[t]);
            uint32_t e[N];                                           // bit 7 of a byte: member t attains the minimum
#pragma unroll
            for (int t = 0; t < N; ++t) {
                e[t] = ge7(mn, v[t]) & H;                            // mn >= v[t] means v[t] attains the minimum
                // keep the masked flag in a register: folded into the pattern expressions the mask becomes
                // a fourth input of every one of them, i.e. a second LOP3 per pattern
                asm volatile("" : "+r"(e[t]));
            }
            if constexpr (N == 3) {
#pragma unroll
                for (int S = 1; S <= NP; ++S) {
                    const uint32_t f = ((S & 1) ? e[0] : ~e[0]) & ((S & 2) ? e[1] : ~e[1]) & ((S & 4) ? e[2] : ~e[2]);
                    pc[S - 1] = __dp4a(f, 0x01010101u, pc[S - 1]);
                }
            } else {
                // the four combinations of members 0 and 1 first (bits outside bit 7 may be set in the
                // negated ones; every pattern has a member, whose flag clears them - except through the
                // pair (not 0, not 1), which is masked here)
                uint32_t p01[4] = {~e[0] & ~e[1] & H, e[0] & ~e[1], ~e[0] & e[1], e[0] & e[1]};
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("" : "+r"(p01[i]));
#pragma unroll
                for (int S = 1; S <= NP; ++S) {
                    const uint32_t f = p01[S & 3] & ((S & 4) ? e[2] : ~e[2]) & ((S & 8) ? e[3] : ~e[3]);
                    pc[S - 1] = __dp4a(f, 0x01010101u, pc[S - 1]);
                }
            }
        }
    }
    const int pad_lo = r0 > n_reads ? r0 : n_reads;
    const uint32_t pad = r1 > pad_lo ? (uint32_t)(r1 - pad_lo) : 0u;
#pragma unroll
    for (int i = 0; i < NP; ++i) pc[i] = warp_sum(pc[i]) >> 7;
    pc[NP - 1] -= pad;                                               // pad reads are all-zero rows: every member ties
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < N; ++t) {
            uint32_t c[N];
#pragma unroll
            for (int q = 0; q < N; ++q) c[q] = 0u;
#pragma unroll
            for (int S = 1; S <= NP; ++S)
                if (S & (1 << t)) c[__popc(S) - 1] += pc[S - 1];
#pragma unroll
            for (int q = 0; q < N; ++q)
                if (c[q]) atomicAdd(out + t * N + q, c[q]);
        }
    }
}

template <int N>
__device__ __forceinline__ void count_set(const uint8_t* const (&rows)[N], int r0, int r1, int n_reads,
                                          uint32_t* __restrict__ out, bool small) {
    if constexpr (N == 2) {
        if (small) count_pair_small(rows[0], rows[1], r0, r1, n_reads, out);
        else count_pair(rows[0], rows[1], r0, r1, n_reads, out);
        return;
    }
    if constexpr (N == 3 || N == 4) {
        if (small) {
            count_set_small<N>(rows, r0, r1, n_reads, out);
            return;
        }
    }
    const int lane = gk_lane();
    uint32_t cnt[N][N];
#pragma unroll
    for (int t = 0; t < N; ++t)
#pragma unroll
        for (int q = 0; q < N; ++q) cnt[t][q] = 0u;

    for (int r = r0 + lane * 16; r < r1; r += 32 * 16) {
        uint4 x[N];
#pragma unroll
        for (int t = 0; t < N; ++t) x[t] = __ldg(reinterpret_cast<const uint4*>(rows[t] + r));
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t v[N];
#pragma unroll
            for (int t = 0; t < N; ++t) v[t] = w == 0 ? x[t].x : w == 1 ? x[t].y : w == 2 ? x[t].z : x[t].w;
            uint32_t mn = v[0];
#pragma unroll
            for (int t = 1; t < N; ++t) {
                const uint32_t keep = (ge80(v[t], mn) >> 7) * 0xffu;        // 0xff where mn stays
                mn = (mn & keep) | (v[t] & ~keep);
            }
            uint32_t eq[N];
            uint32_t q = 0u;
#pragma unroll
            for (int t = 0; t < N; ++t) {
                eq[t] = zero80(v[t] ^ mn) >> 7;
                q += eq[t];
            }
            // binary digits of q (1..N <= 8) per byte
            const uint32_t b0 = q & 0x01010101u, b1 = (q >> 1) & 0x01010101u, b2 = (q >> 2) & 0x01010101u;
#pragma unroll
            for (int qq = 0; qq < N; ++qq) {
                const int c = qq + 1;
                uint32_t sel;
                if (c == 8) {
                    sel = (q >> 3) & 0x01010101u;
                } else {       // q == 8 has the digits 000 and cannot match c in 1..7
                    sel = ((c & 1) ? b0 : ~b0) & ((c & 2) ? b1 : ~b1) & ((c & 4) ? b2 : ~b2) & 0x01010101u;
                }
#pragma unroll
                for (int t = 0; t < N; ++t) cnt[t][qq] = __dp4a(eq[t], sel, cnt[t][qq]);
            }
        }
    }
    // pad reads of this item seen by the whole warp: [max(r0, n_reads), r1)
    const int pad_lo = r0 > n_reads ? r0 : n_reads;
    const uint32_t pad = r1 > pad_lo ? (uint32_t)(r1 - pad_lo) : 0u;
#pragma unroll
    for (int t = 0; t < N; ++t) {
#pragma unroll
        for (int q = 0; q < N; ++q) {
            uint32_t c = warp_sum(cnt[t][q]);
            if (q == N - 1) c -= pad;
            if (lane == 0 && c) atomicAdd(out + t * N + q, c);
        }
    }
}

template <int N>
__global__ void __launch_bounds__(kThreads)
gk_rescore_count_kernel(const GkMatrix* __restrict__ matrices, const GkSearch* __restrict__ searches,
                        const GkCountItem* __restrict__ items, int top_n, const GkStepInfo* __restrict__ info,
                        const int32_t* __restrict__ ids_prev, const int32_t* __restrict__ cand_pool,
                        const int32_t* __restrict__ alive_pool, const uint8_t* __restrict__ LT_pool,
                        uint32_t* __restrict__ cnt_pool) {
    const GkCountItem item = items[blockIdx.x];
    const GkSearch X = searches[item.search];
    const GkMatrix M = matrices[X.matrix];
    const int f = item.f0 + gk_warp();
    const int n_alive = info[item.search].n_alive < X.alive_cap ? info[item.search].n_alive : X.alive_cap;
    if (f >= n_alive) return;
    const int i = alive_pool[X.alive_off + f];
    const int k = i / X.n_cand;
    const int a = cand_pool[X.cand_off + (i - k * X.n_cand)];
    const int32_t* prev = ids_prev + ((int64_t)item.search * top_n + k) * GK_MAX_CN;
    const uint8_t* LT = LT_pool + M.LT_off;
    const uint8_t* rows[N];
#pragma unroll
    for (int t = 0; t < N - 1; ++t) rows[t] = LT + (int64_t)prev[t] * M.r_pad;
    rows[N - 1] = LT + (int64_t)a * M.r_pad;
    count_set<N>(rows, item.r0, item.r1, M.n_reads, cnt_pool + X.cnt_off + (int64_t)f * N * N,
                 M.m_max > 0 && M.m_max < 128);
}

// P tiles: one k-block (GK_KB = 64 kept sets) x 512 reads at a time, over the item's read range.
//   phase 1  one warp per set row: every lane loads 16 reads (128 bits) of each member row of LT and
//            keeps the byte-wise minimum; the row goes to shared memory as 32 quads of 4 words, the
//            quad index XOR-swizzled with (row / 4) so that phase 2 is (almost) conflict free;
//   phase 2  a thread takes 4 sets x 4 reads (four 32-bit words), transposes them with byte permutes
//            and stores 4 sets of one read as 8 bytes: 16 lanes write one 128-byte row of the
//            k-block, and consecutive reads are consecutive rows of the row-blocked P.
constexpr int kPReads = 512;
constexpr int kPWords = kPReads / 4;     // words per set row in shared memory

__global__ void __launch_bounds__(kThreads)
gk_write_p_kernel(const GkMatrix* __restrict__ matrices, const GkSearch* __restrict__ searches,
                  const GkPItem* __restrict__ items, int top_n, int n_set, const int32_t* __restrict__ kept_count,
                  const int32_t* __restrict__ ids, const uint8_t* __restrict__ LT_pool,
                  void* __restrict__ P_pool_raw, int half_mode) {
    __shared__ __align__(16) uint32_t tile[GK_KB * kPWords];
    __shared__ int32_t s_ids[GK_KB * GK_MAX_CN];
    const GkPItem item = items[blockIdx.x];
    const GkSearch X = searches[item.search];
    const GkMatrix M = matrices[X.matrix];
    const int K = kept_count[item.search];
    const int lane = gk_lane();
    const int warp = gk_warp();
    const bool small = M.m_max > 0 && M.m_max < 128;
    // member ids of the 64 sets of this k-block
    for (int i = threadIdx.x; i < GK_KB * GK_MAX_CN; i += kThreads) {
        const int k = item.k_blk * GK_KB + i / GK_MAX_CN;
        s_ids[i] = k < K ? ids[((int64_t)item.search * top_n + k) * GK_MAX_CN + (i % GK_MAX_CN)] : 0;
    }
    __syncthreads();
    const int l16 = threadIdx.x & 15;            // phase 2: sets 4 l16 .. 4 l16 + 3
    const int h16 = threadIdx.x >> 4;            // phase 2: read quads h16, h16 + 16, ...
    for (int r0 = item.r0; r0 < item.r1; r0 += kPReads) {
        const int n_reads = item.r1 - r0 < kPReads ? item.r1 - r0 : kPReads;      // multiple of 128
        const uint8_t* LT = LT_pool + M.LT_off + r0 + lane * 16;
        const bool in_range = lane * 16 < n_reads;
        for (int kl = warp; kl < GK_KB; kl += kWarps) {
            uint4 mn = make_uint4(0u, 0u, 0u, 0u);
            if (in_range && item.k_blk * GK_KB + kl < K) {
                mn = __ldg(reinterpret_cast<const uint4*>(LT + (int64_t)s_ids[kl * GK_MAX_CN] * M.r_pad));
                if (small) {                  // counts below 128: 3 ALU-pipe instructions per byte-wise minimum
                    for (int t = 1; t < n_set; ++t) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(LT + (int64_t)s_ids[kl * GK_MAX_CN + t] * M.r_pad));
                        mn.x = min7(mn.x, v.x);
                        mn.y = min7(mn.y, v.y);
                        mn.z = min7(mn.z, v.z);
                        mn.w = min7(mn.w, v.w);
                    }
                } else {
                    for (int t = 1; t < n_set; ++t) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(LT + (int64_t)s_ids[kl * GK_MAX_CN + t] * M.r_pad));
                        mn.x = __vminu4(mn.x, v.x);
                        mn.y = __vminu4(mn.y, v.y);
                        mn.z = __vminu4(mn.z, v.z);
                        mn.w = __vminu4(mn.w, v.w);
                    }
                }
            }
            *reinterpret_cast<uint4*>(tile + kl * kPWords + 4 * (lane ^ ((kl >> 2) & 15))) = mn;
        }
        __syncthreads();
        for (int rq = h16; rq < n_reads / 4; rq += kThreads / 16) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = tile[(4 * l16 + i) * kPWords + 4 * ((rq >> 2) ^ l16) + (rq & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t o = X.P_off + gk_blk_off(r0 + 4 * rq + j, item.k_blk, X.n_kblk, GK_KB) + 4 * l16;
                const uint32_t sel = 0x0400u + 0x0101u * j;      // byte j of both inputs -> bytes 0 and 2
                if (half_mode) {
                    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(P_pool_raw) + o) =
                        make_uint2(__byte_perm(w[0], w[1], sel) & 0x00ff00ffu, __byte_perm(w[2], w[3], sel) & 0x00ff00ffu);
                } else {
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(P_pool_raw) + o) =
                        make_float4((float)((w[0] >> (8 * j)) & 0xffu), (float)((w[1] >> (8 * j)) & 0xffu),
                                    (float)((w[2] >> (8 * j)) & 0xffu), (float)((w[3] >> (8 * j)) & 0xffu));
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int gk_rescore_count(const GkMatrix* matrices, const GkSearch* searches, const GkCountItem* items,
                                int n_items, int top_n, int n_set, const GkStepInfo* info,
                                const int32_t* ids_prev, const int32_t* cand_pool, const int32_t* alive_pool,
                                const uint8_t* LT_pool, uint32_t* cnt_pool, void* stream) {
    if (n_items <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define GK_COUNT_CASE(N)                                                                                   \
    case N:                                                                                                \
        gk_rescore_count_kernel<N><<<n_items, kThreads, 0, st>>>(matrices, searches, items, top_n, info,   \
                                                                 ids_prev, cand_pool, alive_pool, LT_pool, \
                                                                 cnt_pool);                                \
        break;
    switch (n_set) {
        GK_COUNT_CASE(2)
        GK_COUNT_CASE(3)
        GK_COUNT_CASE(4)
        GK_COUNT_CASE(5)
        GK_COUNT_CASE(6)
        GK_COUNT_CASE(7)
        GK_COUNT_CASE(8)
        default:
            GK_REQUIRE(false, "gk_rescore_count: set size %d outside 2..%d", n_set, GK_MAX_CN);
    }
#undef GK_COUNT_CASE
    GK_CHECK_LAUNCH("gk_rescore_count");
    return 0;
}

extern "C" int gk_write_p(const GkMatrix* matrices, const GkSearch* searches, const GkPItem* items, int n_items,
                          int top_n, int n_set, const int32_t* kept_count, const int32_t* ids,
                          const uint8_t* LT_pool, void* P_pool, int half_mode, void* stream) {
    if (n_items <= 0) return 0;
    GK_REQUIRE(n_set >= 1 && n_set <= GK_MAX_CN, "gk_write_p: set size %d outside 1..%d", n_set, GK_MAX_CN);
    gk_write_p_kernel<<<n_items, kThreads, 0, (cudaStream_t)stream>>>(matrices, searches, items, top_n, n_set,
                                                                      kept_count, ids, LT_pool, P_pool, half_mode);
    GK_CHECK_LAUNCH("gk_write_p");
    return 0;
}
