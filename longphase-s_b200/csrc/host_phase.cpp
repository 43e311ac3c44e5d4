// host_phase.cpp — host-side steps of the `phase` pipeline that sit BETWEEN the kernels.
//
// They are inherently sequential or tiny (O(#alignments sharing a name), O(#clip positions),
// O(#variants * window)) and stay on the host by design, but they
// must be bit-exact because kernel 2 consumes the filters and kernel 3a consumes the sweep:
//   * overlap filter among the alignments of one read name   reference PhasingGraph.cpp:707-781
//   * Clip::getCNVInterval (run twice) + CNV mismatch filter  reference PhasingGraph.cpp:520-692, 1103-1227
//   * edgeConnectResult / Onelongcase (the chain itself)        reference PhasingGraph.cpp:251-283, 286-474
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include "lps_ctx.cuh"

// --------------------------------------------------------------------------------------------
// overlap filter.  Only read names that own more than one alignment in the batch can be affected, so the
// batch is indexed once at submit time (lps_host_index_names); the state machine itself runs on the device, one
// thread per such name (k_overlap_filter, k_edges.cu), on (first called position, last called position, #calls)
// of just their alignments.
// --------------------------------------------------------------------------------------------
void lps_host_index_names(lps_ctx *ctx) {
    const std::vector<int32_t> &rank = ctx->h_name_rank;
    const int n = (int)rank.size();
    ctx->h_multi_members.clear(); ctx->h_multi_group_off.assign(1, 0);
    int max_rank = -1;
    for (int r = 0; r < n; r++) if (rank[r] > max_rank) max_rank = rank[r];
    if (max_rank < 0) return;
    std::vector<int32_t> count((size_t)max_rank + 1, 0);
    for (int r = 0; r < n; r++) count[(size_t)rank[r]]++;
    // slot of every multi-alignment name, in order of first appearance; members keep BAM order
    std::vector<int32_t> slot((size_t)max_rank + 1, -1);
    std::vector<std::vector<int32_t>> groups;
    for (int r = 0; r < n; r++) {
        const int k = rank[r];
        if (count[(size_t)k] < 2) continue;
        if (slot[(size_t)k] < 0) { slot[(size_t)k] = (int32_t)groups.size(); groups.emplace_back(); }
        groups[(size_t)slot[(size_t)k]].push_back(r);
    }
    for (const auto &g : groups) {
        ctx->h_multi_members.insert(ctx->h_multi_members.end(), g.begin(), g.end());
        ctx->h_multi_group_off.push_back((int32_t)ctx->h_multi_members.size());
    }
}

// --------------------------------------------------------------------------------------------
// Clip::getCNVInterval state machine over the clipCount map (ascending positions).
// --------------------------------------------------------------------------------------------
namespace {
struct CnvState {
    bool push = false, slow_up = false, slow_down = false;
    int curr = 0, reject = 0, pull_down = 0, slow_down_count = 0, cand_start = -1, cand_end = -1;
    void reset() { *this = CnvState(); }
    void thresholds(int up) {   // Clip::updateThreshold :1112-1126
        reject = up;
        if (up >= 20) { pull_down = up / 2; slow_down_count = 5; }
        else if (up >= 10) { pull_down = up / 2; slow_down_count = up / 4; }
        else { pull_down = 5; slow_down_count = 2; }
    }
    void arm(int pos, int up, int down, int area) {
        push = true; slow_up = false; slow_down = true;
        curr = up - down; cand_start = pos; cand_end = pos + area;
        thresholds(up);
    }
};
}  // namespace

void lps_host_cnv_intervals(const std::vector<int32_t> &pos, const std::vector<int32_t> &front,
                            const std::vector<int32_t> &back, std::vector<int32_t> &cs, std::vector<int32_t> &ce) {
    const int area = 30000;
    const size_t n = pos.size();
    if (n == 0) return;   // the reference crashes here (rbegin() of an empty map); we return no interval
    CnvState s;
    for (size_t k = 0; k <= n; k++) {
        const bool sentinel = k == n;   // copy of the last entry, AreaSize to the right (:1134)
        const int at = sentinel ? pos[n - 1] + area : pos[k];
        const int up = front[sentinel ? n - 1 : k], down = back[sentinel ? n - 1 : k];
        const bool idle = !s.push && !s.slow_down && !s.slow_up;
        if (idle) {
            if (up >= 5 && s.curr == 0) s.arm(at, up, down, area);
            else if (up > down && s.curr == 0) {
                s.push = false; s.slow_up = true; s.slow_down = false;
                s.curr = up - down; s.cand_start = at; s.cand_end = at + area;
            }
        } else if (s.push && s.slow_down) {
            if (up > s.reject) {
                s.thresholds(up);
                s.cand_start = at; s.cand_end = at + area;
            }
            s.curr += up - down;
            if (s.curr > 30) s.cand_end = at + area;
            if (down >= s.pull_down || (s.curr <= s.slow_down_count && at <= s.cand_end)) {
                cs.push_back(s.cand_start); ce.push_back(at);
                s.reset();
            }
            if (at > s.cand_end || s.curr <= 0 || at - s.cand_start >= 200000) s.reset();
        } else if (s.slow_up) {
            const bool closes = s.curr > 20 ? down >= s.curr / 4 : down >= 5;
            if (closes) { cs.push_back(s.cand_start); ce.push_back(at); s.reset(); }
            else if (up >= 5) s.arm(at, up, down, area);
            else {
                s.curr += up - down;
                if (s.curr > 30) s.cand_end = at + area;
                if (at > s.cand_end || s.curr <= 0 || at - s.cand_start >= 200000) s.reset();
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// CNV mismatch filter (rare: needs clip pile-ups).  Operates on host copies of the calls of the
// surviving alignments; fills `erased` (one byte per call of the batch CSR).
// The interval vector is the state machine's output TWICE (Clip ctor + PhasingProcess.cpp:148), hence
// not sorted; the index walking below follows the reference literally for that reason.
// --------------------------------------------------------------------------------------------
int lps_host_cnv_filter(lps_ctx *ctx, std::vector<uint8_t> &erased) {
    const std::vector<int32_t> &cs = ctx->h_cnv_start, &ce = ctx->h_cnv_end;
    const size_t ncnv = cs.size();
    const int n = (int)ctx->h_status.size();
    erased.assign(ctx->h_calls.size(), 0);
    if (ncnv == 0) return LPS_OK;
    const std::vector<uint64_t> &off = ctx->h_call_off;
    const std::vector<lps_call> &calls = ctx->h_calls;
    const std::vector<int32_t> &vpos = ctx->h_vpos;
    auto inside = [](int p, int s, int e) { return p >= s && p <= e; };
    std::vector<int> alns;
    for (int r = 0; r < n; r++) if (off[(size_t)r + 1] > off[(size_t)r] && !ctx->h_read_dead[(size_t)r]) alns.push_back(r);

    // calculateCnvMismatchRate: ALT calls of each alignment inside each interval, keyed by interval start
    std::vector<std::map<int, int>> alt_in_cnv(alns.size());
    size_t ci = 0;
    for (size_t a = 0; a < alns.size(); a++) {
        const uint64_t c0 = off[(size_t)alns[a]], c1 = off[(size_t)alns[a] + 1];
        const int rs = vpos[(size_t)calls[c0].var], re = vpos[(size_t)calls[c1 - 1].var];
        while (ci > 0 && cs[ci] > rs) ci--;
        size_t i = ci;
        for (; i < ncnv && cs[i] <= re; i++)
            for (uint64_t c = c0; c < c1; c++) {
                const int vp = vpos[(size_t)calls[c].var];
                if (vp > ce[i]) break;
                if (inside(vp, cs[i], ce[i]) && calls[c].allele == 1) alt_in_cnv[a][cs[i]]++;
            }
        ci = i > 0 ? i - 1 : 0;
    }
    // aggregateCnvReadMismatchRate: per (variant, allele) the list of those per-read counts
    std::map<int, std::map<int, std::vector<int>>> per_variant;   // variant index -> allele -> counts
    ci = 0;
    for (size_t a = 0; a < alns.size(); a++) {
        const uint64_t c0 = off[(size_t)alns[a]], c1 = off[(size_t)alns[a] + 1];
        const int rs = vpos[(size_t)calls[c0].var], re = vpos[(size_t)calls[c1 - 1].var];
        while (ci > 0 && cs[ci] > rs) ci--;
        size_t i = ci;
        for (; i < ncnv && cs[i] <= re; i++)
            for (uint64_t c = c0; c < c1; c++) {
                const int vp = vpos[(size_t)calls[c].var];
                if (vp > ce[i]) break;
                auto hit = alt_in_cnv[a].find(cs[i]);
                if (inside(vp, cs[i], ce[i]) && hit != alt_in_cnv[a].end()) per_variant[calls[c].var][calls[c].allele].push_back(hit->second);
            }
        ci = i > 0 ? i - 1 : 0;
    }
    // calculateAverageMismatchRate
    auto mean = [](const std::vector<int> &v) { double s = 0.0; for (int x : v) s += x; return v.empty() ? 0.0 : s / (double)v.size(); };
    std::map<int, double> miss;
    for (const auto &kv : per_variant) {
        const int vp = vpos[(size_t)kv.first];
        for (size_t i = 0; i < ncnv; i++) {   // the reference never advances its start index in this function
            if (cs[i] > vp) break;
            if (!inside(vp, cs[i], ce[i])) continue;
            auto r = kv.second.find(0), al = kv.second.find(1);
            if (r == kv.second.end() || al == kv.second.end()) continue;
            const double mr = mean(r->second), ma = mean(al->second);
            if (mr != 0 && ma != 0) miss[kv.first] = ma / (mr + ma);
        }
    }
    if (miss.empty()) return LPS_OK;
    // filterHighMismatchVariants
    ci = 0;
    for (size_t a = 0; a < alns.size(); a++) {
        const uint64_t c0 = off[(size_t)alns[a]], c1 = off[(size_t)alns[a] + 1];
        const int rs = vpos[(size_t)calls[c0].var];
        while (ci > 0 && cs[ci] > rs) ci--;
        for (uint64_t c = c0; c < c1; c++) {
            const int vp = vpos[(size_t)calls[c].var];
            size_t i = ci;
            for (; i < ncnv && cs[i] <= vp; i++) {
                if (!inside(vp, cs[i], ce[i])) continue;
                auto m = miss.find(calls[c].var);
                if (m != miss.end() && m->second >= 0.7) { erased[c] = 1; break; }
            }
            ci = i > 0 ? i - 1 : 0;
        }
    }
    return LPS_OK;
}

// --------------------------------------------------------------------------------------------
// the sweep (edgeConnectResult, PhasingGraph.cpp:286-474).  It is a left-to-right chain — node k's haplotype comes from the
// weighted votes of its <= W predecessors, then k votes on its W successors — so it runs on the host; but nothing about a vote
// except its DIRECTION depends on the chain, and k_fold_edges' epilogue already reduced every (voter, successor) cell to one
// byte (findBestEdgePair, :166-228):
//   bits 0-1 link (1 same haplotype, 2 opposite, 0 none), bit 2 weight-20 rule, bit 3 (para+cross) <= 1,
//   bit 4 edgeSimilarRatio < 0.2.
// Only those bytes cross PCIe, never the [nodes][W][4] float table.
//   * hpCountMap2 is a float sum in voter order (weights 1, 20, 0.1f): successors receive their votes one voter at a time in
//     ascending voter order, exactly like the reference;
//   * Onelongcase's sums only ever add 1 or 20 (at most 35 x 20) and the count of single-read edges is at most W, so the three
//     live in ONE 32-bit accumulator per node: bits 0-7 singles, 8-19 sum towards haplotype 1, 20-31 towards haplotype 2;
//   * inside a block the REF-allele haplotype telescopes to hp[k]-1 (the first member of a block always has hp 1) and
//     PS = position(block start)+1; single-member blocks are dropped (:425).
// Layout.  The accumulators are cut into ALIGNED blocks of 16 nodes, and the device writes row k of the vote bytes already
// shifted to those blocks: lps_vote_row_stride(W) bytes per row, byte j = vote on node 16*((k+1)/16) + j, zero where there is no
// vote.  A voter therefore adds whole aligned vectors, with no edge masks, and every load of an accumulator block has exactly
// the address and width of the store the previous voter made to it, so store-to-load forwarding always works (the first version
// added at unaligned offsets t0+d: every load straddled two earlier stores and waited for both to retire — 47 ns per node).
// --------------------------------------------------------------------------------------------
int lps_vote_row_stride(int W) { return ((W + 15 + 15) / 16) * 16; }

namespace {

struct SweepAcc {
    std::vector<char> mem;
    float *w1, *w2;
    uint32_t *pk;
    explicit SweepAcc(size_t len) {
        // staggered by 13 cache lines so that w1[i], w2[i], pk[i] never share their low 12 address bits (4K aliasing)
        const size_t bytes = (len * 4 + 4095) & ~(size_t)4095;
        mem.assign(3 * bytes + 3 * 832 + 64, 0);
        char *b = mem.data();
        b += (64 - ((uintptr_t)b & 63)) & 63;
        w1 = (float *)(b);
        w2 = (float *)(b + 1 * (bytes + 832));
        pk = (uint32_t *)(b + 2 * (bytes + 832));
    }
};

// Sums of the node that follows the voter, handed from the voter's registers to the next step of the chain: a scalar load of
// them would have to wait for the voter's vector stores to leave the store buffer.
struct NextNode { float h1, h2; uint32_t pk; int lane; bool valid; };

inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t as_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// Voter k (haplotype hp, node type `type`) votes on the nodes of its row; `base` = first node of the row = 16 * ((k+1) / 16).
// Mask arithmetic instead of branches: the vote directions are data, not control flow (adding +0.0f is exact and the
// accumulators are never -0.0).
inline void cast_votes_scalar(const uint8_t *row, int RS, int hp, unsigned type, const SweepAcc &A, size_t base, NextNode &nx) {
    const uint32_t wtab[2] = {as_bits(type == 4u ? (float)0.1 : 1.f), as_bits(type == 4u ? (float)0.1 : 20.f)};   // :367-369
    const uint32_t itab[2] = {1u, 20u};
    const unsigned same = hp == 1 ? 1u : 2u, other = 3u - same;       // link code that sends the vote to haplotype 1
    const uint32_t type_ok = (type != 3u && type != 4u) ? ~0u : 0u;  // Onelongcase: weight >= 1, voter not an indel (:265)
    for (int j = 0; j < RS; j++) {
        const unsigned info = row[j], link = info & 3u;
        if (!link) continue;
        const uint32_t mA = 0u - (uint32_t)(link == same), mB = 0u - (uint32_t)(link == other);
        const uint32_t wb = wtab[(info >> 2) & 1u];
        A.w1[base + j] += as_float(wb & mA);
        A.w2[base + j] += as_float(wb & mB);
        const uint32_t single = 0u - ((info >> 3) & 1u);
        const uint32_t qual = ~single & (0u - ((info >> 4) & 1u)) & type_ok;
        const uint32_t wi = itab[(info >> 2) & 1u];
        A.pk[base + j] += ((mA | mB) & single & 1u) | ((wi & qual & mA) << 8) | ((wi & qual & mB) << 20);
    }
    nx.h1 = A.w1[base + nx.lane]; nx.h2 = A.w2[base + nx.lane]; nx.pk = A.pk[base + nx.lane];
}

#if defined(__x86_64__)
#define LPS_T_AVX2 __attribute__((target("avx2")))
#define LPS_T_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq")))

struct Avx2Consts {
    __m256i c1, c3, c4, c8, c16, c20;
};

LPS_T_AVX2 inline __attribute__((always_inline)) void cast_votes_avx2(const uint8_t *row, int first, int last, int hp, unsigned type,
                                                                    const SweepAcc &A, size_t base, const Avx2Consts &C, NextNode &nx) {
    const __m256 w_lo = _mm256_set1_ps(type == 4u ? (float)0.1 : 1.f), w_hi = _mm256_set1_ps(type == 4u ? (float)0.1 : 20.f);
    const __m256i same = _mm256_set1_epi32(hp), other = _mm256_set1_epi32(3 - hp);       // link code that sends the vote to haplotype 1 / 2
    const __m256i type_ok = _mm256_set1_epi32((type != 3u && type != 4u) ? -1 : 0);
    for (int b = first; b <= last; b++) {                     // blocks of 8 nodes; the ones before `first` / after `last` hold no vote
        const __m256i info = _mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i *)(row + 8 * b)));
        const __m256i link = _mm256_and_si256(info, C.c3);
        const __m256i mA = _mm256_cmpeq_epi32(link, same), mB = _mm256_cmpeq_epi32(link, other);
        const __m256i heavy = _mm256_cmpeq_epi32(_mm256_and_si256(info, C.c4), C.c4);
        const __m256 wb = _mm256_blendv_ps(w_lo, w_hi, _mm256_castsi256_ps(heavy));
        float *p1 = A.w1 + base + 8 * b, *p2 = A.w2 + base + 8 * b;
        const __m256 v1 = _mm256_add_ps(_mm256_load_ps(p1), _mm256_and_ps(wb, _mm256_castsi256_ps(mA)));
        const __m256 v2 = _mm256_add_ps(_mm256_load_ps(p2), _mm256_and_ps(wb, _mm256_castsi256_ps(mB)));
        _mm256_store_ps(p1, v1);
        _mm256_store_ps(p2, v2);
        const __m256i act = _mm256_or_si256(mA, mB);
        const __m256i single = _mm256_cmpeq_epi32(_mm256_and_si256(info, C.c8), C.c8);
        const __m256i qual = _mm256_and_si256(_mm256_andnot_si256(single, _mm256_cmpeq_epi32(_mm256_and_si256(info, C.c16), C.c16)), type_ok);
        const __m256i wi = _mm256_and_si256(_mm256_blendv_epi8(C.c1, C.c20, heavy), qual);
        const __m256i add = _mm256_or_si256(_mm256_and_si256(_mm256_and_si256(act, single), C.c1),
                                            _mm256_or_si256(_mm256_slli_epi32(_mm256_and_si256(wi, mA), 8),
                                                            _mm256_slli_epi32(_mm256_and_si256(wi, mB), 20)));
        __m256i *pp = (__m256i *)(A.pk + base + 8 * b);
        const __m256i vp = _mm256_add_epi32(_mm256_load_si256(pp), add);
        _mm256_store_si256(pp, vp);
        if (b == first) {                                     // the next node sits in the first block: hand its sums over in registers
            const __m256i lane = _mm256_set1_epi32(nx.lane & 7);
            nx.h1 = _mm256_cvtss_f32(_mm256_permutevar8x32_ps(v1, lane));
            nx.h2 = _mm256_cvtss_f32(_mm256_permutevar8x32_ps(v2, lane));
            nx.pk = (uint32_t)_mm256_cvtsi256_si32(_mm256_permutevar8x32_epi32(vp, lane));
        }
    }
}

struct Avx512Consts {
    __m512i c1, c2, c3, c4, c8, c16, a1_lo, a1_hi, a2_lo, a2_hi;
};

LPS_T_AVX512 inline __attribute__((always_inline)) void cast_votes_avx512(const uint8_t *row, int nblk, int hp, unsigned type,
                                                                        const SweepAcc &A, size_t base, const Avx512Consts &C, NextNode &nx) {
    const __m512 w_lo = _mm512_set1_ps(type == 4u ? (float)0.1 : 1.f), w_hi = _mm512_set1_ps(type == 4u ? (float)0.1 : 20.f);
    const __m512i same = _mm512_set1_epi32(hp), other = _mm512_set1_epi32(3 - hp);       // link code that sends the vote to haplotype 1 / 2
    const __mmask16 tok = (type != 3u && type != 4u) ? (__mmask16)0xFFFF : (__mmask16)0;
    for (int b = 0; b < nblk; b++) {                          // blocks of 16 nodes
        const __m512i info = _mm512_cvtepu8_epi32(_mm_load_si128((const __m128i *)(row + 16 * b)));
        const __m512i link = _mm512_and_si512(info, C.c3);
        const __mmask16 mA = _mm512_cmpeq_epi32_mask(link, same), mB = _mm512_cmpeq_epi32_mask(link, other);
        const __mmask16 heavy = _mm512_test_epi32_mask(info, C.c4), single = _mm512_test_epi32_mask(info, C.c8);
        const __mmask16 qual = _kandn_mask16(single, _mm512_mask_test_epi32_mask(tok, info, C.c16));
        const __m512 wb = _mm512_mask_blend_ps(heavy, w_lo, w_hi);
        float *p1 = A.w1 + base + 16 * b, *p2 = A.w2 + base + 16 * b;
        __m512 v1 = _mm512_load_ps(p1), v2 = _mm512_load_ps(p2);
        v1 = _mm512_mask_add_ps(v1, mA, v1, wb);
        v2 = _mm512_mask_add_ps(v2, mB, v2, wb);
        _mm512_store_ps(p1, v1);
        _mm512_store_ps(p2, v2);
        __m512i *pp = (__m512i *)(A.pk + base + 16 * b);
        __m512i pk = _mm512_load_si512(pp);
        pk = _mm512_mask_add_epi32(pk, _kand_mask16(_kor_mask16(mA, mB), single), pk, C.c1);
        pk = _mm512_mask_add_epi32(pk, _kand_mask16(mA, qual), pk, _mm512_mask_blend_epi32(heavy, C.a1_lo, C.a1_hi));
        pk = _mm512_mask_add_epi32(pk, _kand_mask16(mB, qual), pk, _mm512_mask_blend_epi32(heavy, C.a2_lo, C.a2_hi));
        _mm512_store_si512(pp, pk);
        if (b == 0) {                                         // the next node sits in the first block: hand its sums over in registers
            const __mmask16 one = (__mmask16)(1u << nx.lane);
            nx.h1 = _mm512_cvtss_f32(_mm512_maskz_compress_ps(one, v1));
            nx.h2 = _mm512_cvtss_f32(_mm512_maskz_compress_ps(one, v2));
            nx.pk = (uint32_t)_mm512_cvtsi512_si32(_mm512_maskz_compress_epi32(one, pk));
        }
    }
}
#endif

// the chain itself; CAST(row, hp, type, base, k) adds voter k's votes
#define LPS_SWEEP_CHAIN(CAST)                                                                                                     \
    int block_start = -1, block_size = 0, block_ps = 0, last_connect = -1;                                                        \
    NextNode nx = {0.f, 0.f, 0u, 0, false};                                                                                       \
    for (int k = 0; k + 1 < N; k++) {                                                                                             \
        const bool handed = nx.valid;                                                                                             \
        nx.valid = false;                                                                                                         \
        if (std::abs(node_pos[k + 1] - node_pos[k]) > p->distance) continue;                   /* :318-320 */                     \
        float h1 = handed ? nx.h1 : A.w1[k], h2 = handed ? nx.h2 : A.w2[k];                                                       \
        const uint32_t pk = handed ? nx.pk : A.pk[k];                                                                             \
        const int sg = (int)(pk & 0xFFu), a1 = (int)((pk >> 8) & 0xFFFu), a2 = (int)(pk >> 20);                                   \
        const bool onelong = (sg > 3) & ((a1 | a2) != 0);                                      /* Onelongcase :276-281 */          \
        const float f1 = (float)a1, f2 = (float)a2;                                                                               \
        h1 = onelong ? f1 : h1;                                                                                                   \
        h2 = onelong ? f2 : h2;                                                                                                   \
        int hp;                                                                                                                   \
        if (h1 == h2) {                                                                                                           \
            if (last_connect >= 0 && k < last_connect) continue;                                /* :340-342 */                    \
            if (block_start >= 0 && block_size == 1) { node_ps[block_start] = 0; node_hap_ref[block_start] = -1; }                \
            block_start = k; block_size = 0; block_ps = node_pos[k] + 1; hp = 1;                                                  \
        } else hp = 2 - (int)(h1 > h2);                                                                                           \
        block_size++;                                                                                                             \
        node_ps[k] = block_ps;                                                                                                    \
        node_hap_ref[k] = (int8_t)(hp - 1);                                                                                       \
        const int L = last_link[k];                                                                                               \
        if (L < 0) continue;                                                                                                      \
        const uint8_t *row = votes + (size_t)k * (size_t)RS;                                                                      \
        const size_t base = ((size_t)k + 1) & ~(size_t)15;                                                                        \
        nx.lane = (k + 1) & 15;                                                                                                   \
        CAST;                                                                                                                     \
        nx.valid = true;                                                                                                          \
        last_connect = k + 1 + L;                                                                                                 \
    }                                                                                                                             \
    if (block_start >= 0 && block_size == 1) { node_ps[block_start] = 0; node_hap_ref[block_start] = -1; }

void sweep_scalar(const lps_phase_params *p, int N, int RS, const int32_t *node_pos, const uint8_t *node_type, const uint8_t *votes,
                  const int8_t *last_link, int32_t *node_ps, int8_t *node_hap_ref, const SweepAcc &A) {
    LPS_SWEEP_CHAIN(cast_votes_scalar(row, RS, hp, node_type[k], A, base, nx))
}

#if defined(__x86_64__)
LPS_T_AVX2 void sweep_avx2(const lps_phase_params *p, int N, int RS, const int32_t *node_pos, const uint8_t *node_type,
                           const uint8_t *votes, const int8_t *last_link, int32_t *node_ps, int8_t *node_hap_ref, const SweepAcc &A) {
    const Avx2Consts C = {_mm256_set1_epi32(1), _mm256_set1_epi32(3), _mm256_set1_epi32(4), _mm256_set1_epi32(8), _mm256_set1_epi32(16),
                          _mm256_set1_epi32(20)};
    LPS_SWEEP_CHAIN(cast_votes_avx2(row, (int)((k + 1) & 15) >> 3, ((int)((k + 1) & 15) + L) >> 3, hp, node_type[k], A, base, C, nx))
}

LPS_T_AVX512 void sweep_avx512(const lps_phase_params *p, int N, int RS, const int32_t *node_pos, const uint8_t *node_type,
                               const uint8_t *votes, const int8_t *last_link, int32_t *node_ps, int8_t *node_hap_ref, const SweepAcc &A) {
    const Avx512Consts C = {_mm512_set1_epi32(1), _mm512_set1_epi32(2), _mm512_set1_epi32(3), _mm512_set1_epi32(4), _mm512_set1_epi32(8),
                            _mm512_set1_epi32(16), _mm512_set1_epi32(1 << 8), _mm512_set1_epi32(20 << 8), _mm512_set1_epi32(1 << 20),
                            _mm512_set1_epi32(20 << 20)};
    const int nblk = RS >> 4;
    LPS_SWEEP_CHAIN(cast_votes_avx512(row, nblk, hp, node_type[k], A, base, C, nx))
}
#endif

}  // namespace

// votes: [N] rows of lps_vote_row_stride(W) bytes, 16-byte aligned, in the block-shifted layout described above;
// last_link: [N] the largest successor offset d on which node k has a link, -1 if none.
// Returns the code path taken (0 scalar, 1 AVX2, 2 AVX-512); LPS_SWEEP=scalar|avx2|avx512 forces one (tests).
int lps_host_sweep(const lps_phase_params *p, int32_t N, int32_t W, const int32_t *node_pos, const uint8_t *node_type,
                   const uint8_t *votes, const int8_t *last_link, int32_t *node_ps, int8_t *node_hap_ref) {
    for (int k = 0; k < N; k++) { node_ps[k] = 0; node_hap_ref[k] = -1; }
    int simd = 0;
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx2")) simd = 1;
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
        __builtin_cpu_supports("avx512dq"))
        simd = 2;
    if (const char *env = getenv("LPS_SWEEP")) {
        const int want = !strcmp(env, "scalar") ? 0 : !strcmp(env, "avx2") ? 1 : !strcmp(env, "avx512") ? 2 : simd;
        if (want <= simd) simd = want;
    }
    if (((uintptr_t)votes & 15) != 0) simd = 0;
#endif
    if (N < 2) return simd;
    const int RS = lps_vote_row_stride(W);
    SweepAcc A((size_t)N + (size_t)RS + 32);
#if defined(__x86_64__)
    if (simd == 2) sweep_avx512(p, N, RS, node_pos, node_type, votes, last_link, node_ps, node_hap_ref, A);
    else if (simd == 1) sweep_avx2(p, N, RS, node_pos, node_type, votes, last_link, node_ps, node_hap_ref, A);
    else
#endif
        sweep_scalar(p, N, RS, node_pos, node_type, votes, last_link, node_ps, node_hap_ref, A);
    return simd;
}


// ---- postProcess of the two extract passes (reference src/somatic_haplotag/SomaticVarCaller.cpp:176-210, 520-603) ----
namespace {
inline float vaf_of(int alt, int depth) { return (depth == 0 || alt == 0) ? 0.0f : (float)alt / (float)depth; }      // base_analysis::calculateVAF
inline double imbalance_of(int h1, int h2, int total) {                                                               // calculateHaplotypeImbalanceRatio
    if (h1 > 0 && h2 > 0) return h1 > h2 ? (double)h1 / (double)total : (double)h2 / (double)total;
    if (h1 == 0 && h2 == 0) return 0.0;
    return 1.0;
}
}  // namespace

void lps_host_post_process(int n_tum, const int32_t *tum_var, const uint8_t *t_alt0, const uint16_t *t_ref_len, const uint16_t *t_alt_len,
                           const int32_t *pos_base, const int32_t *read_hp_count, const int32_t *case_count, bool tumor, float *rf,
                           double *rd, int32_t *case_reads) {
    for (int s = 0; s < n_tum; s++) {
        float *f = rf + (size_t)s * LPS_RF_FIELDS;
        double *d = rd + (size_t)s * LPS_RD_FIELDS;
        for (int k = 0; k < LPS_RF_FIELDS; k++) f[k] = 0.0f;
        for (int k = 0; k < LPS_RD_FIELDS; k++) d[k] = 0.0;
        case_reads[s] = 0;
        const int v = tum_var[s], rl = t_ref_len[v], al = t_alt_len[v];
        const bool snp = rl == 1 && al == 1, ins = rl == 1 && al > 1, del = rl > 1 && al == 1;
        if (!snp && !ins && !del) continue;
        const int32_t *pb = pos_base + (size_t)s * LPS_PB_FIELDS, *hp = read_hp_count + (size_t)s * 9;
        // calculateBaseCommonInfo (:13-40): a SNP counts the reads that show the tumor ALT base, an indel the ALT observations
        int alt = pb[LPS_PB_ALT], mpq_alt = pb[LPS_PB_MPQ_ALT];
        if (snp) {
            const int k = t_alt0[v] == 'A' ? 0 : t_alt0[v] == 'C' ? 1 : t_alt0[v] == 'G' ? 2 : t_alt0[v] == 'T' ? 3 : -1;
            alt = k < 0 ? 0 : pb[LPS_PB_A + k];             // getBaseCount throws for other letters; such records do not occur in a VCF
            mpq_alt = k < 0 ? 0 : pb[LPS_PB_MPQ_A + k];
        }
        const int depth = pb[LPS_PB_DEPTH], mpq_depth = pb[LPS_PB_MPQ_DEPTH], dele = pb[LPS_PB_DEL];
        f[LPS_RF_VAF] = vaf_of(alt, depth);
        f[LPS_RF_MPQ_VAF] = vaf_of(mpq_alt, mpq_depth);
        f[LPS_RF_NONDEL_VAF] = vaf_of(alt, depth - dele);
        f[LPS_RF_LOW_MPQ_RATIO] = depth == 0 ? 0.0f : (float)(depth - mpq_depth) / (float)depth;
        f[LPS_RF_DEL_RATIO] = (depth == 0 || dele == 0) ? 0.0f : (float)dele / (float)depth;
        const int h1 = hp[1], h2 = hp[2], germ = h1 + h2;
        d[LPS_RD_GERMLINE_IMBALANCE] = imbalance_of(h1, h2, germ);
        d[LPS_RD_PCT_GERMLINE_HP] = (depth == 0 || germ == 0) ? 0.0 : (double)germ / (double)depth;
        if (tumor) {
            const int32_t *cc = case_count + (size_t)s * LPS_CASE_FIELDS;
            const int clean = cc[LPS_CASE_CLEAN_HP3], mixed = cc[LPS_CASE_MIXED];
            case_reads[s] = clean + mixed;
            if (clean + mixed != 0) {
                const float den = (float)clean + (float)mixed;
                f[LPS_RF_MIXED_RATIO] = (float)mixed / den;
                f[LPS_RF_PURE_H1_1_RATIO] = (float)cc[LPS_CASE_PURE_H1_1] / den;
                f[LPS_RF_PURE_H2_1_RATIO] = (float)cc[LPS_CASE_PURE_H2_1] / den;
                f[LPS_RF_PURE_H3_RATIO] = (float)cc[LPS_CASE_PURE_H3] / den;
            }
            const int b1 = hp[1] + hp[5], b2 = hp[2] + hp[7];                       // H1 + H1_1, H2 + H2_1
            d[LPS_RD_ALLELIC_IMBALANCE] = imbalance_of(b1, b2, b1 + b2);
            d[LPS_RD_SOMATIC_IMBALANCE] = imbalance_of(hp[5], hp[7], hp[5] + hp[7]);
        }
    }
}
