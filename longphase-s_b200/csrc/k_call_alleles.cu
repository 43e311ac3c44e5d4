// k_call_alleles.cu — KERNEL 1: per-read CIGAR walk + allele call at every overlapped SNP / indel, five dialects.
// Replaces BamParser::direct_detect_alleles' read filter, BamParser::get_snp and getClip (reference
// src/phase/ParsingBam.cpp:1243-1301, 1303-1634, 1636-1645), the call-erasing half of SnpParser::filterSNP (:891-911),
// CigarParser::parsingCigar with the hooks of its four parsers (src/haplotag/HaplotagParsingBam.cpp:541-670,
// src/haplotag/HaplotagStrategy.cpp, src/somatic_haplotag/SomaticVarCaller.cpp:123-759,
// src/somatic_haplotag/SomaticHaplotagProcess.cpp:310-579).
//
// Mapping: ONE WARP PER READ, persistent CTAs (one per SM), reads claimed from a global counter.
//   * The CIGAR lives in HBM as a 16-bit stream (len << 4 | op, lengths >= 4095 escaped into a side table): 2 bytes per op.
//   * k_prep_reads turns the SoA batch into one 48-byte descriptor per read that has to be walked (filtered reads never reach
//     the walking kernel) and sorts the descriptors into four work segments, reads with many CIGAR ops first.
//   * Every warp owns two CIGAR buffers in shared memory.  A whole super-chunk of a read (up to 1536 ops, 3 KB) is brought in
//     by ONE cp.async.bulk (1-D TMA) that completes on the warp's own mbarrier; the copy of the NEXT super-chunk - of this
//     read, or the first of the next claimed read - is in flight while the current one is decoded, and the descriptors of
//     the next two reads arrive by cp.async.  No register is spent on data in flight and the only wait on HBM is the
//     mbarrier.
//   * phase 1 (streaming, from shared memory): 16 ops per lane and iteration; two dot-product instructions per pair of ops
//     (IDP.2A of the two 12-bit lengths with the (consumes reference, consumes query) flags of the two op codes, looked up
//     as a pair) give the lane's advance sums; one packed warp scan gives the (reference, query) position at which every
//     group of 8 ops starts; the positions go into a shared-memory index.  No per-op position is ever materialised.
//   * phase 2 (once per super-chunk): every lane takes one pending variant, finds its group by binary search in the index,
//     reads the group's 8 ops back from shared memory (one 128-bit load) and walks them in registers to the covering op.
//     SNP hits are only RECORDED; the scattered SEQ-nibble / QUAL gathers happen afterwards, 32 at a time.
//   * calls are compacted with warp ballots and written contiguously per read into a scratch pool; k_gather_calls rebuilds
//     the CSR in read order after a prefix sum, so the output is deterministic.
//
// Sequential quirks of the reference that are reproduced: cursor == lower_bound of the op start; only the
// FIRST pending variant of a D op is examined (homopolymer >= 3 rule); a variant whose query index is beyond l_qseq drops
// the whole read but keeps the clips seen before it; indel alleles look at the op that follows the M op; clips are FRONT
// iff the CIGAR index is 0.
#include <algorithm>
#include <cub/cub.cuh>
#include "lps_ctx.cuh"
#include "lps_async.cuh"

namespace {

#ifndef LPS_WARPS
#define LPS_WARPS 24
#endif
#ifndef LPS_SC_OPS
#define LPS_SC_OPS 1536
#endif
#ifndef LPS_CAND_CAP
#define LPS_CAND_CAP 192
#endif
#ifndef LPS_PREFETCH_SQ
#define LPS_PREFETCH_SQ 1      // L2 prefetch of the SEQ / QUAL sectors when a SNP candidate is found
#endif
#ifndef LPS_PREFETCH_VREC
#define LPS_PREFETCH_VREC 1    // variant records of the first phase-2 round requested before phase 1
#endif
#ifndef LPS_SPLIT_CHAINS
#define LPS_SPLIT_CHAINS 1
#endif
#ifndef LPS_STATIC_SHARE
#define LPS_STATIC_SHARE 85    // per cent of the work list dealt out statically (item k * warps + warp); the rest is claimed from a global counter
#endif
#ifndef LPS_PAIR_WALK
#define LPS_PAIR_WALK 1        // phase-2 walk over pairs of ops with the pair table (0: op by op)
#endif
constexpr int WARPS_PER_CTA = LPS_WARPS;      // one CTA per SM; its dynamic shared memory (~9.2 KB per warp) pins the L1 / shared split
constexpr int SC = LPS_SC_OPS;                // CIGAR ops per super-chunk (one bulk copy, one phase-2 pass); multiple of 512
constexpr int IT_OPS = 512;                   // ops per phase-1 iteration: 16 per lane
constexpr int NGRP = SC / 8;                  // index entries: one per group of 8 ops
constexpr int SC_BUF = SC + 8;                // + one 16-byte unit: the op that follows the super-chunk's last op
constexpr int CAND_CAP = LPS_CAND_CAP;        // 8-byte candidates buffered per warp in shared memory (half as many 16-byte somatic ones)
constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t PAD16 = 1u;                // zero-length insertion: advances nothing
constexpr uint32_t PAD_WORD = PAD16 | (PAD16 << 16);
static_assert(SC % IT_OPS == 0 && SC >= IT_OPS, "a super-chunk is a whole number of phase-1 iterations");

// candidate encoding: x = kind << 30 | payload
//   kind 0: SNP seen inside an M/=/X op, payload = query index
//   kind 1: SNP seen by the D-op rule,   payload = query index
//   kind 2: resolved indel call,         payload = origin << 2 | allele << 1 | danger
struct Cand { int32_t var; uint32_t x; };

struct K1Args {
    DevBatch b;
    DevVariants v;
    int mapping_quality;
    int have_reference;
    int apply_filter;
    int last_var_pos;
    lps_call *calls_tmp;
    unsigned long long calls_cap;
    uint64_t *tmp_start;
    uint32_t *ncalls;
    uint8_t *status;
    uint32_t *clip_keys;
    uint2 *clip_meta;                 // (read, CIGAR index) of every clip event
    int32_t *abort_of_read;           // CIGAR index at which get_snp dropped the read, INT_MAX otherwise (preset by k_prep_reads)
    const uint4 *work;                // read descriptors (3 x uint4 each), four segments: reads with many CIGAR ops first
    const uint32_t *seg_count;        // [4] descriptors per segment (device memory, written by k_prep_reads)
    uint32_t seg_base[4];             // first slot of each segment
    unsigned long long *dbg_times;    // LPS_DEBUG_K1: per warp {globaltimer at start, at end, reads processed, SM id}
    unsigned long long clip_cap;
    CallCounters *counters;
    // overflow pass
    const uint32_t *overflow_reads;   // null in the main pass; else the work slots of the reads to redo
    const uint64_t *overflow_off;
    Cand *overflow_buf;
    uint32_t *overflow_list_out;      // main pass: work slots of the reads that overflowed
    uint64_t *overflow_need_out;      // main pass: candidates they need
    uint32_t overflow_list_cap;
    // ---- tag dialect (CigarParser::parsingCigar + GermlineHaplotagStrategy) ----
    const int32_t *var_ps;            // per variant: phase set
    int mapq_filter;                  // ParsingBamControl::mappingQualityFilter
    int tag_supplementary;            // ParsingBamConfig::tagSupplementary
    int want_calls;                   // also emit the per-read (variant, haplotype) list
    int count_gathers;                // SEQ/QUAL live in pinned host memory: count the gathered sectors
    double percentage;                // ParsingBamConfig::percentageThreshold
    const int8_t *pq_lut;             // [256][256] PQ by (min, max), built on the host with the host libm
    const uint8_t *hp1_is_alt;        // per variant: HP1 carries ALT (GT 1|0)  (somatic dialects; the others read the packed record)
    int8_t *tag_hp;                   // ReadHP: 0 unTag, 1 H1, 2 H2
    int32_t *tag_ps, *tag_pq, *tag_h1, *tag_h2;
    uint8_t *tag_cat;                 // dispatch category, see LPS_TAG_* in lps.h
    // ---- somatic family (extract-normal, extract-tumor, somatic tagging) ----
    DevSomatic som;
    WdItem *wd_items;                 // tumor pass: window-diff work list
    unsigned long long wd_cap;
    int32_t *tag_h3, *tag_end, *tag_len;
    uint8_t *tag_nps;
    int8_t *tag_hpb;
    float *tag_sim;
};

// ---- packed variant record (k_pack_vrec, k_annotate.cu): x = position, y = ref0 | alt0 << 8 | flags << 16 | homopolymer << 24 ----
constexpr unsigned VF_REF1 = 1u << 16, VF_ALT1 = 1u << 17, VF_DANGER = 1u << 18, VF_FILTERED = 1u << 19, VF_HP1ALT = 1u << 20;

// per-op advance bits, two bits per op code: bit0 = consumes the reference (M D N = X), bit1 = consumes the query (M I S = X)
constexpr uint32_t ADV_LUT = (3u << 0) | (2u << 2) | (1u << 4) | (1u << 6) | (2u << 8) | (3u << 14) | (3u << 16);

struct alignas(16) WarpScratch {
    uint16_t cig[2][SC_BUF];          // destinations of the bulk copies
    int2 grp[NGRP];                   // (reference, query) position at which each group of 8 consecutive ops starts
    Cand cand[CAND_CAP];
    uint4 desc[3][3];                 // ring of read descriptors: the read being walked and the next two
    unsigned long long mbar[2];
};
static_assert((SC_BUF * 2) % 16 == 0 && sizeof(WarpScratch) % 16 == 0, "bulk copy destinations must stay 16-byte aligned");

// true length of an escaped op (12-bit field == 0xFFF): binary search of its global index in the side table
__device__ __noinline__ uint32_t esc_len(const uint64_t *__restrict__ long_at, const uint32_t *__restrict__ long_len, uint32_t n_long, uint64_t gop) {
    uint32_t lo = 0, hi = n_long;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (long_at[mid] < gop) lo = mid + 1; else hi = mid;
    }
    return (lo < n_long && long_at[lo] == gop) ? long_len[lo] : 0xFFFu;
}
__device__ __forceinline__ uint32_t op_len(const DevBatch &b, uint32_t x, uint64_t gop) {
    const uint32_t len = x >> 4;
    return len == 0xFFFu ? esc_len(b.long_at, b.long_len, b.n_long, gop) : len;
}

// HaplotagVariantType (HaplotagType.h:76-85): 1 SNP, 2 INSERTION, 3 DELETION, 4 MNP, 0 = setVariantType would throw
__device__ __forceinline__ int hvt(int rl, int al) {
    if (rl == 1) return al == 1 ? 1 : (al > 1 ? 2 : 0);
    if (rl > 1) return al == 1 ? 3 : (al == rl ? 4 : 0);
    return 0;
}

// CigarParser::countBaseNucleotide (HaplotagParsingBam.cpp:682-720)
__device__ __forceinline__ void count_base(int32_t *pb, char base, bool mpq_ok, bool is_alt, int ttype) {
    const int k = base == 'A' ? 0 : base == 'C' ? 1 : base == 'G' ? 2 : base == 'T' ? 3 : 4;
    if (mpq_ok) { atomicAdd(pb + LPS_PB_MPQ_A + k, 1); if (is_alt) atomicAdd(pb + LPS_PB_MPQ_ALT, 1); atomicAdd(pb + LPS_PB_MPQ_DEPTH, 1); }
    atomicAdd(pb + LPS_PB_A + k, 1);
    if (is_alt) { if (ttype == 3) atomicAdd(pb + LPS_PB_DEL, 1); atomicAdd(pb + LPS_PB_ALT, 1); }
    atomicAdd(pb + LPS_PB_DEPTH, 1);
}

// The hooks of the three somatic parsers over the raw candidates of one read, the per-read decision, and the per-position
// counters.  Pass 1 (32 candidates at a time): hooks + votes; warp reduction; decision (all lanes, uniform); pass 2: counters
// that depend on the read's haplotype, and the per-read variant list.
//   extract-normal  ExtractNorDataCigarParser + ExtractNorDataChrProcessor::processRead   SomaticVarCaller.cpp:123-293
//   extract-tumor   ExtractTumDataCigarParser + ExtractTumDataChrProcessor::processRead   SomaticVarCaller.cpp:334-518, 712-759
//   somatic tagging SomaticHaplotagCigarParser + SomaticHaplotagChrProcessor::judgeHaplotype  SomaticHaplotagProcess.cpp:310-579
struct PoolCursor;
__device__ __forceinline__ unsigned long long pool_alloc(const K1Args &a, PoolCursor &pc, int n, int lane);

template <int MODE>
__device__ __forceinline__ void resolve_somatic(const K1Args &a, int r, int lane, uint4 *cand4, int ncand, int ref_start, int ref_end,
                                                int q_end, int lq, PoolCursor &pc, const uint8_t *__restrict__ seq, const bool mpq_ok) {
    constexpr bool XNOR = MODE == LPS_MODE_EXTRACT_NORMAL, XTUM = MODE == LPS_MODE_EXTRACT_TUMOR, STAG = MODE == LPS_MODE_SOMATIC_TAG;
    const DevSomatic &s = a.som;
    int h1 = 0, h2 = 0, h3 = 0, ps_min = INT_MAX, ps_max = INT_MIN, d1 = 0, d2 = 0;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
        const int c = c0 + lane;
        int var = -1, vhp = 0; unsigned flags = 0; bool ps_counted = false;
        if (c < ncand) {
            const uint4 cd = cand4[c];
            var = (int)cd.x;
            const int qidx = (int)cd.y;
            const unsigned fl = cd.w >> 28;
            const bool N = s.nor_present ? s.nor_present[var] != 0 : true, T = s.tum_present[var] != 0;
            const bool nphased = N && (s.nor_gt ? s.nor_gt[var] == 1 : true);          // NORMAL record is PHASED_HETERO
            const int slot = s.slot_of_var[var];
            const int nrl = a.v.ref_len[var], nal = a.v.alt_len[var];
            const int trl = T ? s.t_ref_len[var] : 0, tal = T ? s.t_alt_len[var] : 0;
            const int nty = N ? hvt(nrl, nal) : 0, tty = T ? hvt(trl, tal) : 0;
            const char nrb = (char)a.v.ref0[var], nab = (char)a.v.alt0[var];
            const char trb = T ? (char)s.t_ref0[var] : 0, tab = T ? (char)s.t_alt0[var] : 0;
            const bool h1alt = N && a.hp1_is_alt[var] != 0;
            if (!(fl & 8u)) {
                // ---- processMatchOperation ----
                const char base = "=ACMGRSVTWYHKDBN"[(seq[qidx >> 1] >> ((~qidx & 1) << 2)) & 0xfu];
                // IsAltIndel (HaplotagParsingBam.cpp:650-670) on the NORMAL record when there is one, else on the TUMOR record
                const int sty = N ? nty : tty;
                const bool is_alt = sty == 1 ? base == (N ? nab : tab) : sty == 2 ? (fl & 2u) != 0 : sty == 3 ? (fl & 4u) != 0 : false;
                if (XNOR) {
                    if (T && tty >= 1 && tty <= 3) { flags |= 1u; count_base(s.pos_base + (size_t)slot * LPS_PB_FIELDS, base, mpq_ok, is_alt, tty); }
                    if (mpq_ok && nphased) {
                        // GermlineHaplotagStrategy::judgeSnpHap (HaplotagStrategy.cpp:20-130)
                        if (nty == 1) {
                            if (base == nrb || base == nab) {
                                ps_counted = true;
                                if (base == (h1alt ? nab : nrb)) h1++;
                                if (base == (h1alt ? nrb : nab)) h2++;
                            }
                        } else if ((nty == 2 || nty == 3) && (fl & 1u)) {
                            const bool has = nty == 2 ? (fl & 2u) != 0 : (fl & 4u) != 0;
                            const int l1 = h1alt ? nal : nrl, l2 = h1alt ? nrl : nal;
                            if (l1 != 1 && l2 == 1) { if (has) h1++; else h2++; }
                            else if (l1 == 1 && l2 != 1) { if (has) h2++; else h1++; }
                            ps_counted = true;
                        }
                    }
                } else {
                    if (STAG || mpq_ok) {
                        // SomaticJudgeHapStrategy::judgeSomaticSnpHap (HaplotagStrategy.cpp:315-389)
                        if (N) {
                            if (nphased) {
                                if (nty == 2 || nty == 3) {          // base := whole ALT / REF string, compared with HP1 / HP2
                                    if (is_alt == h1alt) { h1++; vhp = 1; } else { h2++; vhp = 2; }
                                    ps_counted = true;
                                } else if (nty == 1 && (base == nrb || base == nab)) {
                                    if (base == (h1alt ? nab : nrb)) { h1++; vhp = 1; }
                                    if (base == (h1alt ? nrb : nab)) { h2++; vhp = 2; }
                                    ps_counted = true;
                                }
                            }
                        } else if (T) {
                            const int gt = s.t_gt[var];
                            const bool indel = tty == 2 || tty == 3;
                            if (gt >= 1 && gt <= 3 && (indel || (tty == 1 && (base == trb || base == tab)))) {
                                const bool base_is_alt = indel ? is_alt : base == tab;
                                // judgeTumorOnlySnpHap: extract (:617-638) counts every ALT, tagging (:653-668) only somatic variants
                                if (base_is_alt && (XTUM || s.is_somatic[var])) { h3++; vhp = 3; if (XTUM) flags |= 2u; }
                            }
                        }
                        if (XTUM && T) flags |= 1u;               // tumorSnpPosVec
                    }
                    if (XTUM && T && tty >= 1 && tty <= 3) {
                        if (tty != 1 || base == trb || base == tab) {
                            atomicAdd(s.allele_count + (size_t)slot * 2 + (is_alt ? 1 : 0), 1);
                            const unsigned long long k = atomicAdd(&a.counters->wd_items.v, 1ull);
                            if (k < a.wd_cap) {
                                WdItem it;
                                it.read = (uint32_t)r; it.slot2 = (uint32_t)slot * 2u + (is_alt ? 1u : 0u); it.opi = cd.z; it.qidx = cd.y;
                                it.off = cd.w & 0x0fffffffu;
                                a.wd_items[k] = it;
                            }
                        }
                        count_base(s.pos_base + (size_t)slot * LPS_PB_FIELDS, base, mpq_ok, is_alt, tty);
                    }
                    if (STAG && s.is_somatic[var]) flags |= 4u;     // somaticVarDeriveHP entry
                }
            } else if (XNOR || XTUM) {
                // ---- processDeletionOperation ----
                if (T) {
                    int32_t *pb = s.pos_base + (size_t)slot * LPS_PB_FIELDS;
                    if (XNOR) flags |= 1u;
                    if (tty == 1) { atomicAdd(pb + LPS_PB_DEL, 1); atomicAdd(pb + LPS_PB_DEPTH, 1); }
                    else if (tty == 3) { atomicAdd(pb + LPS_PB_ALT, 1); atomicAdd(pb + LPS_PB_DEL, 1); atomicAdd(pb + LPS_PB_DEPTH, 1); }
                }
                // first phased-het NORMAL variant of the D op only (alreadyJudgeDel); judgeDeletionHap (HaplotagStrategy.cpp:147-209)
                if (XNOR && mpq_ok && nphased && s.prev_nor[var] < (int)cd.z && a.have_reference && a.v.hom[var] >= 3) {
                    if (nty == 1) {
                        if (qidx < lq) {
                            const char base = "=ACMGRSVTWYHKDBN"[(seq[qidx >> 1] >> ((~qidx & 1) << 2)) & 0xfu];
                            if (base == (h1alt ? nab : nrb)) h1++;
                            if (base == (h1alt ? nrb : nab)) h2++;
                            ps_counted = true;
                        }
                    } else if (nty == 3) {
                        const int l1 = h1alt ? nal : nrl, l2 = h1alt ? nrl : nal;
                        if (l1 != 1 && l2 == 1) h1++; else if (l1 == 1 && l2 != 1) h2++;
                        ps_counted = true;
                    }
                }
            }
            if (ps_counted) { const int ps = a.var_ps[var]; ps_min = min(ps_min, ps); ps_max = max(ps_max, ps); }
            if (STAG && (flags & 4u) && vhp == 3) { const int d = s.derive_hp[var]; d1 += d == 1; d2 += d == 2; }
        }
        __syncwarp();
        if (c < ncand) cand4[c] = make_uint4((unsigned)var, (unsigned)vhp | (flags << 8), 0u, 0u);
    }
    h1 = (int)__reduce_add_sync(FULL, (unsigned)h1); h2 = (int)__reduce_add_sync(FULL, (unsigned)h2);
    h3 = (int)__reduce_add_sync(FULL, (unsigned)h3);
    ps_min = __reduce_min_sync(FULL, ps_min); ps_max = __reduce_max_sync(FULL, ps_max);
    const bool ps_seen = ps_min != INT_MAX, ps_multi = ps_seen && ps_min != ps_max;
    const int imx = h1 > h2 ? h1 : h2, imn = h1 > h2 ? h2 : h1;
    int hp = 0, hp_before = 0, pq;
    float sim = 0.f;
    // PQ from the germline counts (HaplotagStrategy.cpp:279-288, :589-597); -1: beyond the table, the host fills it in
    if (imx == 0) pq = 0; else if (imn == 0) pq = 40; else pq = imx < 256 ? (int)a.pq_lut[imn * 256 + imx] : -1;
    if (XNOR) {
        // GermlineHaplotagStrategy::judgeReadHap (HaplotagStrategy.cpp:243-300)
        const double mx = (double)imx, mn = (double)imn;
        if (!(mx / (mx + mn) < a.percentage)) { if (h1 > h2) hp = 1; if (h1 < h2) hp = 2; }
        if (ps_multi) hp = 0;
    } else {
        // SomaticJudgeHapStrategy::judgeSomaticReadHap (HaplotagStrategy.cpp:452-602); hpCount[4] stays 0, so the tumor
        // similarity is 1 whenever hpCount[3] != 0, and on a tie the normal maximum is H2
        const int max_n = h1 > h2 ? 1 : 2;
        const double nsim = imx == 0 ? 0.0 : (double)imx / ((double)imx + (double)imn);
        if (h3 != 0) {
            if (1.0 >= a.percentage) hp = nsim >= a.percentage ? (max_n == 1 ? 5 : 7) : 3;   // H1_1 / H2_1 / H3
            pq = 40;
        } else if (imx != 0) {
            if (nsim >= a.percentage) hp = max_n;
        }
        if (ps_multi) hp = 0;
        hp_before = hp;
        if (STAG && hp == 3) {
            // inheritHaplotype (SomaticHaplotagProcess.cpp:461-527), similarity in float
            d1 = (int)__reduce_add_sync(FULL, (unsigned)d1); d2 = (int)__reduce_add_sync(FULL, (unsigned)d2);
            const int mx = d1 > d2 ? d1 : d2, mn = d1 > d2 ? d2 : d1;
            sim = mx == 0 ? 0.0f : (float)mx / ((float)mx + (float)mn);
            if ((double)sim >= a.percentage) hp = d1 > d2 ? 5 : 7;
        }
    }
    if (lane == 0) {
        a.tag_hp[r] = (int8_t)hp; a.tag_pq[r] = pq; a.tag_h1[r] = h1; a.tag_h2[r] = h2; a.tag_h3[r] = h3;
        a.tag_ps[r] = XNOR ? (hp ? ps_min : 0) : (hp ? (ps_seen ? ps_min : -1) : 0);       // PS rule (SomaticHaplotagProcess.cpp:409-430)
        a.tag_nps[r] = (uint8_t)(ps_multi ? 2 : (ps_seen ? 1 : 0));
        a.tag_end[r] = ref_end; a.tag_len[r] = q_end; a.tag_hpb[r] = (int8_t)hp_before; a.tag_sim[r] = sim;
    }
    __syncwarp();
    // ---- pass 2: counters keyed by the read's haplotype; the per-read variant list ----
    const bool record = !ps_multi, clean = h1 == 0 || h2 == 0;
    int nout = 0;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
        const int c = c0 + lane;
        bool emit = false;
        uint4 cd = make_uint4(0u, 0u, 0u, 0u);
        if (c < ncand) {
            cd = cand4[c];
            const int var = (int)cd.x, vhp = (int)(cd.y & 0xffu);
            const unsigned flags = cd.y >> 8;
            const size_t sl = (size_t)max(s.slot_of_var[var], 0);
            if (XNOR) {
                if (flags & 1u) atomicAdd(s.read_hp_count + sl * 9 + hp, 1);
            } else if (XTUM) {
                if (flags & 2u) {
                    // classifyReadsByCase (SomaticVarCaller.cpp:462-518) + somaticReadHpCount (:395-404)
                    int32_t *cc = s.case_count + sl * LPS_CASE_FIELDS;
                    if (!record) atomicAdd(cc + LPS_CASE_UNTAG, 1);
                    else if (clean) {
                        atomicAdd(cc + LPS_CASE_CLEAN_HP3, 1);
                        if (h1 == 0 && h2 == 0) atomicAdd(cc + LPS_CASE_PURE_H3, 1);
                        else if (h1 != 0) atomicAdd(cc + LPS_CASE_PURE_H1_1, 1);
                        else atomicAdd(cc + LPS_CASE_PURE_H2_1, 1);
                    } else atomicAdd(cc + LPS_CASE_MIXED, 1);
                    if (hp == 5 || hp == 7 || hp == 3 || hp == 0) atomicAdd(s.somatic_read_hp_count + sl * 9 + hp, 1);
                }
                if (flags & 1u) atomicAdd(s.read_hp_count + sl * 9 + hp, 1);
                emit = vhp != 0 || (flags & 1u);
            } else {
                if (flags & 4u) {
                    // chrReadHpResult::recordReadHp / recordAlignCoverRegion (HaplotagLogging.cpp:13-72)
                    atomicAdd(s.hp_before_count + sl * 9 + hp_before, 1);
                    if (hp_before != 0 && vhp == 3) atomicAdd(s.h3_before_count + sl * 9 + hp_before, 1);
                    atomicAdd(s.hp_after_count + sl * 9 + hp, 1);
                    if (hp != 0 && vhp == 3) atomicAdd(s.h3_after_count + sl * 9 + hp, 1);
                    if (hp != 0) { atomicMin(s.cover_start + sl, ref_start + 1); atomicMax(s.cover_end + sl, ref_end); }
                }
                emit = vhp != 0;
            }
        }
        if (a.want_calls) {
            const unsigned m = __ballot_sync(FULL, emit);
            __syncwarp();
            if (emit) {
                const int dst = nout + __popc(m & ((1u << lane) - 1u));
                const unsigned q = XTUM ? ((cd.y >> 8) & 3u) : ((cd.y >> 8) & 4u);
                cand4[dst] = make_uint4(cd.x, q | ((cd.y & 0xffu) << 16), 0u, 0u);   // dst <= c
            }
            nout += __popc(m);
            __syncwarp();
        }
    }
    const unsigned long long start = pool_alloc(a, pc, nout, lane);
    if (lane == 0) { a.ncalls[r] = (uint32_t)nout; a.tmp_start[r] = start; a.status[r] = LPS_READ_OK; }
    if (start + nout <= a.calls_cap) {
        for (int c = lane; c < nout; c += 32) {
            const uint4 cd = cand4[c];
            lps_call out;
            out.var = (int32_t)cd.x; out.quality = (int16_t)(cd.y & 0xffffu); out.allele = (int8_t)((cd.y >> 16) & 0xffu); out.origin = 0;
            a.calls_tmp[start + c] = out;
        }
    }
}

// Slots of the scratch call pool are handed out in blocks of POOL_BLOCK per warp (one global atomic per ~50 reads instead of
// one per read on the critical path); the unused tail of a block stays empty, k_gather_calls compacts the pool anyway.
constexpr unsigned POOL_BLOCK = 1024;
struct PoolCursor { unsigned long long next, end, used; };

__device__ __forceinline__ unsigned long long pool_alloc(const K1Args &a, PoolCursor &pc, int n, int lane) {
    if (n == 0) return 0ull;
    if (pc.next + (unsigned)n > pc.end) {
        const unsigned long long want = (unsigned)n > POOL_BLOCK ? (unsigned long long)n : (unsigned long long)POOL_BLOCK;
        unsigned long long s = 0;
        if (lane == 0) s = atomicAdd(&a.counters->tmp_calls.v, want);
        s = __shfl_sync(FULL, s, 0);
        pc.next = s; pc.end = s + want;
    }
    const unsigned long long start = pc.next;
    pc.next += (unsigned)n;
    pc.used += (unsigned)n;
    return start;
}
// ---- read descriptor (written by k_prep_reads, 3 x uint4) ----
struct ReadDesc {
    uint64_t cigar_off, seq_off, qual_off;
    int ref_start, lq, ncig, first_var;
    uint32_t r;                        // read index in the batch; 0xFFFFFFFF = no read (the work list is exhausted)
    uint32_t flags;                    // mapq | flag << 8
};
__device__ __forceinline__ ReadDesc unpack_desc(const uint4 d0, const uint4 d1, const uint4 d2) {
    ReadDesc D;
    D.cigar_off = (uint64_t)d0.x | ((uint64_t)d0.y << 32);
    D.seq_off = (uint64_t)d0.z | ((uint64_t)d0.w << 32);
    D.qual_off = (uint64_t)d1.x | ((uint64_t)d1.y << 32);
    D.ref_start = (int)d1.z; D.lq = (int)d1.w;
    D.ncig = (int)d2.x; D.first_var = (int)d2.y; D.r = d2.z; D.flags = d2.w;
    return D;
}

// work item -> slot of the descriptor array: the four segments are handed out one after the other
__device__ __forceinline__ uint32_t slot_of_item(const K1Args &a, uint32_t t, uint32_t c0, uint32_t c1, uint32_t c2) {
    if (t < c0) return a.seg_base[0] + t;
    if (t < c0 + c1) return a.seg_base[1] + (t - c0);
    if (t < c0 + c1 + c2) return a.seg_base[2] + (t - c0 - c1);
    return a.seg_base[3] + (t - c0 - c1 - c2);
}

// state of a warp's two CIGAR buffers
struct Pipe {
    int buf;                           // buffer the NEXT super-chunk to be consumed arrives in
    uint32_t phase;                    // bit b = parity the next wait on mbar[b] uses
};

// one bulk copy: super-chunk s of the read, into buffer `buf`.  All positions are in the read's ALIGNED op space: index 0 is the
// 16-byte boundary at or before the read's first op (mis = cigar_off & 7 ops before it), so every super-chunk starts on a 16-byte
// boundary of both the global stream and the shared buffer.  The copy carries one extra 16-byte unit (the op after the last one),
// clamped to the padded end of the stream.
__device__ __forceinline__ void issue_sc(const K1Args &a, WarpScratch &S, const ReadDesc &D, int s, int buf, int lane) {
    if (lane == 0) {
        const int mis = (int)(D.cigar_off & 7ull);
        const int tot_a = mis + D.ncig;
        const int cnt = min(SC, tot_a - s * SC);
        const uint64_t start = D.cigar_off - (uint64_t)mis + (uint64_t)s * SC;          // multiple of 8 ops
        uint64_t ops = (uint64_t)((cnt + 7) & ~7) + 8;
        const uint64_t avail = ((a.b.cigar_len + 7ull) & ~7ull) - start;
        if (ops > avail) ops = avail;
        const uint32_t bytes = (uint32_t)ops * 2u;
        const uint32_t bar = smem_u32(&S.mbar[buf]);
        fence_proxy_async();           // orders this warp's earlier generic accesses to the buffer before the async write
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(&S.cig[buf][0]), a.b.cigar16 + start, bytes, bar);
    }
}

// One read, one warp.  `cand` is the warp's shared candidate buffer (or the global list of the overflow pass).
//   next: descriptor of the read this warp walks next (r == 0xFFFFFFFF: none); its first super-chunk is requested while the
//   last super-chunk of this read is being decoded.
template <int MODE>
__device__ __forceinline__ void process_read(const K1Args &a, WarpScratch &S, const uint32_t *__restrict__ s_pair, const ReadDesc &D,
                                             const ReadDesc &next, Pipe &pp, Cand *cand, int cand_cap, const uint32_t work_slot,
                                             const int lane, PoolCursor &pc) {
    constexpr bool TAG = MODE != LPS_MODE_PHASE;          // CigarParser::parsingCigar instead of BamParser::get_snp
    constexpr bool SOM = MODE >= LPS_MODE_EXTRACT_NORMAL; // raw 16-byte candidates, resolved by resolve_somatic
    if (SOM) cand_cap >>= 1;                              // 16-byte candidates
    uint4 *cand4 = reinterpret_cast<uint4 *>(cand);
    const int nv = a.v.n;
    const int r = (int)D.r;
    const int ref_start = D.ref_start, lq = D.lq, ncig = D.ncig;
    int cur = D.first_var;                                // lower_bound(variants, ref_start), from k_prep_reads
    const uint2 *__restrict__ vrec = a.v.vrec;
    const int mis = (int)(D.cigar_off & 7ull);
    const int tot_a = mis + ncig;
    const int n_sc = ncig > 0 ? (tot_a + SC - 1) / SC : 0;
    const uint64_t gop0 = D.cigar_off - (uint64_t)mis;    // global op index of aligned position 0

    const uint8_t *__restrict__ seq = a.b.seq4 + D.seq_off;
    const uint8_t *__restrict__ qual = a.b.qual + D.qual_off;
    // resolve gathers one SEQ nibble (phase: and one QUAL byte) per SNP candidate from anywhere in the read: ask L2 for the sectors
    // as soon as the query index is known, the rest of phase 2 runs while they travel
    auto prefetch_base = [&](int qi) {
#if LPS_PREFETCH_SQ
        if (!TAG && a.b.sq) {          // interleaved rows: base and quality share one 16-byte unit
            asm volatile("prefetch.global.L2 [%0];" ::"l"(seq + (size_t)sq_unit_of((uint32_t)qi) * SQ_UNIT));
            return;
        }
        asm volatile("prefetch.global.L2 [%0];" ::"l"(seq + (qi >> 1)));
#ifndef LPS_DEBUG_NO_QUAL
        if (!TAG) asm volatile("prefetch.global.L2 [%0];" ::"l"(qual + qi));
#endif
#else
        (void)qi;
#endif
    };
    int ref_pos = ref_start, qpos = 0;
    int ncand = 0;
    bool aborted = false;
    int abort_op = INT_MAX, bad_op = INT_MAX;

    if (n_sc == 0 && next.r != 0xFFFFFFFFu && next.ncig > 0) issue_sc(a, S, next, 0, pp.buf, lane);   // keeps "the next read's copy is under way"
    for (int s = 0; s < n_sc; s++) {
        const int buf = pp.buf;
        // the variant records the first phase-2 round of this super-chunk looks at: requested now, used after phase 1
        uint2 vr_first = make_uint2((unsigned)INT_MAX, 0u);
#if LPS_PREFETCH_VREC
        if (cur + lane < nv) vr_first = vrec[cur + lane];
#endif
        // request what this warp decodes next, then wait for what it decodes now
        if (s + 1 < n_sc) issue_sc(a, S, D, s + 1, buf ^ 1, lane);
        else if (next.r != 0xFFFFFFFFu && next.ncig > 0) issue_sc(a, S, next, 0, buf ^ 1, lane);
        mbar_wait(smem_u32(&S.mbar[buf]), (pp.phase >> buf) & 1u);
        pp.phase ^= 1u << buf;
        pp.buf = buf ^ 1;
        uint16_t *__restrict__ cg = S.cig[buf];
        const int a_base = s * SC;                        // aligned position of cg[0]
        const int cnt = min(SC, tot_a - a_base);          // aligned positions of this super-chunk
        const int cnt16 = (cnt + 15) & ~15;
        // positions before the read's first op and behind its last one hold the neighbours' ops: blank them
        if (s == 0 && lane < mis) cg[lane] = (uint16_t)PAD16;
        if (s == n_sc - 1 && lane < 16 && cnt + lane < cnt16) cg[cnt + lane] = (uint16_t)PAD16;
        fence_proxy_async();           // these generic writes precede the next bulk copy into this buffer
        __syncwarp();

        // ================= phase 1: index the group starts of this super-chunk =================
        const int n_it = (cnt + IT_OPS - 1) / IT_OPS;
        for (int it = 0; it < n_it; it++) {
            const int a0 = it * IT_OPS + lane * 16;       // this lane's 16 ops start here (position inside the super-chunk)
            uint32_t w[8];
            if (a0 < cnt16) {
                const uint4 lo = *reinterpret_cast<const uint4 *>(cg + a0), hi = *reinterpret_cast<const uint4 *>(cg + a0 + 8);
                w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w; w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) w[j] = PAD_WORD;
            }
            // advance sums of the lane's two groups of 8 ops: per pair of ops one table word (flags of both op codes) and two
            // dot products of the two 12-bit lengths with them
            unsigned rs = 0, qs = 0, r0 = 0, q0 = 0, acc = 0;
#if LPS_SPLIT_CHAINS
            // the two groups of 8 ops accumulate independently: dependent chains of 4 dot products instead of 8
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned x = w[j], y = w[j + 4];
                const unsigned fx = s_pair[(x & 0xFu) | ((x >> 12) & 0xF0u)], fy = s_pair[(y & 0xFu) | ((y >> 12) & 0xF0u)];
                // the length fields stay where they are (16 x the length per half word); the sums are scaled back once
                const unsigned lx = x & 0xFFF0FFF0u, ly = y & 0xFFF0FFF0u;
                r0 = __dp2a_lo(lx, fx, r0); q0 = __dp2a_hi(lx, fx, q0);
                rs = __dp2a_lo(ly, fy, rs); qs = __dp2a_hi(ly, fy, qs);
                acc |= x | y;
            }
            rs = (rs + r0) >> 4; qs = (qs + q0) >> 4; r0 >>= 4; q0 >>= 4;
#else
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const unsigned x = w[j];
                const unsigned lens = (x >> 4) & 0x0FFF0FFFu;
                const unsigned fl = s_pair[(x & 0xFu) | ((x >> 12) & 0xF0u)];
                rs = __dp2a_lo(lens, fl, rs);
                qs = __dp2a_hi(lens, fl, qs);
                acc |= x;
                if (j == 3) { r0 = rs; q0 = qs; }
            }
#endif
            // an escaped length (field == 0xFFF) somewhere in the lane's ops?  (the OR can only err towards "yes")
            const bool esc = ((acc >> 4) & 0xFFFu) == 0xFFFu || (acc >> 20) == 0xFFFu;
            if (__any_sync(FULL, esc)) {
                if (esc) {
                    rs = 0; qs = 0;
#pragma unroll 1
                    for (int j = 0; j < 16; j++) {
                        const unsigned x = a0 < cnt16 ? (unsigned)cg[a0 + j] : PAD16;   // from shared memory: no dynamic register indexing
                        const unsigned t = ADV_LUT >> ((x << 1) & 30u);
                        const unsigned len = op_len(a.b, x, gop0 + (uint64_t)(a_base + a0 + j));
                        rs += (t & 1u) * len; qs += ((t >> 1) & 1u) * len;
                        if (j == 7) { r0 = rs; q0 = qs; }
                    }
                }
            }
            // exclusive scan of both cursors: one packed scan when every lane sum fits 11 bits (so the totals fit 16)
            int er, eq, rtot, qtot;
            if (!__any_sync(FULL, (rs | qs) >= 2048u)) {
                const unsigned pk2 = rs | (qs << 16);
                unsigned inc = pk2;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned t = __shfl_up_sync(FULL, inc, d);
                    if (lane >= d) inc += t;
                }
                const unsigned exc = inc - pk2, tt = __shfl_sync(FULL, inc, 31);
                er = (int)(exc & 0xffffu); eq = (int)(exc >> 16);
                rtot = (int)(tt & 0xffffu); qtot = (int)(tt >> 16);
            } else {
                int ri = (int)rs, qi = (int)qs;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(FULL, ri, d);
                    const int u = __shfl_up_sync(FULL, qi, d);
                    if (lane >= d) { ri += t; qi += u; }
                }
                er = ri - (int)rs; eq = qi - (int)qs;
                rtot = __shfl_sync(FULL, ri, 31); qtot = __shfl_sync(FULL, qi, 31);
            }
            *reinterpret_cast<int4 *>(&S.grp[it * 64 + lane * 2]) =
                make_int4(ref_pos + er, qpos + eq, ref_pos + er + (int)r0, qpos + eq + (int)q0);
            // ---- clips (S/H longer than 5) and unsupported ops: op codes with bit 2 or 3 set (S H P = X and 9..15).  Only the pairs
            // that hold such a code are looked at; the reference position of a clip is summed up on demand (clips are rare). ----
            if (__any_sync(FULL, ((acc | (acc >> 16)) & 0xCu) != 0)) {
                unsigned pairs = 0u;                       // bit j: pair j of this lane holds an op code >= 4
#pragma unroll
                for (int j = 0; j < 8; j++) pairs |= ((w[j] & 0x000C000Cu) != 0u ? 1u : 0u) << j;
                while (pairs) {
                    const int j = __ffs(pairs) - 1;
                    pairs &= pairs - 1u;
                    for (int h = 0; h < 2; h++) {
                        const int jj = 2 * j + h;
                        const unsigned x = cg[a0 + jj];    // back from shared memory: no dynamically indexed registers
                        const unsigned op = x & 15u;
                        if (op < 4u || op == 7u || op == 8u) continue;
                        const int g = a_base + a0 + jj - mis;   // CIGAR index inside the read
                        if (!TAG && (op == 4u || op == 5u) && (int)op_len(a.b, x, gop0 + (uint64_t)(a_base + a0 + jj)) > 5) {
                            // getClip (ParsingBam.cpp:1636-1645); a later abort of the read cancels the events at or after the aborting op
                            int rr = ref_pos + er;
                            for (int t = 0; t < jj; t++) {
                                const unsigned y = cg[a0 + t];
                                if ((0x18Du >> (y & 15u)) & 1u) rr += (int)op_len(a.b, y, gop0 + (uint64_t)(a_base + a0 + t));
                            }
                            const unsigned long long slot = atomicAdd(&a.counters->clips.v, 1ull);
                            if (slot < a.clip_cap) {
                                a.clip_keys[slot] = ((uint32_t)rr << 1) | (g == 0 ? 0u : 1u);
                                a.clip_meta[slot] = make_uint2((unsigned)r, (unsigned)g);
                            }
                        }
                        if (op > 8u) bad_op = min(bad_op, g);
                    }
                }
            }
            ref_pos += rtot; qpos += qtot;
        }
        const int ng = n_it * 64;
        __syncwarp();

        // ================= phase 2: every pending variant below ref_pos, one per lane =================
        bool first_round = true;
        while (true) {
            const int vi = cur + lane;
            uint2 vr = vr_first;
            if (!LPS_PREFETCH_VREC || !first_round) { vr = make_uint2((unsigned)INT_MAX, 0u); if (vi < nv) vr = vrec[vi]; }
            first_round = false;
            const int vp = (int)vr.x;
            const bool mine = vp < ref_pos;
            const unsigned mmask = __ballot_sync(FULL, mine);
            if (mmask == 0) break;
            int cand_var = -1; uint32_t cand_x = 0;
            uint4 c4 = make_uint4(0u, 0u, 0u, 0u);
            bool ab = false, in_del = false;
            int o_r = 0, o_q = 0, opi = 0;
            if (mine) {
                // group: the last one that starts at or before vp
                int g = 0;
#pragma unroll
                for (int step = (NGRP >= 256 ? 256 : 128); step >= 1; step >>= 1) {
                    const int midg = g + step;
                    if (midg < ng && S.grp[midg].x <= vp) g = midg;
                }
                const int2 gs = S.grp[g];
                int wr = gs.x, wq = gs.y;
                // covering op: the last op of the group that starts at or before vp.  One 128-bit load; the four PAIRS of ops are summed
                // with the same table + dot products as phase 1, the pair is picked by counting starts <= vp, then one of its two ops
                const uint4 g8 = *reinterpret_cast<const uint4 *>(cg + g * 8);
                const unsigned gw[4] = {g8.x, g8.y, g8.z, g8.w};
                const unsigned gor = g8.x | g8.y | g8.z | g8.w;
                const bool gesc = (((gor >> 4) & 0xFFFu) == 0xFFFu) || ((gor >> 20) == 0xFFFu);
                unsigned c = PAD16;
                int o_len = 0, jsel = 0;
                if (LPS_PAIR_WALK && !gesc) {
                    unsigned fl[4], pr[4], pq[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        fl[k] = s_pair[(gw[k] & 0xFu) | ((gw[k] >> 12) & 0xF0u)];
                        const unsigned lens = (gw[k] >> 4) & 0x0FFF0FFFu;
                        pr[k] = __dp2a_lo(lens, fl[k], 0u);
                        pq[k] = __dp2a_hi(lens, fl[k], 0u);
                    }
                    const int s1 = wr + (int)pr[0], s2 = s1 + (int)pr[1], s3 = s2 + (int)pr[2];
                    const int q1 = wq + (int)pq[0], q2 = q1 + (int)pq[1], q3 = q2 + (int)pq[2];
                    const int k = (s1 <= vp) + (s2 <= vp) + (s3 <= vp);          // starts never decrease: the last pair that starts at or before vp
                    const unsigned wsel = k == 0 ? gw[0] : k == 1 ? gw[1] : k == 2 ? gw[2] : gw[3];
                    const unsigned fsel = k == 0 ? fl[0] : k == 1 ? fl[1] : k == 2 ? fl[2] : fl[3];
                    const int ps = k == 0 ? wr : k == 1 ? s1 : k == 2 ? s2 : s3, qs_ = k == 0 ? wq : k == 1 ? q1 : k == 2 ? q2 : q3;
                    const int lenA = (int)((wsel >> 4) & 0xFFFu), lenB = (int)(wsel >> 20);
                    const int sB = ps + (int)(fsel & 1u) * lenA, qB = qs_ + (int)((fsel >> 16) & 1u) * lenA;
                    const bool second = sB <= vp;
                    c = second ? (wsel >> 16) : (wsel & 0xFFFFu);
                    o_r = second ? sB : ps; o_q = second ? qB : qs_; o_len = second ? lenB : lenA;
                    jsel = 2 * k + (second ? 1 : 0);
                } else {
#pragma unroll 1
                    for (int j = 0; j < 8; j++) {
                        const unsigned cc = (unsigned)cg[g * 8 + j];
                        const int len = (int)op_len(a.b, cc, gop0 + (uint64_t)(a_base + g * 8 + j));
                        if (wr <= vp) { c = cc; o_r = wr; o_q = wq; jsel = j; o_len = len; }
                        const unsigned t = ADV_LUT >> ((cc << 1) & 30u);
                        wr += (int)(t & 1u) * len;
                        wq += (int)((t >> 1) & 1u) * len;
                    }
                }
                const int lidx = g * 8 + jsel;            // position inside the super-chunk; the buffer also holds position lidx + 1
                const int o_op = (int)(c & 15u);
                opi = a_base + lidx - mis;
                const unsigned nop = (unsigned)cg[lidx + 1] & 15u;   // only meaningful when opi + 1 < ncig
                const bool rl1 = (vr.y & VF_REF1) != 0, al1 = (vr.y & VF_ALT1) != 0;
                if (SOM) {
                    // union map: every variant inside an M/=/X op (HaplotagParsingBam.cpp:585-614) or a D op (:623-631) becomes a
                    // raw candidate {variant, query index, op index | op start, offset | flags}; the hooks run in resolve_somatic
                    if (o_op == 0 || o_op == 7 || o_op == 8) {
                        const int off = vp - o_r;
                        if (o_q + off < lq) {                      // beyond SEQ the reference reads undefined memory: dropped
                            unsigned fl = 0;
                            if (opi + 1 < ncig) {
                                fl = 1u;                           // i + 1 < n_cigar
                                if (o_r + o_len - 1 == vp) fl |= (nop == 1u ? 2u : 0u) | (nop == 2u ? 4u : 0u);
                            }
                            cand_var = vi;
                            c4 = make_uint4((unsigned)vi, (unsigned)(o_q + off), (unsigned)opi, (unsigned)off | (fl << 28));
                            prefetch_base(o_q + off);
                        }
                    } else if (o_op == 2) {
                        cand_var = vi;
                        c4 = make_uint4((unsigned)vi, (unsigned)o_q, (unsigned)o_r, 8u << 28);
                    }
                } else if (TAG) {
                    // CigarParser::parsingCigar M branch (HaplotagParsingBam.cpp:585-614) + judgeSnpHap (HaplotagStrategy.cpp:20-130)
                    if (o_op == 0 || o_op == 7 || o_op == 8) {
                        const int off = vp - o_r;
                        if (rl1 && al1) {
                            // the reference reads seq[query_pos+offset] without a bounds check; past l_qseq that is
                            // memory of the BAM record (undefined) — such a hit is dropped here
                            if (o_q + off < lq) { cand_var = vi; cand_x = (uint32_t)(o_q + off); prefetch_base(o_q + off); }
                        } else if (rl1 != al1) {
                            if (opi + 1 < ncig) {
                                const unsigned want = rl1 ? 1u : 2u;
                                const bool has = (o_r + o_len - 1 == vp && nop == want);
                                const bool h1alt = (vr.y & VF_HP1ALT) != 0;
                                const bool l1_is1 = h1alt ? al1 : rl1, l2_is1 = h1alt ? rl1 : al1;   // lengths of the HP1 / HP2 allele strings == 1
                                // read shows the indel: the haplotype whose allele string is not 1 long gets the vote, else the other
                                int hpbit = -1;
                                if (!l1_is1 && l2_is1) hpbit = has ? 0 : 1;
                                else if (l1_is1 && !l2_is1) hpbit = has ? 1 : 0;
                                cand_var = vi;
                                cand_x = (2u << 30) | (hpbit >= 0 ? 2u : 0u) | (unsigned)(hpbit > 0);
                            }
                        }
                    } else if (o_op == 2) in_del = true;
                } else if (o_op == 0 || o_op == 7 || o_op == 8) {
                    const int off = vp - o_r;
                    if (o_q + off + 1 > lq) ab = true;                                              // :1453-1455
                    else {
                        if (rl1 && al1) { cand_var = vi; cand_x = (uint32_t)(o_q + off); prefetch_base(o_q + off); }
                        else if (rl1 != al1) {
                            if (opi + 1 < ncig) {                                                   // :1470, :1495
                                const unsigned want = rl1 ? 1u : 2u;                                // I after an insertion anchor, D after a deletion anchor
                                const int allele = (o_r + o_len - 1 == vp && nop == want) ? 1 : 0;
                                cand_var = vi;
                                cand_x = (2u << 30) | ((unsigned)allele << 1) | ((vr.y & VF_DANGER) ? 1u : 0u);
                            }
                        }
                    }
                } else if (o_op == 2) in_del = true;   // decided below, with the position of the previous variant
            }
            // D-op rule: only the FIRST pending variant of the op (the previous variant lies before the op)
            if (!SOM && in_del) {
                const int pv = vi > 0 ? (int)vrec[vi - 1].x : INT_MIN;
                if (a.have_reference && pv < o_r && (vr.y >> 24) >= 3u) {
                    const bool rl1 = (vr.y & VF_REF1) != 0, al1 = (vr.y & VF_ALT1) != 0;
                    if (TAG) {
                        // processDeletionOperation (HaplotagProcess.cpp:492-501): first variant of the D op only;
                        // judgeDeletionHap (HaplotagStrategy.cpp:147-209): homopolymer >= 3, SNP compares the next aligned base
                        if (rl1 && al1) {
                            if (o_q < lq) { cand_var = vi; cand_x = (1u << 30) | (uint32_t)o_q; prefetch_base(o_q); }
                        } else if (!rl1 && al1) {
                            const bool h1alt = (vr.y & VF_HP1ALT) != 0;
                            const bool l1_is1 = h1alt ? al1 : rl1, l2_is1 = h1alt ? rl1 : al1;
                            int hpbit = -1;
                            if (!l1_is1 && l2_is1) hpbit = 0; else if (l1_is1 && !l2_is1) hpbit = 1;
                            cand_var = vi;
                            cand_x = (2u << 30) | (hpbit >= 0 ? 2u : 0u) | (unsigned)(hpbit > 0);
                        }
                    } else {
                        // get_snp D branch (:1539-1607)
                        if (o_q + 1 > lq) ab = true;                                                // :1559-1561
                        else if (rl1 && al1) { cand_var = vi; cand_x = (1u << 30) | (uint32_t)o_q; prefetch_base(o_q); }
                        else if (!rl1 && al1) { cand_var = vi; cand_x = (2u << 30) | (1u << 2) | (1u << 1); }
                    }
                }
            }
            // the first aborting variant (in order) drops the read; nothing after it matters
            const unsigned abmask = __ballot_sync(FULL, ab);
            unsigned keep = __ballot_sync(FULL, cand_var >= 0);
            if (abmask) {
                const int fl = __ffs(abmask) - 1;
                abort_op = __shfl_sync(FULL, opi, fl);
                aborted = true;
                keep &= (1u << fl) - 1u;
            }
            if (cand_var >= 0 && ((keep >> lane) & 1u)) {
                const int dst = ncand + __popc(keep & ((1u << lane) - 1u));
                if (dst < cand_cap) {
                    if (SOM) cand4[dst] = c4;
                    else { cand[dst].var = cand_var; cand[dst].x = cand_x; }
                }
            }
            ncand += __popc(keep);
            if (aborted) break;
            const int handled = __popc(mmask);
            cur += handled;
            if (handled < 32) break;
        }
        __syncwarp();   // the next super-chunk overwrites the index; the buffer may be refilled from the next iteration on
        if (aborted) {
            // the remaining super-chunks are not decoded, but the copies already requested must be consumed to keep the
            // pipeline's parities in step: drain the one in flight (if it belongs to this read) and request the next read's
            if (s + 1 < n_sc) {
                const int b2 = pp.buf;
                mbar_wait(smem_u32(&S.mbar[b2]), (pp.phase >> b2) & 1u);
                pp.phase ^= 1u << b2;
                __syncwarp();
                if (next.r != 0xFFFFFFFFu && next.ncig > 0) issue_sc(a, S, next, 0, b2, lane);
            }
            break;
        }
    }
    const bool bad = __any_sync(FULL, bad_op < abort_op);
    if (bad && lane == 0) atomicAdd(&a.counters->bad_cigar.v, 1ull);
    if (!TAG && lane == 0 && aborted) {
        a.abort_of_read[r] = abort_op;
        atomicAdd(&a.counters->aborted_reads.v, 1ull);
    }

    if (aborted || bad) {
        if (lane == 0) { a.ncalls[r] = 0; a.tmp_start[r] = 0; a.status[r] = LPS_READ_ABORTED; }
        return;
    }
    if (ncand > cand_cap) {
        // rare: more candidates than the shared buffer holds — redo this read in the overflow pass
        if (lane == 0) {
            unsigned k = (unsigned)atomicAdd(&a.counters->overflow_reads.v, 1ull);
            atomicAdd(&a.counters->overflow_cands.v, (unsigned long long)ncand);
            if (k < a.overflow_list_cap) { a.overflow_list_out[k] = work_slot; a.overflow_need_out[k] = (uint64_t)ncand; }
            a.ncalls[r] = 0; a.tmp_start[r] = 0; a.status[r] = LPS_READ_OK;
        }
        return;
    }
    __syncwarp();

    if (SOM) {
        resolve_somatic<MODE>(a, r, lane, cand4, ncand, ref_start, ref_pos, qpos, lq, pc, seq, (int)(D.flags & 0xFFu) >= a.mapping_quality);
        return;
    }
    if (TAG) {
        // ---- resolve: per candidate the haplotype bit and the "counts towards countPS" flag, then judgeReadHap ----
        int h1 = 0, h2 = 0, ps_min = INT_MAX, ps_max = INT_MIN, nout = 0, ngather = 0;
        for (int c0 = 0; c0 < ncand; c0 += 32) {
            const int c = c0 + lane;
            int hpbit = -1; bool ps_counted = false; int var = -1; unsigned kind = 0;
            if (c < ncand) {
                const Cand cd = cand[c];
                kind = cd.x >> 30; var = cd.var;
                if (kind == 2) { ps_counted = true; if (cd.x & 2u) hpbit = (int)(cd.x & 1u); }
                else {
                    ngather++;
                    const int qi = (int)(cd.x & 0x3fffffffu);
                    const unsigned code = (seq[qi >> 1] >> ((~qi & 1) << 2)) & 0xfu;
                    const char base = "=ACMGRSVTWYHKDBN"[code];
                    const unsigned vy = vrec[var].y;
                    const char rb = (char)(vy & 0xFFu), ab = (char)((vy >> 8) & 0xFFu);
                    const bool h1alt = (vy & VF_HP1ALT) != 0;
                    const char hp1 = h1alt ? ab : rb, hp2 = h1alt ? rb : ab;
                    // M op: only a REF/ALT base counts at all (HaplotagStrategy.cpp:39); D-op rule: countPS unconditionally (:186)
                    if (kind == 1 || base == rb || base == ab) {
                        ps_counted = true;
                        if (base == hp1) hpbit = 0;
                        if (base == hp2) hpbit = 1;     // HP2 assigned last, as in the reference (:52-59)
                    }
                }
            }
            if (ps_counted) { const int ps = a.var_ps[var]; ps_min = min(ps_min, ps); ps_max = max(ps_max, ps); }
            if (hpbit == 0) h1++; else if (hpbit == 1) h2++;
            // a base equal to both HP1 and HP2 cannot happen for a het; H1 and H2 counts are therefore exclusive
            if (a.want_calls) {
                const unsigned m = __ballot_sync(FULL, ps_counted);
                __syncwarp();
                if (ps_counted) {
                    const int dst = nout + __popc(m & ((1u << lane) - 1u));
                    Cand packed;
                    packed.var = var;
                    packed.x = ((uint32_t)(uint16_t)(int16_t)1) | ((uint32_t)(uint8_t)(int8_t)hpbit << 16) | ((uint32_t)kind << 24);
                    cand[dst] = packed;
                }
                nout += __popc(m);
                __syncwarp();
            }
        }
        if (a.count_gathers) {
            ngather = (int)__reduce_add_sync(FULL, (unsigned)ngather);
            if (lane == 0 && ngather) atomicAdd(&a.counters->gathers.v, (unsigned long long)ngather);
        }
        h1 = (int)__reduce_add_sync(FULL, (unsigned)h1);
        h2 = (int)__reduce_add_sync(FULL, (unsigned)h2);
        ps_min = __reduce_min_sync(FULL, ps_min);
        ps_max = __reduce_max_sync(FULL, ps_max);
        if (lane == 0) {
            // GermlineHaplotagStrategy::judgeReadHap (HaplotagStrategy.cpp:243-300)
            const double mx = (double)(h1 > h2 ? h1 : h2), mn = (double)(h1 > h2 ? h2 : h1);
            int hp = 0;
            if (!(mx / (mx + mn) < a.percentage)) { if (h1 > h2) hp = 1; if (h1 < h2) hp = 2; }
            if (ps_min != INT_MAX && ps_min != ps_max) hp = 0;                  // countPS.size() > 1: crosses two blocks
            const int imx = h1 > h2 ? h1 : h2, imn = h1 > h2 ? h2 : h1;
            int pq;
            if (imx == 0) pq = 0; else if (imn == 0) pq = 40;
            else pq = (imx < 256) ? (int)a.pq_lut[imn * 256 + imx] : -1;         // -1: the host fills it in with its libm
            a.tag_hp[r] = (int8_t)hp; a.tag_ps[r] = hp ? ps_min : 0; a.tag_pq[r] = pq; a.tag_h1[r] = h1; a.tag_h2[r] = h2;
        }
        const unsigned long long start = pool_alloc(a, pc, nout, lane);
        if (lane == 0) { a.ncalls[r] = (uint32_t)nout; a.tmp_start[r] = start; a.status[r] = LPS_READ_OK; }
        if (start + nout <= a.calls_cap) {
            for (int c = lane; c < nout; c += 32) {
                const Cand cd = cand[c];
                lps_call out;
                out.var = cd.var;
                out.quality = (int16_t)(cd.x & 0xffffu);
                out.allele = (int8_t)((cd.x >> 16) & 0xffu);
                out.origin = (int8_t)((cd.x >> 24) & 0xffu);
                a.calls_tmp[start + c] = out;
            }
        }
        return;
    }

    // ---- resolve candidates: gather base + quality, decide the allele, drop filterSNP variants ----
    int nvalid = 0, ngather = 0;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
        const int c = c0 + lane;
        bool valid = false;
        lps_call out;
        out.var = 0; out.quality = 0; out.allele = 0; out.origin = 0;
        if (c < ncand) {
            const Cand cd = cand[c];
            const unsigned kind = cd.x >> 30;
            out.var = cd.var;
            const unsigned vy = vrec[cd.var].y;
            if (kind == 2) {
                out.allele = (int8_t)((cd.x >> 1) & 1u);
                out.origin = (int8_t)((cd.x >> 2) & 1u);
                out.quality = (cd.x & 1u) ? -5 : -4;
                valid = true;
            } else {
                ngather++;
                const int qi = (int)(cd.x & 0x3fffffffu);
                unsigned code, q;
                if (a.b.sq) {
                    // one aligned 16-byte load: the unit of ten bases that holds this query index (lps_read_batch.sq)
                    const uint32_t u = sq_unit_of((uint32_t)qi);
                    const uint4 w = __ldg(reinterpret_cast<const uint4 *>(seq) + u);
                    sq_extract(w.x, w.y, w.z, w.w, (uint32_t)qi - u * SQ_BASES, code, q);
                } else {
                    const unsigned byte = seq[qi >> 1];
                    code = (byte >> ((~qi & 1) << 2)) & 0xfu;                        // bam_seqi
#ifdef LPS_DEBUG_NO_QUAL
                    q = 30;                              // timing experiment only (bench refuses the result): what do the QUAL gathers cost over PCIe?
#else
                    q = qual[qi];
#endif
                }
                const char base = "=ACMGRSVTWYHKDBN"[code];                          // seq_nt16_str
                const char rb = (char)(vy & 0xFFu), ab = (char)((vy >> 8) & 0xFFu);
                out.quality = (int16_t)q;
                out.origin = (int8_t)kind;
                if (base == rb) { out.allele = 0; valid = true; }
                else if (base == ab) { out.allele = 1; valid = true; }
            }
            if (valid && a.apply_filter && (vy & VF_FILTERED)) valid = false;
        }
        // compact inside the candidate buffer (reused as the staging area of the final write)
        const unsigned m = __ballot_sync(FULL, valid);
        __syncwarp();
        if (valid) {
            const int dst = nvalid + __popc(m & ((1u << lane) - 1u));
            Cand packed;
            packed.var = out.var;
            packed.x = ((uint32_t)(uint16_t)out.quality) | ((uint32_t)(uint8_t)out.allele << 16) | ((uint32_t)(uint8_t)out.origin << 24);
            cand[dst] = packed;   // dst <= c, and every read of this round happened before the __syncwarp
        }
        nvalid += __popc(m);
        __syncwarp();
    }
    if (a.count_gathers) {
        ngather = (int)__reduce_add_sync(FULL, (unsigned)ngather);
        if (lane == 0 && ngather) atomicAdd(&a.counters->gathers.v, (unsigned long long)ngather);
    }
    const unsigned long long start = pool_alloc(a, pc, nvalid, lane);
    if (lane == 0) { a.ncalls[r] = (uint32_t)nvalid; a.tmp_start[r] = start; a.status[r] = LPS_READ_OK; }
    if (start + nvalid <= a.calls_cap) {
        for (int c = lane; c < nvalid; c += 32) {
            const Cand cd = cand[c];
            lps_call out;
            out.var = cd.var;
            out.quality = (int16_t)(cd.x & 0xffffu);
            out.allele = (int8_t)((cd.x >> 16) & 0xffu);
            out.origin = (int8_t)((cd.x >> 24) & 0xffu);
            a.calls_tmp[start + c] = out;
        }
    }
}

__device__ __forceinline__ ReadDesc load_desc(const uint4 *slot) { return unpack_desc(slot[0], slot[1], slot[2]); }

// Persistent launch: every warp takes its reads from a global counter until the work list is exhausted, so a CTA is never held
// hostage by its longest read (read lengths are log-normal).  Three descriptors are resident per warp (the read being walked and
// the next two), fetched with cp.async two reads ahead; the claim that feeds the ring is issued one read ahead.  The overflow pass
// walks its list one read per warp.
template <int MODE>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 1) k_call_alleles(K1Args a) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    WarpScratch *s_all = reinterpret_cast<WarpScratch *>(s_dyn);
    // flags of a PAIR of op codes (low 4 bits: first op, high 4 bits: second op), one byte each:
    // consumes-reference(op0), consumes-reference(op1), consumes-query(op0), consumes-query(op1)  ->  operands of IDP.2A lo / hi
    __shared__ uint32_t s_pair[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const unsigned t0 = (ADV_LUT >> (2 * (i & 15))) & 3u, t1 = (ADV_LUT >> (2 * (i >> 4))) & 3u;
        s_pair[i] = (t0 & 1u) | ((t1 & 1u) << 8) | ((t0 >> 1) << 16) | ((t1 >> 1) << 24);
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpScratch &S = s_all[wib];
    if (lane == 0) { mbar_init(smem_u32(&S.mbar[0]), 1u); mbar_init(smem_u32(&S.mbar[1]), 1u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    PoolCursor pc;
    pc.next = 0; pc.end = 0; pc.used = 0;
    Pipe pp;
    pp.buf = 0; pp.phase = 0u;
    ReadDesc none;
    none.cigar_off = 0; none.seq_off = 0; none.qual_off = 0; none.ref_start = 0; none.lq = 0; none.ncig = 0; none.first_var = 0;
    none.r = 0xFFFFFFFFu; none.flags = 0;
    unsigned long long dbg_t0 = 0, dbg_n = 0;
    if (a.dbg_times) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
    if (a.overflow_reads != nullptr) {
        const long long wid = (long long)blockIdx.x * WARPS_PER_CTA + wib;
        if (wid >= a.overflow_list_cap) return;
        const uint32_t slot = a.overflow_reads[wid];
        if (lane < 3) cp_async16(smem_u32(&S.desc[0][lane]), a.work + (size_t)slot * 3 + lane);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        const ReadDesc cur = load_desc(S.desc[0]);
        if (cur.ncig > 0) issue_sc(a, S, cur, 0, 0, lane);
        process_read<MODE>(a, S, s_pair, cur, none, pp, a.overflow_buf + a.overflow_off[wid], (int)(a.overflow_off[wid + 1] - a.overflow_off[wid]),
                           slot, lane, pc);
        if (lane == 0 && pc.used) atomicAdd(&a.counters->n_calls.v, pc.used);
        return;
    }
    const uint32_t c0 = a.seg_count[0], c1 = a.seg_count[1], c2 = a.seg_count[2], n_items = c0 + c1 + c2 + a.seg_count[3];
    // slot of the descriptor array each ring entry came from (reported when a read overflows the candidate buffer)
    uint32_t ring_slot[3] = {0u, 0u, 0u};
    auto fetch = [&](int k, uint32_t t) -> uint32_t {
        uint32_t slot = 0;
        if (t < n_items) {
            slot = slot_of_item(a, t, c0, c1, c2);
            if (lane < 3) cp_async16(smem_u32(&S.desc[k][lane]), a.work + (size_t)slot * 3 + lane);
        } else if (lane == 2) S.desc[k][2].z = 0xFFFFFFFFu;
        return slot;
    };
    // Work distribution.  ~100 k claims on ONE counter saturate its L2 slice (the atomic's round trip grew to a read's duration and
    // showed up as the largest long-scoreboard stall of the kernel), so the first LPS_STATIC_SHARE per cent of the list - which
    // starts with the long reads - is dealt out without atomics, item k * warps + warp to every warp, and only the rest, which
    // evens out the warps' finishing times, is claimed from the counter, one claim ahead of its use.
    const uint32_t n_warps = gridDim.x * WARPS_PER_CTA, w_global = blockIdx.x * WARPS_PER_CTA + wib;
    const uint32_t n_static = (uint32_t)(((unsigned long long)n_items * LPS_STATIC_SHARE / 100ull) / n_warps);   // items per warp
    const uint32_t dyn_base = n_static * n_warps;
    uint32_t k_static = 0, claim = 0;
    if (lane == 0) claim = (uint32_t)atomicAdd(&a.counters->next_read.v, 1ull);       // first dynamic item; stays in flight while the static share lasts
    auto next_item = [&]() -> uint32_t {
        if (k_static < n_static) return (k_static++) * n_warps + w_global;
        const uint32_t t = dyn_base + __shfl_sync(FULL, claim, 0);
        if (lane == 0) claim = (uint32_t)atomicAdd(&a.counters->next_read.v, 1ull);
        return t;
    };
    ring_slot[0] = fetch(0, next_item()); ring_slot[1] = fetch(1, next_item()); ring_slot[2] = fetch(2, next_item());
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    {
        const ReadDesc first = load_desc(S.desc[0]);
        if (first.r != 0xFFFFFFFFu && first.ncig > 0) issue_sc(a, S, first, 0, 0, lane);
    }
    int sl = 0;
    while (true) {
        cp_async_wait<1>();            // everything but the refill issued at the end of the previous read has landed
        __syncwarp();
        const ReadDesc cur = load_desc(S.desc[sl]);
        if (cur.r == 0xFFFFFFFFu) break;
        const int sn = sl == 2 ? 0 : sl + 1;
        const ReadDesc next = load_desc(S.desc[sn]);
        process_read<MODE>(a, S, s_pair, cur, next, pp, S.cand, CAND_CAP, sl == 0 ? ring_slot[0] : sl == 1 ? ring_slot[1] : ring_slot[2], lane, pc);
        __syncwarp();
        // refill this ring entry with the read claimed one read ago; claim the one after it
        const uint32_t t = next_item();
        const uint32_t ns = fetch(sl, t);
        if (sl == 0) ring_slot[0] = ns; else if (sl == 1) ring_slot[1] = ns; else ring_slot[2] = ns;
        cp_async_commit();
        sl = sn;
        dbg_n++;
    }
    cp_async_wait<0>();
    if (lane == 0 && pc.used) atomicAdd(&a.counters->n_calls.v, pc.used);
    if (a.dbg_times && lane == 0) {
        unsigned long long t1; unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long *d = a.dbg_times + 4ull * ((unsigned long long)blockIdx.x * WARPS_PER_CTA + wib);
        d[0] = dbg_t0; d[1] = t1; d[2] = dbg_n; d[3] = smid;
    }
}

// ---- k_prep_reads: one thread per alignment ------------------------------------------------------------------------------
// The read filter of BamParser::direct_detect_alleles (iterator region "chr:1-lastSNP", ParsingBam.cpp:1273; MAPQ / flag filter
// :1282-1291) or the dispatch of ChromosomeProcessor::processSingleChrom (HaplotagParsingBam.cpp:457-486, in its order): alignments
// that are not walked get their (empty) products here and never reach k_call_alleles.  Every other alignment gets a 48-byte
// descriptor - with lower_bound(variants, ref_start), the reference's stateful firstVariantIter (ParsingBam.cpp:1318-1319,
// HaplotagParsingBam.cpp:555-563), which equals it for a coordinate-sorted batch - appended to one of four work segments: reads
// with more than thr0 / thr1 / thr2 CIGAR ops (> 4x, > 2.5x, > 1.6x the mean) and the rest.  The walking kernel starts the long ones
// first, so that they do not form its tail.
struct PrepArgs {
    DevBatch b;
    int nv;
    const int32_t *vpos;
    int mode, mapping_quality, mapq_filter, tag_supplementary, last_var_pos;
    uint32_t thr[3];
    uint32_t seg_base[4];
    uint4 *work;
    uint32_t *seg_count;
    uint64_t *tmp_start;
    uint32_t *ncalls;
    uint8_t *status;
    int32_t *abort_of_read;
    int8_t *tag_hp;
    int32_t *tag_ps, *tag_pq, *tag_h1, *tag_h2;
    uint8_t *tag_cat;
    int32_t *tag_h3, *tag_end, *tag_len;
    uint8_t *tag_nps;
    int8_t *tag_hpb;
    float *tag_sim;
};

__global__ void __launch_bounds__(256) k_prep_reads(PrepArgs p) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in_range = r < p.b.n_reads;
    int seg = -1;
    int ref_start = 0, lq = 0, flag = 0, mapq = 0;
    uint32_t ncig = 0;
    if (in_range) {
        ref_start = p.b.ref_start[r]; lq = p.b.l_qseq[r]; ncig = p.b.n_cigar[r]; flag = p.b.flag[r]; mapq = p.b.mapq[r];
        const bool tag = p.mode != LPS_MODE_PHASE, som = p.mode >= LPS_MODE_EXTRACT_NORMAL;
        bool walk;
        if (!tag) {
            walk = !(ref_start >= p.last_var_pos || mapq < p.mapping_quality || (flag & (0x4 | 0x100 | 0x400)));
            p.abort_of_read[r] = INT_MAX;
        } else {
            int cat = LPS_TAG_PROCESSED;
            if (mapq < p.mapping_quality && p.mapq_filter) cat = LPS_TAG_LOW_MAPQ;
            else if (flag & 0x4) cat = LPS_TAG_UNMAPPED;
            else if (flag & 0x100) cat = LPS_TAG_SECONDARY;
            else if ((flag & 0x800) && !p.tag_supplementary) cat = LPS_TAG_SUPPLEMENTARY;
            else if (p.nv == 0) cat = LPS_TAG_EMPTY_VARIANTS;
            else if (!(ref_start <= p.last_var_pos)) cat = LPS_TAG_OTHER;
            p.tag_cat[r] = (uint8_t)cat;
            walk = cat == LPS_TAG_PROCESSED;
            if (!walk) {
                p.tag_hp[r] = 0; p.tag_ps[r] = 0; p.tag_pq[r] = 0; p.tag_h1[r] = 0; p.tag_h2[r] = 0;
                if (som) { p.tag_h3[r] = 0; p.tag_end[r] = 0; p.tag_len[r] = 0; p.tag_nps[r] = 0; p.tag_hpb[r] = 0; p.tag_sim[r] = 0.f; }
            }
        }
        if (!walk) { p.ncalls[r] = 0; p.tmp_start[r] = 0; p.status[r] = LPS_READ_FILTERED; }
        else seg = ncig > p.thr[0] ? 0 : ncig > p.thr[1] ? 1 : ncig > p.thr[2] ? 2 : 3;
    }
    // one atomic per warp and segment
    uint32_t slot = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned m = __ballot_sync(FULL, seg == k);
        if (m == 0) continue;
        uint32_t base = 0;
        const int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(p.seg_count + k, (uint32_t)__popc(m));
        base = __shfl_sync(FULL, base, leader);
        if (seg == k) slot = p.seg_base[k] + base + (uint32_t)__popc(m & ((1u << lane) - 1u));
    }
    if (seg < 0) return;
    int lo = 0, hi = p.nv;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (p.vpos[mid] < ref_start) lo = mid + 1; else hi = mid;
    }
    const uint64_t co = p.b.cigar_off[r], so = p.b.seq_off[r], qo = p.b.sq ? 0ull : p.b.qual_off[r];
    uint4 *w = p.work + (size_t)slot * 3;
    w[0] = make_uint4((uint32_t)co, (uint32_t)(co >> 32), (uint32_t)so, (uint32_t)(so >> 32));
    w[1] = make_uint4((uint32_t)qo, (uint32_t)(qo >> 32), (uint32_t)ref_start, (uint32_t)lq);
    w[2] = make_uint4(ncig, (uint32_t)lo, (uint32_t)r, (uint32_t)mapq | ((uint32_t)flag << 8));
}

// clip events of aborted reads: the reference stops walking at the aborting op, so events at or after it never happened
__global__ void k_clip_filter(unsigned long long n, uint32_t *__restrict__ keys, const uint2 *__restrict__ meta,
                              const int32_t *__restrict__ abort_of_read) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 m = meta[i];
    if ((int)m.y >= abort_of_read[m.x]) keys[i] = 0xFFFFFFFFu;
}

// scratch pool (allocation order) -> CSR in read order; one warp per read
__global__ void k_gather_calls(int n_reads, const uint64_t *__restrict__ tmp_start, const uint32_t *__restrict__ ncalls,
                               const uint64_t *__restrict__ call_off, const lps_call *__restrict__ tmp,
                               lps_call *__restrict__ out) {
    // sixteen lanes per read (a read carries ~18 calls)
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    int lane = threadIdx.x & 15;
    if (wid >= n_reads) return;
    uint32_t n = ncalls[wid];
    uint64_t s = tmp_start[wid], d = call_off[wid];
    const uint64_t *src = reinterpret_cast<const uint64_t *>(tmp);
    uint64_t *dst = reinterpret_cast<uint64_t *>(out);
    for (uint32_t c = lane; c < n; c += 16) dst[d + c] = src[s + c];
}

__global__ void k_widen_u32(int n, const uint32_t *__restrict__ in, uint64_t *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

}  // namespace

static_assert(sizeof(lps_call) == 8, "lps_call must be 8 bytes");

namespace {
constexpr size_t K1_SMEM = sizeof(WarpScratch) * WARPS_PER_CTA;

template <int MODE> cudaError_t prepare_k1() {
    cudaError_t e = cudaFuncSetAttribute(k_call_alleles<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_call_alleles<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (getenv("LPS_DEBUG_OCC")) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_call_alleles<MODE>, WARPS_PER_CTA * 32, K1_SMEM);
        fprintf(stderr, "k_call_alleles<%d>: %d resident CTAs per SM, %zu bytes of dynamic shared memory\n", MODE, nb, K1_SMEM);
    }
    return cudaSuccess;
}

void launch_k1(int mode, int grid, cudaStream_t st, const K1Args &a) {
    const int tb = WARPS_PER_CTA * 32;
    switch (mode) {
        case LPS_MODE_PHASE: k_call_alleles<LPS_MODE_PHASE><<<grid, tb, K1_SMEM, st>>>(a); break;
        case LPS_MODE_GERMLINE: k_call_alleles<LPS_MODE_GERMLINE><<<grid, tb, K1_SMEM, st>>>(a); break;
        case LPS_MODE_EXTRACT_NORMAL: k_call_alleles<LPS_MODE_EXTRACT_NORMAL><<<grid, tb, K1_SMEM, st>>>(a); break;
        case LPS_MODE_EXTRACT_TUMOR: k_call_alleles<LPS_MODE_EXTRACT_TUMOR><<<grid, tb, K1_SMEM, st>>>(a); break;
        default: k_call_alleles<LPS_MODE_SOMATIC_TAG><<<grid, tb, K1_SMEM, st>>>(a); break;
    }
}
}  // namespace

// Function attributes belong to the device (its primary context), not to the process: called from lps_ctx_create with the
// context's device current, so every GPU a process opens a context on gets the shared-memory opt-in.
int lps_prepare_call_alleles(lps_ctx *ctx) {
    LPS_CUDA(ctx, prepare_k1<LPS_MODE_PHASE>());
    LPS_CUDA(ctx, prepare_k1<LPS_MODE_GERMLINE>());
    LPS_CUDA(ctx, prepare_k1<LPS_MODE_EXTRACT_NORMAL>());
    LPS_CUDA(ctx, prepare_k1<LPS_MODE_EXTRACT_TUMOR>());
    LPS_CUDA(ctx, prepare_k1<LPS_MODE_SOMATIC_TAG>());
    return LPS_OK;
}

int lps_launch_call_alleles(lps_ctx *ctx, const lps_phase_params *p, const lps_tag_params *t, int want_calls, int mode, bool defer_clips) {
    const bool tag = t != nullptr;
    if (mode < 0) mode = tag ? LPS_MODE_GERMLINE : LPS_MODE_PHASE;
    const bool som = mode >= LPS_MODE_EXTRACT_NORMAL;
    if (ctx->batch.sq && mode != LPS_MODE_PHASE)
        return ctx->fail(LPS_E_STATE, "a batch with interleaved SEQ + QUAL rows (lps_read_batch.sq) serves the phase calls only");
    const int n = ctx->batch.n_reads;
    const int nv = ctx->var.n;
    cudaStream_t st = ctx->stream;
    LPS_CUDA(ctx, ctx->d_tmp_start.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_call_off.reserve((size_t)n + 2));
    LPS_CUDA(ctx, ctx->d_ncalls.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_status.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_counters.reserve(1));
    LPS_CUDA(ctx, ctx->d_abort_of_read.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_seg_count.reserve(4));
    if (tag) {
        LPS_CUDA(ctx, ctx->d_tag_hp.reserve((size_t)n + 1)); LPS_CUDA(ctx, ctx->d_tag_ps.reserve((size_t)n + 1));
        LPS_CUDA(ctx, ctx->d_tag_pq.reserve((size_t)n + 1)); LPS_CUDA(ctx, ctx->d_tag_h1.reserve((size_t)n + 1));
        LPS_CUDA(ctx, ctx->d_tag_h2.reserve((size_t)n + 1)); LPS_CUDA(ctx, ctx->d_tag_cat.reserve((size_t)n + 1));
    }
    if (som) {
        LPS_CUDA(ctx, ctx->d_tag_h3.reserve((size_t)n + 1)); LPS_CUDA(ctx, ctx->d_tag_end.reserve((size_t)n + 1));
        LPS_CUDA(ctx, ctx->d_tag_len.reserve((size_t)n + 1)); LPS_CUDA(ctx, ctx->d_tag_nps.reserve((size_t)n + 1));
        LPS_CUDA(ctx, ctx->d_tag_hpb.reserve((size_t)n + 1)); LPS_CUDA(ctx, ctx->d_tag_sim.reserve((size_t)n + 1));
    }
    // ---- work list: descriptors of the alignments that are walked, long reads first ----
    // Segment capacities follow from Markov's inequality on the batch's own mean: at most n / k reads have more than k x mean ops.
    const double mean_ops = n > 0 ? (double)ctx->batch.cigar_len / (double)n : 0.0;
    const char *lpt_env = getenv("LPS_LPT");
    const bool lpt = !(lpt_env && lpt_env[0] == '0');
    PrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.thr[0] = lpt ? (uint32_t)(4.0 * mean_ops) + 64 : 0xFFFFFFFFu;
    pa.thr[1] = lpt ? (uint32_t)(2.5 * mean_ops) + 64 : 0xFFFFFFFFu;
    pa.thr[2] = lpt ? (uint32_t)(1.6 * mean_ops) + 64 : 0xFFFFFFFFu;
    const size_t seg_cap[4] = {(size_t)n / 4 + 2, (size_t)(n / 2.5) + 2, (size_t)(n / 1.6) + 2, (size_t)n + 2};
    pa.seg_base[0] = 0;
    for (int k = 1; k < 4; k++) pa.seg_base[k] = pa.seg_base[k - 1] + (uint32_t)seg_cap[k - 1];
    LPS_CUDA(ctx, ctx->d_work.reserve(3 * ((size_t)pa.seg_base[3] + seg_cap[3]) + 3));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_seg_count.p, 0, 4 * sizeof(uint32_t), st));
    pa.b = ctx->batch; pa.nv = nv; pa.vpos = ctx->var.pos; pa.mode = mode;
    pa.mapping_quality = tag ? t->mapping_quality : p->mapping_quality;
    pa.mapq_filter = tag ? t->mapq_filter : 0; pa.tag_supplementary = tag ? t->tag_supplementary : 0;
    pa.last_var_pos = nv ? ctx->h_vpos[nv - 1] : -1;
    pa.work = ctx->d_work.p; pa.seg_count = ctx->d_seg_count.p;
    pa.tmp_start = ctx->d_tmp_start.p; pa.ncalls = ctx->d_ncalls.p; pa.status = ctx->d_status.p; pa.abort_of_read = ctx->d_abort_of_read.p;
    pa.tag_hp = ctx->d_tag_hp.p; pa.tag_ps = ctx->d_tag_ps.p; pa.tag_pq = ctx->d_tag_pq.p; pa.tag_h1 = ctx->d_tag_h1.p; pa.tag_h2 = ctx->d_tag_h2.p;
    pa.tag_cat = ctx->d_tag_cat.p; pa.tag_h3 = ctx->d_tag_h3.p; pa.tag_end = ctx->d_tag_end.p; pa.tag_len = ctx->d_tag_len.p;
    pa.tag_nps = ctx->d_tag_nps.p; pa.tag_hpb = ctx->d_tag_hpb.p; pa.tag_sim = ctx->d_tag_sim.p;
    if (n > 0) {
        k_prep_reads<<<(n + 255) / 256, 256, 0, st>>>(pa);
        ctx->stats.kernel_launches++;
    }
    // tumor pass: one window-diff work item per (alignment, covered tumor position); sized from the tumor density, re-run on overflow
    size_t wd_cap = 0;
    if (mode == LPS_MODE_EXTRACT_TUMOR) {
        double tden = 0.0;
        if (nv > 1) tden = (double)ctx->som.n_tum / ((double)ctx->h_vpos[nv - 1] - (double)ctx->h_vpos[0] + 1.0);
        wd_cap = (size_t)((double)ctx->sum_l_qseq * tden * 1.5) + (size_t)n + 4096;
        if (wd_cap < ctx->d_wd_items.cap) wd_cap = ctx->d_wd_items.cap;
    }
    // scratch pool capacity from the variant density of the contig; re-run on overflow
    double density = 0.0;
    if (nv > 1) density = (double)nv / ((double)ctx->h_vpos[nv - 1] - (double)ctx->h_vpos[0] + 1.0);
    size_t cap = (size_t)((double)ctx->sum_l_qseq * density * 1.5) + (size_t)n * 4 + 4096 +
                 (size_t)ctx->sm_count * WARPS_PER_CTA * POOL_BLOCK;   // every warp may leave one block partly unused
    if (cap < ctx->d_calls_tmp.cap) cap = ctx->d_calls_tmp.cap;
    size_t clip_cap = (size_t)n * 2 + 1024;
    if (clip_cap < ctx->d_clip_keys.cap) clip_cap = ctx->d_clip_keys.cap;
    size_t ovf_cap = 1u << 14;                                         // reads beyond the shared candidate buffer; grown on demand
    if (ovf_cap < ctx->d_overflow_reads.cap) ovf_cap = ctx->d_overflow_reads.cap;

    CallCounters hc;
    for (int attempt = 0; attempt < 4; attempt++) {
        LPS_CUDA(ctx, ctx->d_calls_tmp.reserve(cap));
        LPS_CUDA(ctx, ctx->d_clip_keys.reserve(clip_cap));
        LPS_CUDA(ctx, ctx->d_clip_meta.reserve(ctx->d_clip_keys.cap));
        LPS_CUDA(ctx, ctx->d_overflow_reads.reserve(ovf_cap));
        LPS_CUDA(ctx, ctx->d_overflow_cand.reserve(ctx->d_overflow_reads.cap));
        LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(CallCounters), st));
        if (som) {
            // per-slot counters start from zero on every attempt (a re-run after a pool overflow must not count twice)
            LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_som_counters.p, 0, ctx->som_counter_words * sizeof(int32_t), st));
            if (ctx->som.n_tum) {
                std::vector<int32_t> init((size_t)ctx->som.n_tum * 2, INT_MIN);
                std::fill(init.begin(), init.begin() + ctx->som.n_tum, INT_MAX);
                LPS_CUDA(ctx, cudaMemcpyAsync(ctx->som.cover_start, init.data(), init.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
                LPS_CUDA(ctx, cudaStreamSynchronize(st));
            }
            if (wd_cap) LPS_CUDA(ctx, ctx->d_wd_items.reserve(wd_cap));
        }
        K1Args a;
        memset(&a, 0, sizeof(a));
        a.b = ctx->batch; a.v = ctx->var;
        if (!tag) {
            a.mapping_quality = p->mapping_quality; a.have_reference = p->have_reference && ctx->ref_len > 0;
            a.apply_filter = p->is_ont;
        } else {
            a.mapping_quality = t->mapping_quality; a.have_reference = t->have_reference && ctx->ref_len > 0;
            a.mapq_filter = t->mapq_filter; a.tag_supplementary = t->tag_supplementary; a.want_calls = want_calls;
            a.percentage = t->percentage_threshold;
            a.hp1_is_alt = ctx->d_vhp1_is_alt.p; a.var_ps = ctx->d_vps.p; a.pq_lut = ctx->d_pq_lut.p;
            a.tag_hp = ctx->d_tag_hp.p; a.tag_ps = ctx->d_tag_ps.p; a.tag_pq = ctx->d_tag_pq.p;
            a.tag_h1 = ctx->d_tag_h1.p; a.tag_h2 = ctx->d_tag_h2.p; a.tag_cat = ctx->d_tag_cat.p;
        }
        if (som) {
            a.som = ctx->som;
            a.wd_items = ctx->d_wd_items.p; a.wd_cap = mode == LPS_MODE_EXTRACT_TUMOR ? ctx->d_wd_items.cap : 0;
            a.tag_h3 = ctx->d_tag_h3.p; a.tag_end = ctx->d_tag_end.p; a.tag_len = ctx->d_tag_len.p; a.tag_nps = ctx->d_tag_nps.p;
            a.tag_hpb = ctx->d_tag_hpb.p; a.tag_sim = ctx->d_tag_sim.p;
        }
        a.last_var_pos = nv ? ctx->h_vpos[nv - 1] : -1;
        a.calls_tmp = ctx->d_calls_tmp.p; a.calls_cap = ctx->d_calls_tmp.cap;
        a.tmp_start = ctx->d_tmp_start.p; a.ncalls = ctx->d_ncalls.p; a.status = ctx->d_status.p;
        a.clip_keys = ctx->d_clip_keys.p; a.clip_cap = ctx->d_clip_keys.cap; a.clip_meta = ctx->d_clip_meta.p;
        a.abort_of_read = ctx->d_abort_of_read.p;
        a.work = ctx->d_work.p; a.seg_count = ctx->d_seg_count.p;
        for (int k = 0; k < 4; k++) a.seg_base[k] = pa.seg_base[k];
        const int max_warps = ctx->sm_count * WARPS_PER_CTA;
        if (getenv("LPS_DEBUG_K1")) {
            LPS_CUDA(ctx, ctx->d_dbg_times.reserve(4 * (size_t)max_warps + 64));
            LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_dbg_times.p, 0, 8 * ctx->d_dbg_times.cap, st));
            a.dbg_times = ctx->d_dbg_times.p;
        }
        a.counters = ctx->d_counters.p;
        a.count_gathers = ctx->zero_copy ? 1 : 0;
        a.overflow_list_out = ctx->d_overflow_reads.p; a.overflow_need_out = ctx->d_overflow_cand.p;
        a.overflow_list_cap = (uint32_t)std::min<size_t>(ctx->d_overflow_reads.cap, 0xFFFFFFFFu);
        int grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        if (grid > ctx->sm_count) grid = ctx->sm_count;   // resident CTAs only (one per SM), reads fetched dynamically
        if (grid > 0) {
            cudaEventRecord(ctx->kev[0], st);
            launch_k1(mode, grid, st, a);
            cudaEventRecord(ctx->kev[1], st);
            ctx->stats.kernel_launches++;
        }
        LPS_CUDA(ctx, cudaGetLastError());
        LPS_CUDA(ctx, cudaMemcpyAsync(&hc, ctx->d_counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        LPS_CUDA(ctx, cudaStreamSynchronize(st));
        if (grid > 0) cudaEventElapsedTime(&ctx->stats.ms_kernel_call_alleles, ctx->kev[0], ctx->kev[1]);
        if (getenv("LPS_DEBUG_K1") && a.dbg_times) {
            const size_t nw = (size_t)grid * WARPS_PER_CTA;
            std::vector<unsigned long long> h(4 * nw);
            cudaMemcpy(h.data(), a.dbg_times, 32 * nw, cudaMemcpyDeviceToHost);
            unsigned long long t0 = ~0ull, t1 = 0;
            for (size_t w = 0; w < nw; w++) { if (h[4 * w] && h[4 * w] < t0) t0 = h[4 * w]; if (h[4 * w + 1] > t1) t1 = h[4 * w + 1]; }
            std::vector<double> st_, en, cnt;
            for (size_t w = 0; w < nw; w++) { st_.push_back((double)(h[4 * w] - t0) / 1e3); en.push_back((double)(h[4 * w + 1] - t0) / 1e3); cnt.push_back((double)h[4 * w + 2]); }
            std::sort(st_.begin(), st_.end()); std::sort(en.begin(), en.end()); std::sort(cnt.begin(), cnt.end());
            auto q = [&](std::vector<double> &v, double f) { return v[(size_t)(f * (v.size() - 1))]; };
            fprintf(stderr, "k1 warps=%zu span=%.1f us | start us p0 %.1f p50 %.1f p99 %.1f max %.1f | end us p0 %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f | reads/warp min %.0f p50 %.0f max %.0f\n",
                    nw, (double)(t1 - t0) / 1e3, q(st_, 0), q(st_, .5), q(st_, .99), q(st_, 1), q(en, 0), q(en, .1), q(en, .5), q(en, .9), q(en, 1), q(cnt, 0), q(cnt, .5), q(cnt, 1));
        }
        if (getenv("LPS_DEBUG_K1"))
            fprintf(stderr, "k1 mode=%d attempt=%d grid=%d ms=%.4f pool=%llu/%zu calls=%llu clips=%llu overflow_reads=%u aborted=%u\n", mode, attempt, grid,
                    ctx->stats.ms_kernel_call_alleles, hc.tmp_calls.v, ctx->d_calls_tmp.cap, hc.n_calls.v, hc.clips.v, (unsigned)hc.overflow_reads.v, (unsigned)hc.aborted_reads.v);
        // zero-copy accounting: one 32-byte sector of SEQ (phase: and one of QUAL) crosses PCIe per gathered candidate
        if (ctx->zero_copy) ctx->stats.h2d_bytes += hc.gathers.v * ((tag || ctx->batch.sq) ? 32ull : 64ull);
        if (hc.bad_cigar.v) return ctx->fail(LPS_E_CIGAR, "alignment find unsupported CIGAR operation");
        if (hc.overflow_reads.v > ctx->d_overflow_reads.cap) {
            // more reads overflow the shared candidate buffer than the list holds (dense variants, long reads): grow the list and redo
            // the attempt, like the call pool
            if (attempt == 3) return ctx->fail(LPS_E_NOMEM, "overflow list sizing did not converge");
            ovf_cap = (size_t)hc.overflow_reads.v + (size_t)hc.overflow_reads.v / 8 + 1024;
            if (hc.tmp_calls.v > ctx->d_calls_tmp.cap) cap = (size_t)hc.tmp_calls.v + (size_t)hc.tmp_calls.v / 8 + 65536;
            continue;
        }
        if (hc.overflow_reads.v) {
            // second pass for the reads with more candidates than the shared buffer: same kernel, candidate lists in global memory
            // sized from the first pass
            std::vector<uint64_t> need(hc.overflow_reads.v), off(hc.overflow_reads.v + 1, 0);
            LPS_CUDA(ctx, cudaMemcpy(need.data(), ctx->d_overflow_cand.p, 8 * (size_t)hc.overflow_reads.v, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < hc.overflow_reads.v; i++) off[i + 1] = off[i] + need[i] * (som ? 2 : 1);   // 8-byte units
            LPS_CUDA(ctx, ctx->d_overflow_off.reserve(off.size()));
            LPS_CUDA(ctx, cudaMemcpy(ctx->d_overflow_off.p, off.data(), 8 * off.size(), cudaMemcpyHostToDevice));
            // persistent scratch of the context: the dense configs (C5) take this path on every call
            LPS_CUDA(ctx, ctx->d_ovf_cand.reserve(8 * ((size_t)off.back() + 2)));
            K1Args b2 = a;
            b2.overflow_reads = ctx->d_overflow_reads.p; b2.overflow_off = ctx->d_overflow_off.p; b2.overflow_buf = reinterpret_cast<Cand *>(ctx->d_ovf_cand.p);
            b2.overflow_list_cap = (uint32_t)hc.overflow_reads.v;
            // the overflow pass must not append the clips / counters of these reads a second time
            b2.clip_cap = 0;
            LPS_CUDA(ctx, ctx->d_counters2.reserve(1));
            LPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_counters2.p, ctx->d_counters.p, sizeof(CallCounters), cudaMemcpyDeviceToDevice, st));
            b2.counters = ctx->d_counters2.p;
            const int g2 = ((int)hc.overflow_reads.v + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
            launch_k1(mode, g2, st, b2);
            ctx->stats.kernel_launches++;
            LPS_CUDA(ctx, cudaGetLastError());
            CallCounters h2;
            LPS_CUDA(ctx, cudaMemcpyAsync(&h2, ctx->d_counters2.p, sizeof(h2), cudaMemcpyDeviceToHost, st));
            LPS_CUDA(ctx, cudaStreamSynchronize(st));
            hc.tmp_calls.v = h2.tmp_calls.v; hc.n_calls.v = h2.n_calls.v; hc.wd_items.v = h2.wd_items.v; hc.aborted_reads.v = h2.aborted_reads.v;
        }
        ctx->n_wd_items = hc.wd_items.v;
        if (hc.tmp_calls.v <= ctx->d_calls_tmp.cap && hc.clips.v <= ctx->d_clip_keys.cap && hc.wd_items.v <= ctx->d_wd_items.cap) break;
        cap = (size_t)hc.tmp_calls.v + (size_t)hc.tmp_calls.v / 8 + 65536;   // the block hand-out is not deterministic: leave slack
        clip_cap = (size_t)hc.clips.v + 1024;
        if (wd_cap) wd_cap = (size_t)hc.wd_items.v + 1024;
        if (attempt == 3) return ctx->fail(LPS_E_NOMEM, "call pool sizing did not converge");
    }


    // CSR offsets in read order
    const int tb = 256;
    k_widen_u32<<<(n + tb) / tb, tb, 0, st>>>(n, ctx->d_ncalls.p, ctx->d_call_off.p + 0);
    ctx->stats.kernel_launches++;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->d_call_off.p, ctx->d_call_off.p, n + 1, st);
    LPS_CUDA(ctx, ctx->d_cub_tmp.reserve(tmp_bytes));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_call_off.p + n, 0, 8, st));
    cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, tmp_bytes, ctx->d_call_off.p, ctx->d_call_off.p, n + 1, st);
    ctx->stats.kernel_launches++;
    ctx->n_calls = hc.n_calls.v;
    LPS_CUDA(ctx, ctx->d_calls.reserve((size_t)ctx->n_calls + 1));
    if (n > 0) {
        const long long threads = (long long)n * 16;
        k_gather_calls<<<(unsigned)((threads + tb - 1) / tb), tb, 0, st>>>(n, ctx->d_tmp_start.p, ctx->d_ncalls.p, ctx->d_call_off.p,
                                                                         ctx->d_calls_tmp.p, ctx->d_calls.p);
        ctx->stats.kernel_launches++;
    }
    LPS_CUDA(ctx, cudaGetLastError());

    // clipCount map: sort the (pos << 1 | side) keys, run-length encode
    const int nclip = (int)hc.clips.v;
    if (nclip > 0 && hc.aborted_reads.v) {
        k_clip_filter<<<(nclip + 255) / 256, 256, 0, st>>>((unsigned long long)nclip, ctx->d_clip_keys.p, ctx->d_clip_meta.p, ctx->d_abort_of_read.p);
        ctx->stats.kernel_launches++;
    }
    // the sorted, run-length encoded keys travel to pinned memory asynchronously; lps_finish_clips turns them into the map
    ctx->n_clip_events = nclip;
    ctx->clips_pending = true;
    if (nclip > 0) {
        LPS_CUDA(ctx, ctx->d_clip_keys_sorted.reserve((size_t)nclip));
        LPS_CUDA(ctx, ctx->d_clip_unique.reserve((size_t)nclip));
        LPS_CUDA(ctx, ctx->d_clip_counts.reserve((size_t)nclip));
        LPS_CUDA(ctx, ctx->d_num_runs.reserve(1));
        LPS_CUDA(ctx, ctx->p_clip_unique.reserve((size_t)nclip));
        LPS_CUDA(ctx, ctx->p_clip_counts.reserve((size_t)nclip));
        LPS_CUDA(ctx, ctx->p_num_runs.reserve(1));
        // off the critical path: nothing on the device waits for the clip map (the host turns it into CNV intervals), so it is sorted,
        // run-length encoded and copied on the context's side stream while the main stream goes on with the graph
        cudaStream_t sd = ctx->stream_side;
        LPS_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
        LPS_CUDA(ctx, cudaStreamWaitEvent(sd, ctx->ev_fork, 0));
        const int key_bits = 32;
        size_t b1 = 0, b2 = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, b1, ctx->d_clip_keys.p, ctx->d_clip_keys_sorted.p, nclip, 0, key_bits, sd);
        cub::DeviceRunLengthEncode::Encode(nullptr, b2, ctx->d_clip_keys_sorted.p, ctx->d_clip_unique.p, ctx->d_clip_counts.p,
                                           ctx->d_num_runs.p, nclip, sd);
        LPS_CUDA(ctx, ctx->d_cub_tmp_side.reserve(b1 > b2 ? b1 : b2));
        cub::DeviceRadixSort::SortKeys(ctx->d_cub_tmp_side.p, b1, ctx->d_clip_keys.p, ctx->d_clip_keys_sorted.p, nclip, 0, key_bits, sd);
        cub::DeviceRunLengthEncode::Encode(ctx->d_cub_tmp_side.p, b2, ctx->d_clip_keys_sorted.p, ctx->d_clip_unique.p, ctx->d_clip_counts.p,
                                           ctx->d_num_runs.p, nclip, sd);
        ctx->stats.kernel_launches += 2;
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_num_runs.p, ctx->d_num_runs.p, 4, cudaMemcpyDeviceToHost, sd));
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_clip_unique.p, ctx->d_clip_unique.p, 4 * (size_t)nclip, cudaMemcpyDeviceToHost, sd));
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_clip_counts.p, ctx->d_clip_counts.p, 4 * (size_t)nclip, cudaMemcpyDeviceToHost, sd));
        LPS_CUDA(ctx, cudaEventRecord(ctx->ev_clips, sd));
    }
    ctx->have_calls = !tag;
    ctx->host_calls_valid = false;
    if (!defer_clips) return lps_finish_clips(ctx);
    return LPS_OK;
}

int lps_finish_clips(lps_ctx *ctx) {
    if (!ctx->clips_pending) return LPS_OK;
    ctx->clips_pending = false;
    ctx->h_clip_pos.clear(); ctx->h_clip_front.clear(); ctx->h_clip_back.clear();
    if (ctx->n_clip_events <= 0) return LPS_OK;
    LPS_CUDA(ctx, cudaEventSynchronize(ctx->ev_clips));
    const int runs = ctx->p_num_runs.p[0];
    const uint32_t *keys = ctx->p_clip_unique.p, *cnts = ctx->p_clip_counts.p;
    ctx->stats.d2h_bytes += 8ull * (uint64_t)runs;
    for (int i = 0; i < runs; i++) {
        if (keys[i] == 0xFFFFFFFFu) continue;   // cancelled by k_clip_filter
        int32_t pos = (int32_t)(keys[i] >> 1);
        if (ctx->h_clip_pos.empty() || ctx->h_clip_pos.back() != pos) {
            ctx->h_clip_pos.push_back(pos); ctx->h_clip_front.push_back(0); ctx->h_clip_back.push_back(0);
        }
        if (keys[i] & 1u) ctx->h_clip_back.back() += (int32_t)cnts[i]; else ctx->h_clip_front.back() += (int32_t)cnts[i];
    }
    return LPS_OK;
}
