/*
 * oracle_somatic.c — CPU restatement of the somatic family of the tag dialect (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 *   CIGAR walk + IsAltIndel       CigarParser::parsingCigar              src/haplotag/HaplotagParsingBam.cpp:541-670
 *   countBaseNucleotide           CigarParser::countBaseNucleotide       src/haplotag/HaplotagParsingBam.cpp:682-730
 *   normal pass                   ExtractNorDataChrProcessor / CigarParser  src/somatic_haplotag/SomaticVarCaller.cpp:123-293
 *   tumor pass                    ExtractTumDataChrProcessor / CigarParser  src/somatic_haplotag/SomaticVarCaller.cpp:334-518, 712-759
 *   window diff                   getWindowsDiffRef / getOrderWindowsDiffRef / processCigarOperation  :627-710
 *   somatic tagging               SomaticHaplotagChrProcessor::judgeHaplotype / inheritHaplotype, SomaticHaplotagCigarParser
 *                                                                        src/somatic_haplotag/SomaticHaplotagProcess.cpp:310-579
 *   strategies                    judgeSomaticSnpHap / judgeNormalSnpHap / judgeSomaticReadHap / judgeTumorOnlySnpHap
 *                                                                        src/haplotag/HaplotagStrategy.cpp:315-668
 *   germline votes                GermlineHaplotagStrategy::judgeSnpHap / judgeDeletionHap / judgeReadHap  :20-300
 * Parity status: pinned against the unmodified reference through oracle/ref_tap_somatic.cpp.
 *
 * Reads whose CIGAR addresses a base beyond l_qseq are undefined in the reference (it reads past SEQ); such a hit is
 * dropped here, as in oracle_tag.c.
 */
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static const char NT16[] = "=ACMGRSVTWYHKDBN";

enum { VT_NONE = 0, VT_SNP = 1, VT_INS = 2, VT_DEL = 3, VT_MNP = 4 };   /* HaplotagVariantType (HaplotagType.h:76-85) */
enum { GT_PHASED_HETERO = 1 };
enum { HP_UNTAG = 0, HP_H1, HP_H2, HP_H3, HP_H4, HP_H1_1, HP_H1_2, HP_H2_1, HP_H2_2 };

static int var_type(int rl, int al) {
    if (rl == 1 && al == 1) return VT_SNP;
    if (rl == 1 && al > 1) return VT_INS;
    if (rl > 1 && al == 1) return VT_DEL;
    if (rl > 1 && rl == al) return VT_MNP;
    return VT_NONE;   /* VarData::setVariantType throws for anything else (HaplotagType.h:129-141) */
}

typedef struct { lps_call *a; uint64_t n, cap; } callvec;
static void cv_push(callvec *c, lps_call x) {
    if (c->n == c->cap) { c->cap = c->cap * 2 + 1024; c->a = (lps_call *)realloc(c->a, sizeof(lps_call) * c->cap); }
    c->a[c->n++] = x;
}

typedef struct {
    const lps_read_batch *b; const lps_variants *v; const lps_tumor_variants *t; const uint8_t *hom;
    const char *ref; int64_t ref_len;
} som_in;

static inline int nor_present(const som_in *s, int vi) { return s->t->nor_present ? s->t->nor_present[vi] != 0 : 1; }
static inline int nor_gt(const som_in *s, int vi) { return s->v->gt_kind ? s->v->gt_kind[vi] : GT_PHASED_HETERO; }

/* per-read scratch: the variants that were touched, in walk order */
typedef struct { int32_t var; int8_t vhp; uint8_t flags; } touch;   /* flags: 1 tumorSnpPosVec / tumVarPosVec, 2 tumorAllelePosVec, 4 somatic variant */
typedef struct { touch *a; int n, cap; } touchvec;
static touch *tv_get(touchvec *t, int var) {
    if (t->n && t->a[t->n - 1].var == var) return &t->a[t->n - 1];
    if (t->n == t->cap) { t->cap = t->cap * 2 + 64; t->a = (touch *)realloc(t->a, sizeof(touch) * (size_t)t->cap); }
    touch x = {var, 0, 0};
    t->a[t->n++] = x;
    return &t->a[t->n - 1];
}

typedef struct { int h1, h2, h3; int ps_seen, ps_min, ps_multi; } read_acc;
static void count_ps(read_acc *a, int ps) {
    if (!a->ps_seen) { a->ps_seen = 1; a->ps_min = ps; }
    else { if (ps != a->ps_min) a->ps_multi = 1; if (ps < a->ps_min) a->ps_min = ps; }
}

/* CigarParser::countBaseNucleotide (HaplotagParsingBam.cpp:682-720) */
static void count_base(int32_t *pb, char base, int mpq_ok, int is_alt, int ttype) {
    int k = base == 'A' ? 0 : base == 'C' ? 1 : base == 'G' ? 2 : base == 'T' ? 3 : 4;
    if (mpq_ok) { pb[LPS_PB_MPQ_A + k]++; if (is_alt) pb[LPS_PB_MPQ_ALT]++; pb[LPS_PB_MPQ_DEPTH]++; }
    pb[LPS_PB_A + k]++;
    if (is_alt) { if (ttype == VT_DEL) pb[LPS_PB_DEL]++; pb[LPS_PB_ALT]++; }
    pb[LPS_PB_DEPTH]++;
}

/* GermlineHaplotagStrategy::judgeSnpHap (HaplotagStrategy.cpp:20-130) on the NORMAL record: 0/1 in touch.vhp */
static void germline_match_vote(const som_in *s, int vi, char base, int has_next, int ends_next_i, int ends_next_d, read_acc *acc,
                                touchvec *tv) {
    const lps_variants *v = s->v;
    int rl = v->ref_len[vi], al = v->alt_len[vi], ty = var_type(rl, al), h1alt = v->hp1_is_alt[vi] != 0;
    if (ty == VT_SNP) {
        char rb = (char)v->ref0[vi], ab = (char)v->alt0[vi];
        if (base == rb || base == ab) {
            if (base == (h1alt ? ab : rb)) { acc->h1++; tv_get(tv, vi)->vhp = 0; }
            if (base == (h1alt ? rb : ab)) { acc->h2++; tv_get(tv, vi)->vhp = 1; }
            count_ps(acc, v->ps[vi]);
        }
    } else if ((ty == VT_INS || ty == VT_DEL) && has_next) {
        int has = ty == VT_INS ? ends_next_i : ends_next_d;
        int l1 = h1alt ? al : rl, l2 = h1alt ? rl : al, hpbit = -1;
        if (l1 != 1 && l2 == 1) hpbit = has ? 0 : 1;
        else if (l1 == 1 && l2 != 1) hpbit = has ? 1 : 0;
        if (hpbit == 0) { acc->h1++; tv_get(tv, vi)->vhp = 0; }
        if (hpbit == 1) { acc->h2++; tv_get(tv, vi)->vhp = 1; }
        count_ps(acc, v->ps[vi]);
    }
}

/* GermlineHaplotagStrategy::judgeDeletionHap (HaplotagStrategy.cpp:147-209) */
static void germline_del_vote(const som_in *s, int vi, int ref_pos, int len, int qpos, int lq, const uint8_t *seq, int have_ref,
                              read_acc *acc, touchvec *tv) {
    const lps_variants *v = s->v;
    int vp = v->pos[vi];
    if (!have_ref || ref_pos + len + 1 == vp || !(vp >= ref_pos && vp < ref_pos + len) || s->hom[vi] < 3) return;
    int rl = v->ref_len[vi], al = v->alt_len[vi], ty = var_type(rl, al), h1alt = v->hp1_is_alt[vi] != 0;
    if (ty == VT_SNP) {
        if (qpos >= lq) return;   /* undefined in the reference */
        char c = NT16[(seq[qpos >> 1] >> ((~qpos & 1) << 2)) & 0xf];
        char rb = (char)v->ref0[vi], ab = (char)v->alt0[vi];
        if (c == (h1alt ? ab : rb)) { acc->h1++; tv_get(tv, vi)->vhp = 0; }
        if (c == (h1alt ? rb : ab)) { acc->h2++; tv_get(tv, vi)->vhp = 1; }
        count_ps(acc, v->ps[vi]);
    } else if (ty == VT_DEL) {
        int l1 = h1alt ? al : rl, l2 = h1alt ? rl : al;
        if (l1 != 1 && l2 == 1) { acc->h1++; tv_get(tv, vi)->vhp = 0; }
        else if (l1 == 1 && l2 != 1) { acc->h2++; tv_get(tv, vi)->vhp = 1; }
        count_ps(acc, v->ps[vi]);
    }
}

/* SomaticJudgeHapStrategy::judgeSomaticSnpHap (HaplotagStrategy.cpp:315-389) with judgeNormalSnpHap (:403-437) and the
 * two judgeTumorOnlySnpHap (:617-638 extract, :653-668 tagging).  touch.vhp takes SnpHP values 1/2/3.              */
static void somatic_match_vote(const som_in *s, int vi, char base, int is_alt, int tagging, read_acc *acc, touchvec *tv) {
    const lps_variants *v = s->v; const lps_tumor_variants *t = s->t;
    if (nor_present(s, vi)) {
        if (nor_gt(s, vi) != GT_PHASED_HETERO) return;
        int rl = v->ref_len[vi], al = v->alt_len[vi], ty = var_type(rl, al), h1alt = v->hp1_is_alt[vi] != 0;
        if (ty == VT_INS || ty == VT_DEL) {
            /* base := the whole ALT / REF string, compared with the HP1 / HP2 strings */
            if (is_alt == h1alt) { acc->h1++; tv_get(tv, vi)->vhp = 1; } else { acc->h2++; tv_get(tv, vi)->vhp = 2; }
            count_ps(acc, v->ps[vi]);
        } else if (ty == VT_SNP) {
            char rb = (char)v->ref0[vi], ab = (char)v->alt0[vi];
            if (base == rb || base == ab) {
                if (base == (h1alt ? ab : rb)) { acc->h1++; tv_get(tv, vi)->vhp = 1; }
                if (base == (h1alt ? rb : ab)) { acc->h2++; tv_get(tv, vi)->vhp = 2; }
                count_ps(acc, v->ps[vi]);
            }
        }
        /* MNP: a one-character base never equals a longer REF/ALT string */
    } else if (t->tum_present[vi]) {
        int gt = t->gt_kind[vi];
        if (gt < 1 || gt > 3) return;
        int rl = t->ref_len[vi], al = t->alt_len[vi], ty = var_type(rl, al);
        int indel = ty == VT_INS || ty == VT_DEL;
        int snp_hit = ty == VT_SNP && (base == (char)t->ref0[vi] || base == (char)t->alt0[vi]);
        if (!(snp_hit || indel)) return;
        int base_is_alt = indel ? is_alt : base == (char)t->alt0[vi];
        if (!tagging) {
            if (base_is_alt) { acc->h3++; touch *x = tv_get(tv, vi); x->vhp = 3; x->flags |= 2; }
        } else if (t->is_somatic[vi]) {
            if (base_is_alt) { acc->h3++; tv_get(tv, vi)->vhp = 3; }
        }
        /* tumCountPS only feeds the read log */
    }
}

/* SomaticJudgeHapStrategy::judgeSomaticReadHap (HaplotagStrategy.cpp:452-602); hpCount[4] is always 0 */
static int somatic_read_hap(const read_acc *a, double pct, int *pq) {
    double tmin = 0, tmax = 0, nmin, nmax; int max_n;
    if (a->h3 > 0) { tmax = a->h3; tmin = 0; } else { tmax = 0; tmin = a->h3; }   /* h3 > h4(=0) ? ... */
    if (a->h1 > a->h2) { nmin = a->h2; nmax = a->h1; max_n = 1; } else { nmin = a->h1; nmax = a->h2; max_n = 2; }
    double tsim = tmax == 0 ? 0.0 : tmax / (tmax + tmin);
    double nsim = nmax == 0 ? 0.0 : nmax / (nmax + nmin);
    int hp = HP_UNTAG;
    *pq = 0;
    if (tmax != 0) {
        if (tsim >= pct) {
            if (nsim >= pct) hp = max_n == 1 ? HP_H1_1 : HP_H2_1;
            else hp = HP_H3;
        }
    } else if (nmax != 0) {
        if (nsim >= pct) hp = max_n;
    }
    if (a->ps_multi) hp = HP_UNTAG;
    if (nmax == 0 && tmax == 0) *pq = 0;
    else if (tmax != 0) *pq = tmax == tmax + tmin ? 40 : (int)(-10 * (log10((double)tmin / (double)(tmax + tmin))));
    else *pq = nmax == nmax + nmin ? 40 : (int)(-10 * (log10((double)nmin / (double)(nmax + nmin))));
    return hp;
}

/* processCigarOperation (SomaticVarCaller.cpp:627-654) */
static int wd_next_op(const uint32_t *cig, int *ci, int ci_end, int dir, int *remaining, int *read_pos, int *ref_pos, int *op) {
    *ci += dir;
    while (*ci < ci_end && *ci >= 0) {
        *op = (int)(cig[*ci] & 15);
        int len = (int)(cig[*ci] >> 4);
        if (*op == 0 || *op == 3 || *op == 6 || *op == 7 || *op == 8) { *remaining += len; return 1; }
        else if (*op == 1) *read_pos += len * dir;
        else if (*op == 2) *ref_pos += len * dir;
        else return 0;
        *ci += dir;
    }
    return 0;
}

/* getOrderWindowsDiffRef (:655-686): bins the offsets instead of listing (offset, base) pairs */
static void wd_scan(const som_in *s, const uint32_t *cig, int ci, int ncig, const uint8_t *seq, int lq, int read_pos, int remaining,
                    int ref_pos, int dir, int32_t *hist) {
    int op = (int)(cig[ci] & 15);
    for (int i = 1; i <= LPS_WINDOW; i++) {
        remaining--;
        if (remaining == 0 || remaining == -1)
            if (!wd_next_op(cig, &ci, ncig, dir, &remaining, &read_pos, &ref_pos, &op)) return;
        if (op == 2 || op == 1 || op == 3 || op == 6 || op == 8) continue;
        read_pos += dir; ref_pos += dir;
        if (read_pos > lq || (int64_t)ref_pos > s->ref_len || read_pos < 0 || ref_pos < 0) return;
        if (read_pos == lq) return;   /* one past SEQ: undefined in the reference */
        char rb = NT16[(seq[read_pos >> 1] >> ((~read_pos & 1) << 2)) & 0xf];
        char fb = (int64_t)ref_pos == s->ref_len ? '\0' : s->ref[ref_pos];   /* std::string::operator[](size()) is '\0' */
        if (rb != fb) hist[i * dir + LPS_WINDOW]++;
    }
}

/* getWindowsDiffRef (:688-710) */
static void window_diff(const som_in *s, const uint32_t *cig, int ci, int ncig, const uint8_t *seq, int lq, int query_pos, int offset,
                        int var_pos, int32_t *hist) {
    int oplen = (int)(cig[ci] >> 4), op = (int)(cig[ci] & 15);
    int fwd = 0, rev = 0;
    if (op != 1) { fwd = oplen - offset > 0 ? oplen - offset : 0; rev = offset > 0 ? offset : 0; }
    wd_scan(s, cig, ci, ncig, seq, lq, query_pos + offset, rev, var_pos, -1, hist);
    wd_scan(s, cig, ci, ncig, seq, lq, query_pos + offset, fwd, var_pos, 1, hist);
}

static void fill_slots(const lps_tumor_variants *t, int nv, int32_t **slot_of_var, int32_t **tum_var, int *n_tum) {
    *slot_of_var = (int32_t *)malloc(sizeof(int32_t) * ((size_t)nv + 1));
    int k = 0;
    for (int i = 0; i < nv; i++) (*slot_of_var)[i] = t->tum_present[i] ? k++ : -1;
    *n_tum = k;
    *tum_var = (int32_t *)malloc(sizeof(int32_t) * ((size_t)k + 1));
    for (int i = 0; i < nv; i++) if (t->tum_present[i]) (*tum_var)[(*slot_of_var)[i]] = i;
}

int orc_somatic(int mode, const lps_read_batch *b, const lps_variants *v, const lps_tumor_variants *t, const uint8_t *hom, const char *ref,
                int64_t ref_len, const lps_tag_params *p, orc_somatic_out *out) {
    memset(out, 0, sizeof(*out));
    if (!p->have_reference) ref_len = 0;   /* ref_string == "" */
    som_in s = {b, v, t, hom, ref, ref_len};
    const int n = b->n_reads, nv = v->n;
    int32_t *slot;
    fill_slots(t, nv, &slot, &out->tum_var, &out->n_tum);
    const size_t nt = (size_t)out->n_tum + 1;
    out->n_reads = n;
    out->category = (uint8_t *)calloc((size_t)n + 1, 1);
    out->read_hp = (int8_t *)calloc((size_t)n + 1, 1); out->hp_before = (int8_t *)calloc((size_t)n + 1, 1);
    out->ps = (int32_t *)calloc((size_t)n + 1, 4); out->pq = (int32_t *)calloc((size_t)n + 1, 4);
    out->h1 = (int32_t *)calloc((size_t)n + 1, 4); out->h2 = (int32_t *)calloc((size_t)n + 1, 4); out->h3 = (int32_t *)calloc((size_t)n + 1, 4);
    out->n_ps = (uint8_t *)calloc((size_t)n + 1, 1);
    out->end_pos = (int32_t *)calloc((size_t)n + 1, 4); out->read_len = (int32_t *)calloc((size_t)n + 1, 4);
    out->derive_similarity = (float *)calloc((size_t)n + 1, 4);
    out->pos_base = (int32_t *)calloc(nt * LPS_PB_FIELDS, 4);
    out->read_hp_count = (int32_t *)calloc(nt * 9, 4);
    out->somatic_read_hp_count = (int32_t *)calloc(nt * 9, 4);
    out->case_count = (int32_t *)calloc(nt * LPS_CASE_FIELDS, 4);
    out->allele_count = (int32_t *)calloc(nt * 2, 4);
    out->window_hist = (int32_t *)calloc(nt * 2 * LPS_WINDOW_BINS, 4);
    out->hp_before_count = (int32_t *)calloc(nt * 9, 4); out->hp_after_count = (int32_t *)calloc(nt * 9, 4);
    out->h3_before_count = (int32_t *)calloc(nt * 9, 4); out->h3_after_count = (int32_t *)calloc(nt * 9, 4);
    out->cover_start = (int32_t *)malloc(nt * 4); out->cover_end = (int32_t *)malloc(nt * 4);
    out->ratios_f = (float *)calloc(nt * LPS_RF_FIELDS, 4); out->ratios_d = (double *)calloc(nt * LPS_RD_FIELDS, 8);
    out->case_read_count = (int32_t *)calloc(nt, 4);
    for (size_t i = 0; i < nt; i++) { out->cover_start[i] = INT_MAX; out->cover_end[i] = INT_MIN; }
    out->call_off = (uint64_t *)calloc((size_t)n + 2, 8);
    callvec cv = {0, 0, 0};
    touchvec tv = {0, 0, 0};
    const int last_pos = nv ? v->pos[nv - 1] : -1;
    int rc = 0;
    for (int r = 0; r < n && !rc; r++) {
        out->call_off[r] = cv.n;
        const int flag = b->flag[r];
        int cat = LPS_TAG_PROCESSED;
        if ((int)b->mapq[r] < p->mapping_quality && p->mapq_filter) cat = LPS_TAG_LOW_MAPQ;
        else if (flag & 0x4) cat = LPS_TAG_UNMAPPED;
        else if (flag & 0x100) cat = LPS_TAG_SECONDARY;
        else if ((flag & 0x800) && !p->tag_supplementary) cat = LPS_TAG_SUPPLEMENTARY;
        else if (nv == 0) cat = LPS_TAG_EMPTY_VARIANTS;
        else if (!(b->ref_start[r] <= last_pos)) cat = LPS_TAG_OTHER;
        out->category[r] = (uint8_t)cat;
        if (cat != LPS_TAG_PROCESSED) continue;

        const int mpq_ok = (int)b->mapq[r] >= p->mapping_quality;
        const int have_ref = p->have_reference && ref_len > 0;
        read_acc acc = {0, 0, 0, 0, 0, 0};
        tv.n = 0;
        int ref_pos = b->ref_start[r], qpos = 0;
        const int lq = b->l_qseq[r];
        int lo = 0, hi = nv;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (v->pos[mid] < ref_pos) lo = mid + 1; else hi = mid; }
        int cur = lo;
        const uint32_t *cig = b->cigar + b->cigar_off[r];
        const uint8_t *seq = b->seq4 + b->seq_off[r];
        int ncig = (int)b->n_cigar[r];
        if (cur == nv) ncig = 0;
        for (int i = 0; i < ncig; i++) {
            const int op = (int)(cig[i] & 15), len = (int)(cig[i] >> 4);
            while (cur < nv && v->pos[cur] < ref_pos) cur++;
            if (op == 0 || op == 7 || op == 8) {
                for (; cur < nv && v->pos[cur] < ref_pos + len; cur++) {
                    const int vp = v->pos[cur], off = vp - ref_pos;
                    if (qpos + off >= lq) continue;   /* undefined in the reference */
                    const char base = NT16[(seq[(qpos + off) >> 1] >> ((~(qpos + off) & 1) << 2)) & 0xf];
                    const int N = nor_present(&s, cur), T = t->tum_present[cur] != 0;
                    const int has_next = i + 1 < ncig;
                    const int ends = ref_pos + len - 1 == vp;
                    const int nop = has_next ? (int)(cig[i + 1] & 15) : -1;
                    /* IsAltIndel (HaplotagParsingBam.cpp:650-670) on NORMAL if present, else TUMOR */
                    int is_alt = 0;
                    if (N || T) {
                        const int rl = N ? v->ref_len[cur] : t->ref_len[cur], al = N ? v->alt_len[cur] : t->alt_len[cur];
                        const char ab = (char)(N ? v->alt0[cur] : t->alt0[cur]);
                        const int ty = var_type(rl, al);
                        if (ty == VT_SNP) is_alt = base == ab;
                        else if (ty == VT_INS && has_next) is_alt = ends && nop == 1;
                        else if (ty == VT_DEL && has_next) is_alt = ends && nop == 2;
                    }
                    const int ttype = T ? var_type(t->ref_len[cur], t->alt_len[cur]) : VT_NONE;
                    int32_t *pb = T ? out->pos_base + (size_t)slot[cur] * LPS_PB_FIELDS : NULL;
                    if (mode == ORC_SOM_EXTRACT_NORMAL) {
                        if (T && (ttype == VT_SNP || ttype == VT_INS || ttype == VT_DEL)) {
                            tv_get(&tv, cur)->flags |= 1;
                            count_base(pb, base, mpq_ok, is_alt, ttype);
                        }
                        if (mpq_ok && N && nor_gt(&s, cur) == GT_PHASED_HETERO)
                            germline_match_vote(&s, cur, base, has_next, ends && nop == 1, ends && nop == 2, &acc, &tv);
                    } else if (mode == ORC_SOM_EXTRACT_TUMOR) {
                        if (mpq_ok) {
                            somatic_match_vote(&s, cur, base, is_alt, 0, &acc, &tv);
                            if (T) tv_get(&tv, cur)->flags |= 1;
                        }
                        if (T && (ttype == VT_SNP || ttype == VT_INS || ttype == VT_DEL)) {
                            if (ttype != VT_SNP || base == (char)t->ref0[cur] || base == (char)t->alt0[cur]) {
                                out->allele_count[(size_t)slot[cur] * 2 + is_alt]++;
                                window_diff(&s, cig, i, (int)b->n_cigar[r], seq, lq, qpos, off, vp,
                                            out->window_hist + ((size_t)slot[cur] * 2 + is_alt) * LPS_WINDOW_BINS);
                                out->n_window_items++;
                            }
                            count_base(pb, base, mpq_ok, is_alt, ttype);
                        }
                    } else {
                        somatic_match_vote(&s, cur, base, is_alt, 1, &acc, &tv);
                        if (t->is_somatic[cur]) tv_get(&tv, cur)->flags |= 4;   /* somaticVarDeriveHP entry */
                    }
                }
                qpos += len; ref_pos += len;
            } else if (op == 1) qpos += len;
            else if (op == 2) {
                int judged = 0;
                for (; cur < nv && v->pos[cur] < ref_pos + len; cur++) {
                    const int T = t->tum_present[cur] != 0, N = nor_present(&s, cur);
                    const int ttype = T ? var_type(t->ref_len[cur], t->alt_len[cur]) : VT_NONE;
                    if (mode == ORC_SOM_EXTRACT_NORMAL || mode == ORC_SOM_EXTRACT_TUMOR) {
                        if (T) {
                            int32_t *pb = out->pos_base + (size_t)slot[cur] * LPS_PB_FIELDS;
                            if (mode == ORC_SOM_EXTRACT_NORMAL) tv_get(&tv, cur)->flags |= 1;
                            if (ttype == VT_SNP) { pb[LPS_PB_DEL]++; pb[LPS_PB_DEPTH]++; }
                            else if (ttype == VT_DEL) { pb[LPS_PB_ALT]++; pb[LPS_PB_DEL]++; pb[LPS_PB_DEPTH]++; }
                        }
                        if (mode == ORC_SOM_EXTRACT_NORMAL && mpq_ok && N && !judged && nor_gt(&s, cur) == GT_PHASED_HETERO) {
                            judged = 1;
                            germline_del_vote(&s, cur, ref_pos, len, qpos, lq, seq, have_ref, &acc, &tv);
                        }
                    }
                    /* somatic tagging: recordDelReadCount only feeds the benchmark */
                }
                ref_pos += len;
            } else if (op == 3) ref_pos += len;
            else if (op == 4) qpos += len;
            else if (op == 5 || op == 6) {}
            else { rc = LPS_E_CIGAR; break; }
        }
        out->h1[r] = acc.h1; out->h2[r] = acc.h2; out->h3[r] = acc.h3;
        out->n_ps[r] = (uint8_t)(acc.ps_multi ? 2 : acc.ps_seen);
        out->end_pos[r] = ref_pos; out->read_len[r] = qpos;
        int hp, pq = 0;
        if (mode == ORC_SOM_EXTRACT_NORMAL) {
            /* GermlineHaplotagStrategy::judgeReadHap (HaplotagStrategy.cpp:243-300) */
            double mx = acc.h1 > acc.h2 ? acc.h1 : acc.h2, mn = acc.h1 > acc.h2 ? acc.h2 : acc.h1;
            hp = 0;
            if (!(mx / (mx + mn) < p->percentage_threshold)) { if (acc.h1 > acc.h2) hp = 1; if (acc.h1 < acc.h2) hp = 2; }
            if (mx == 0) pq = 0; else if (mx == (mx + mn)) pq = 40; else pq = (int)(-10 * (log10((double)mn / (double)(mx + mn))));
            if (acc.ps_multi) hp = 0;
            for (int k = 0; k < tv.n; k++)
                if (tv.a[k].flags & 1) out->read_hp_count[(size_t)slot[tv.a[k].var] * 9 + hp]++;
            out->ps[r] = hp ? acc.ps_min : 0;
        } else if (mode == ORC_SOM_EXTRACT_TUMOR) {
            hp = somatic_read_hap(&acc, p->percentage_threshold, &pq);
            /* classifyReadsByCase (SomaticVarCaller.cpp:462-518) + somaticReadHpCount (:395-404) */
            const int record = !acc.ps_multi, clean = acc.h1 == 0 || acc.h2 == 0;
            for (int k = 0; k < tv.n; k++) {
                const touch *x = &tv.a[k];
                const size_t sl = (size_t)(slot[x->var] < 0 ? 0 : slot[x->var]);
                if (x->flags & 2) {
                    int32_t *cc = out->case_count + sl * LPS_CASE_FIELDS;
                    if (!record) cc[LPS_CASE_UNTAG]++;
                    else if (clean) {
                        cc[LPS_CASE_CLEAN_HP3]++;
                        if (acc.h1 == 0 && acc.h2 == 0) cc[LPS_CASE_PURE_H3]++;
                        else if (acc.h1 != 0 && acc.h2 == 0) cc[LPS_CASE_PURE_H1_1]++;
                        else if (acc.h1 == 0 && acc.h2 != 0) cc[LPS_CASE_PURE_H2_1]++;
                    } else cc[LPS_CASE_MIXED]++;
                    if (hp == HP_H1_1 || hp == HP_H2_1 || hp == HP_H3 || hp == HP_UNTAG) out->somatic_read_hp_count[sl * 9 + hp]++;
                }
                if (x->flags & 1) out->read_hp_count[sl * 9 + hp]++;
            }
            out->ps[r] = hp ? (acc.ps_seen ? acc.ps_min : -1) : 0;
            for (int k = 0; k < tv.n; k++) {
                const touch *x = &tv.a[k];
                if (x->vhp || (x->flags & 1)) { lps_call c = {x->var, (int16_t)(x->flags & 3), x->vhp, 0}; cv_push(&cv, c); }
            }
        } else {
            hp = somatic_read_hap(&acc, p->percentage_threshold, &pq);
            const int before = hp;
            float sim = 0.f;
            if (hp == HP_H3) {
                /* inheritHaplotype (SomaticHaplotagProcess.cpp:461-527) */
                int d1 = 0, d2 = 0;
                for (int k = 0; k < tv.n; k++)
                    if ((tv.a[k].flags & 4) && tv.a[k].vhp == 3) { int d = t->derive_hp[tv.a[k].var]; if (d == 1) d1++; else if (d == 2) d2++; }
                int mx = d1 > d2 ? d1 : d2, mn = d1 > d2 ? d2 : d1, mxhp = d1 > d2 ? 1 : 2;
                sim = mx == 0 ? 0.0f : (float)mx / ((float)mx + (float)mn);
                if (sim >= p->percentage_threshold) hp = mxhp == 1 ? HP_H1_1 : HP_H2_1;
            }
            out->hp_before[r] = (int8_t)before; out->derive_similarity[r] = sim;
            for (int k = 0; k < tv.n; k++) {
                const touch *x = &tv.a[k];
                if (!(x->flags & 4)) continue;
                const size_t sl = (size_t)slot[x->var];
                const int base_h3 = x->vhp == 3;
                out->hp_before_count[sl * 9 + before]++;
                if (before != HP_UNTAG && base_h3) out->h3_before_count[sl * 9 + before]++;
                out->hp_after_count[sl * 9 + hp]++;
                if (hp != HP_UNTAG && base_h3) out->h3_after_count[sl * 9 + hp]++;
                if (hp != HP_UNTAG) {
                    const int start = b->ref_start[r] + 1;
                    if (out->cover_start[sl] > start) out->cover_start[sl] = start;
                    if (out->cover_end[sl] < ref_pos) out->cover_end[sl] = ref_pos;
                }
            }
            /* PS rule (:409-430) */
            out->ps[r] = hp ? (acc.ps_seen ? acc.ps_min : -1) : 0;
            for (int k = 0; k < tv.n; k++) {
                const touch *x = &tv.a[k];
                if (x->vhp) { lps_call c = {x->var, (int16_t)(x->flags & 4), x->vhp, 0}; cv_push(&cv, c); }
            }
        }
        out->read_hp[r] = (int8_t)hp; out->pq[r] = pq;
    }
    /* postProcess of the extract passes: calculateBaseCommonInfo (SomaticVarCaller.cpp:13-40) via base_analysis
     * (HaplotagStrategy.h:162-191), ExtractNorDataChrProcessor::postProcess (:176-210), ExtractTumDataChrProcessor::postProcess
     * (:520-603).  Untouched positions keep the constructors' zeros, which is also what the formulas give for them. */
    if (mode != ORC_SOM_TAG) {
        for (int sl = 0; sl < out->n_tum; sl++) {
            const int vi = out->tum_var[sl];
            const int ty = var_type(t->ref_len[vi], t->alt_len[vi]);
            if (ty != VT_SNP && ty != VT_INS && ty != VT_DEL) continue;
            const int32_t *pb = out->pos_base + (size_t)sl * LPS_PB_FIELDS, *hpc = out->read_hp_count + (size_t)sl * 9;
            float *f = out->ratios_f + (size_t)sl * LPS_RF_FIELDS;
            double *d = out->ratios_d + (size_t)sl * LPS_RD_FIELDS;
            int alt = pb[LPS_PB_ALT], malt = pb[LPS_PB_MPQ_ALT];
            if (ty == VT_SNP) {
                const char ab = (char)t->alt0[vi];
                const int kk = ab == 'A' ? 0 : ab == 'C' ? 1 : ab == 'G' ? 2 : ab == 'T' ? 3 : -1;
                alt = kk < 0 ? 0 : pb[LPS_PB_A + kk]; malt = kk < 0 ? 0 : pb[LPS_PB_MPQ_A + kk];
            }
            const int depth = pb[LPS_PB_DEPTH], mdepth = pb[LPS_PB_MPQ_DEPTH], del = pb[LPS_PB_DEL];
#define VAF_(a_, d_) (((d_) == 0 || (a_) == 0) ? 0.0f : (float)(a_) / (float)(d_))
            f[LPS_RF_VAF] = VAF_(alt, depth); f[LPS_RF_MPQ_VAF] = VAF_(malt, mdepth); f[LPS_RF_NONDEL_VAF] = VAF_(alt, depth - del);
            f[LPS_RF_LOW_MPQ_RATIO] = depth == 0 ? 0.0f : (float)(depth - mdepth) / (float)depth;
            f[LPS_RF_DEL_RATIO] = VAF_(del, depth);
#define IMB_(a_, b_) (((a_) > 0 && (b_) > 0) ? ((a_) > (b_) ? (double)(a_) / (double)((a_) + (b_)) : (double)(b_) / (double)((a_) + (b_))) : (((a_) == 0 && (b_) == 0) ? 0.0 : 1.0))
            const int g1 = hpc[1], g2 = hpc[2];
            d[LPS_RD_GERMLINE_IMBALANCE] = IMB_(g1, g2);
            d[LPS_RD_PCT_GERMLINE_HP] = (depth == 0 || g1 + g2 == 0) ? 0.0 : (double)(g1 + g2) / (double)depth;
            if (mode == ORC_SOM_EXTRACT_TUMOR) {
                const int32_t *cc = out->case_count + (size_t)sl * LPS_CASE_FIELDS;
                const int clean = cc[LPS_CASE_CLEAN_HP3], mixed = cc[LPS_CASE_MIXED];
                out->case_read_count[sl] = clean + mixed;
                if (clean + mixed != 0) {
                    const float den = (float)clean + (float)mixed;
                    f[LPS_RF_MIXED_RATIO] = (float)mixed / den; f[LPS_RF_PURE_H1_1_RATIO] = (float)cc[LPS_CASE_PURE_H1_1] / den;
                    f[LPS_RF_PURE_H2_1_RATIO] = (float)cc[LPS_CASE_PURE_H2_1] / den; f[LPS_RF_PURE_H3_RATIO] = (float)cc[LPS_CASE_PURE_H3] / den;
                }
                const int b1 = hpc[1] + hpc[5], b2 = hpc[2] + hpc[7];
                d[LPS_RD_ALLELIC_IMBALANCE] = IMB_(b1, b2);
                d[LPS_RD_SOMATIC_IMBALANCE] = IMB_(hpc[5], hpc[7]);
            }
        }
    }
    for (int r = n; r >= 0; r--) if (r == n || out->call_off[r] > cv.n) out->call_off[r] = cv.n;
    out->call_off[n] = cv.n;
    out->n_calls = cv.n;
    out->calls = cv.a ? cv.a : (lps_call *)calloc(1, sizeof(lps_call));
    free(tv.a); free(slot);
    return rc;
}

void orc_somatic_free(orc_somatic_out *o) {
    free(o->tum_var); free(o->category); free(o->read_hp); free(o->hp_before); free(o->ps); free(o->pq); free(o->h1); free(o->h2);
    free(o->h3); free(o->n_ps); free(o->end_pos); free(o->read_len); free(o->derive_similarity); free(o->pos_base);
    free(o->read_hp_count); free(o->somatic_read_hp_count); free(o->case_count); free(o->allele_count); free(o->window_hist);
    free(o->hp_before_count); free(o->hp_after_count); free(o->h3_before_count); free(o->h3_after_count); free(o->cover_start);
    free(o->cover_end); free(o->call_off); free(o->calls); free(o->ratios_f); free(o->ratios_d); free(o->case_read_count);
    memset(o, 0, sizeof(*o));
}
