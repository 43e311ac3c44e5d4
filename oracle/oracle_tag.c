/*
 * oracle_tag.c — CPU restatement of the germline `haplotag` hot path (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 *   dispatch            ChromosomeProcessor::processSingleChrom      src/haplotag/HaplotagParsingBam.cpp:457-486
 *   CIGAR walk          CigarParser::parsingCigar                    src/haplotag/HaplotagParsingBam.cpp:541-647
 *   per-variant votes   GermlineHaplotagStrategy::judgeSnpHap        src/haplotag/HaplotagStrategy.cpp:20-130
 *                       ...::judgeDeletionHap                        src/haplotag/HaplotagStrategy.cpp:147-209
 *                       GermlineHaplotagCigarParser hooks            src/haplotag/HaplotagProcess.cpp:486-501
 *   per-read decision   GermlineHaplotagStrategy::judgeReadHap       src/haplotag/HaplotagStrategy.cpp:243-300
 * Parity status: pinned against the unmodified reference through oracle/ref_tap_tag.cpp.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static const char NT16[] = "=ACMGRSVTWYHKDBN";

typedef struct { lps_call *a; uint64_t n, cap; } callvec;
static void cv_push(callvec *c, lps_call x) {
    if (c->n == c->cap) { c->cap = c->cap * 2 + 1024; c->a = (lps_call *)realloc(c->a, sizeof(lps_call) * c->cap); }
    c->a[c->n++] = x;
}

int orc_tag_reads(const lps_read_batch *b, const lps_variants *v, const uint8_t *hom, const lps_tag_params *p, orc_tags *out) {
    memset(out, 0, sizeof(*out));
    int n = b->n_reads, nv = v->n;
    out->n_reads = n;
    out->category = (uint8_t *)calloc((size_t)n + 1, 1);
    out->hp = (int8_t *)calloc((size_t)n + 1, 1);
    out->ps = (int32_t *)calloc((size_t)n + 1, 4); out->pq = (int32_t *)calloc((size_t)n + 1, 4);
    out->h1 = (int32_t *)calloc((size_t)n + 1, 4); out->h2 = (int32_t *)calloc((size_t)n + 1, 4);
    out->call_off = (uint64_t *)calloc((size_t)n + 2, 8);
    callvec cv = {0, 0, 0};
    int last_pos = nv ? v->pos[nv - 1] : -1;
    int rc = 0;
    for (int r = 0; r < n && !rc; r++) {
        out->call_off[r] = cv.n;
        int flag = b->flag[r];
        int cat = LPS_TAG_PROCESSED;
        if ((int)b->mapq[r] < p->mapping_quality && p->mapq_filter) cat = LPS_TAG_LOW_MAPQ;
        else if (flag & 0x4) cat = LPS_TAG_UNMAPPED;
        else if (flag & 0x100) cat = LPS_TAG_SECONDARY;
        else if ((flag & 0x800) && !p->tag_supplementary) cat = LPS_TAG_SUPPLEMENTARY;
        else if (nv == 0) cat = LPS_TAG_EMPTY_VARIANTS;
        else if (!(b->ref_start[r] <= last_pos)) cat = LPS_TAG_OTHER;
        out->category[r] = (uint8_t)cat;
        if (cat != LPS_TAG_PROCESSED) continue;

        int h1 = 0, h2 = 0, ps_min = 0, ps_seen = 0, ps_multi = 0;
#define COUNT_PS(vi) do { int ps__ = v->ps[vi]; if (!ps_seen) { ps_seen = 1; ps_min = ps__; } else { if (ps__ != ps_min) ps_multi = 1; if (ps__ < ps_min) ps_min = ps__; } } while (0)
        int ref_pos = b->ref_start[r], qpos = 0, lq = b->l_qseq[r];
        int lo = 0, hi = nv;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (v->pos[mid] < ref_pos) lo = mid + 1; else hi = mid; }
        int cur = lo;                                             /* firstVariantIter (:555-563) */
        const uint32_t *cig = b->cigar + b->cigar_off[r];
        const uint8_t *seq = b->seq4 + b->seq_off[r];
        int ncig = (int)b->n_cigar[r];
        if (cur == nv) ncig = 0;                                  /* "return" when no variant is left (:559-561) */
        for (int i = 0; i < ncig; i++) {
            int op = (int)(cig[i] & 15), len = (int)(cig[i] >> 4);
            while (cur < nv && v->pos[cur] < ref_pos) cur++;
            if (op == 0 || op == 7 || op == 8) {
                while (cur < nv && v->pos[cur] < ref_pos + len) {
                    int vp = v->pos[cur], off = vp - ref_pos;
                    int rl = v->ref_len[cur], al = v->alt_len[cur];
                    int h1alt = v->hp1_is_alt[cur] != 0;
                    int hpbit = -1, counted = 0, kind = 0;
                    if (rl == 1 && al == 1) {
                        /* the reference reads seq[query_pos+offset] unchecked; beyond l_qseq the result is undefined and the hit is dropped */
                        if (qpos + off < lq) {
                            char c = NT16[(seq[(qpos + off) >> 1] >> ((~(qpos + off) & 1) << 2)) & 0xf];
                            char rb = (char)v->ref0[cur], ab = (char)v->alt0[cur];
                            if (c == rb || c == ab) {
                                counted = 1;
                                if (c == (h1alt ? ab : rb)) hpbit = 0;
                                if (c == (h1alt ? rb : ab)) hpbit = 1;
                            }
                        }
                    } else if ((rl == 1) != (al == 1)) {
                        if (i + 1 < ncig) {
                            int want = rl == 1 ? 1 : 2;
                            int has = (ref_pos + len - 1 == vp && (int)(cig[i + 1] & 15) == want);
                            int l1 = h1alt ? al : rl, l2 = h1alt ? rl : al;
                            if (l1 != 1 && l2 == 1) hpbit = has ? 0 : 1;
                            else if (l1 == 1 && l2 != 1) hpbit = has ? 1 : 0;
                            counted = 1; kind = 2;
                        }
                    }
                    if (counted) {
                        COUNT_PS(cur);
                        if (hpbit == 0) h1++; else if (hpbit == 1) h2++;
                        lps_call c = {cur, 1, (int8_t)hpbit, (int8_t)kind};
                        cv_push(&cv, c);
                    }
                    cur++;
                }
                qpos += len; ref_pos += len;
            } else if (op == 1) {
                qpos += len;
            } else if (op == 2) {
                int judged = 0;
                while (cur < nv && v->pos[cur] < ref_pos + len) {
                    /* every NORMAL variant of a phased VCF is PHASED_HETERO: the first one in the op is judged */
                    if (!judged) {
                        judged = 1;
                        int vp = v->pos[cur];
                        if (p->have_reference && !(ref_pos + len + 1 == vp) && vp >= ref_pos && hom[cur] >= 3) {
                            int rl = v->ref_len[cur], al = v->alt_len[cur];
                            int h1alt = v->hp1_is_alt[cur] != 0;
                            if (rl == 1 && al == 1) {
                                if (qpos < lq) {
                                    char c = NT16[(seq[qpos >> 1] >> ((~qpos & 1) << 2)) & 0xf];
                                    char rb = (char)v->ref0[cur], ab = (char)v->alt0[cur];
                                    int hpbit = -1;
                                    if (c == (h1alt ? ab : rb)) hpbit = 0;
                                    if (c == (h1alt ? rb : ab)) hpbit = 1;
                                    COUNT_PS(cur);
                                    if (hpbit == 0) h1++; else if (hpbit == 1) h2++;
                                    lps_call cc = {cur, 1, (int8_t)hpbit, 1};
                                    cv_push(&cv, cc);
                                }
                            } else if (rl != 1 && al == 1) {
                                int l1 = h1alt ? al : rl, l2 = h1alt ? rl : al, hpbit = -1;
                                if (l1 != 1 && l2 == 1) hpbit = 0; else if (l1 == 1 && l2 != 1) hpbit = 1;
                                COUNT_PS(cur);
                                if (hpbit == 0) h1++; else if (hpbit == 1) h2++;
                                lps_call cc = {cur, 1, (int8_t)hpbit, 2};
                                cv_push(&cv, cc);
                            }
                        }
                    }
                    cur++;
                }
                ref_pos += len;
            } else if (op == 3) ref_pos += len;
            else if (op == 4) qpos += len;
            else if (op == 5 || op == 6) {}
            else { rc = LPS_E_CIGAR; break; }
        }
        /* judgeReadHap */
        double mx = h1 > h2 ? h1 : h2, mn = h1 > h2 ? h2 : h1;
        int hp = 0, pq;
        if (!(mx / (mx + mn) < p->percentage_threshold)) { if (h1 > h2) hp = 1; if (h1 < h2) hp = 2; }
        if (mx == 0) pq = 0; else if (mx == (mx + mn)) pq = 40; else pq = (int)(-10 * (log10((double)mn / (double)(mx + mn))));
        if (ps_multi) hp = 0;
        out->hp[r] = (int8_t)hp; out->ps[r] = hp ? ps_min : 0; out->pq[r] = pq; out->h1[r] = h1; out->h2[r] = h2;
    }
    for (int r = n; r >= 0; r--) if (r == n || out->call_off[r] > cv.n) out->call_off[r] = cv.n;
    out->call_off[n] = cv.n;
    out->n_calls = cv.n;
    out->calls = cv.a ? cv.a : (lps_call *)calloc(1, sizeof(lps_call));
    return rc;
}

void orc_tags_free(orc_tags *t) {
    free(t->category); free(t->hp); free(t->ps); free(t->pq); free(t->h1); free(t->h2); free(t->call_off); free(t->calls);
    memset(t, 0, sizeof(*t));
}
