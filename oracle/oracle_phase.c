/*
 * oracle_phase.c — CPU restatement of the `phase` hot path (TEST INFRASTRUCTURE ONLY).
 * See oracle.h for the parity status and the rules on who may call this.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */
static const char NT16[] = "=ACMGRSVTWYHKDBN"; /* htslib seq_nt16_str (htslib/hts.c:257) */

static inline char ref_at(const char *ref, int64_t len, int64_t i) {
    /* std::string operator[] returns '\0' at size(); beyond is UB in the reference — read as NUL */
    return (i >= 0 && i < len) ? ref[i] : '\0';
}

/* homopolymerLength — src/shared/Util.cpp:21-54 */
static int homopolymer_len(const char *ref, int64_t len, int pos) {
    int h = 1;
    if ((int64_t)pos + 1 >= len) return h;
    char e = ref[pos];
    int64_t p = (int64_t)pos - 1;
    /* the reference calls ref.at(-1) (throws) when pos == 0; we stop instead */
    while (p >= 0 && ref[p] == e) {
        p--; h++;
        if (h >= 10 || p < 0) break;
    }
    p = (int64_t)pos + 1;
    if (p < len) {
        while (ref[p] == e) {
            p++; h++;
            if (p >= len) break;
            if (h >= 10) break;
        }
    }
    return h;
}

int orc_annotate(const char *ref, int64_t ref_len, const lps_variants *v, int is_ont,
                 uint8_t *hom, uint8_t *danger, uint8_t *filtered) {
    for (int i = 0; i < v->n; i++) {
        int pos = v->pos[i];
        hom[i] = (uint8_t)homopolymer_len(ref, ref_len, pos);
        /* getVariants_markindel — src/phase/ParsingBam.cpp:378-417: the 2-mer right behind the
         * variant position must repeat five times in a row (the first comparison is with itself) */
        int d = 0;
        if (v->ref_len[i] > 1 || v->alt_len[i] > 1) {
            char r0 = ref_at(ref, ref_len, (int64_t)pos + 1), r1 = ref_at(ref, ref_len, (int64_t)pos + 2);
            if ((int64_t)pos + 2 >= ref_len) r1 = '\0';
            int64_t rp = pos;
            int k = 0;
            while (k < 5) {
                if (r0 != ref_at(ref, ref_len, rp + 1) || r1 != ref_at(ref, ref_len, rp + 2)) break;
                rp += 2; k++;
            }
            d = (k == 5);
        }
        danger[i] = (uint8_t)d;
        filtered[i] = 0;
    }
    if (is_ont) {
        /* SnpParser::filterSNP — src/phase/ParsingBam.cpp:866-888: walking left to right, the right
         * member of a pair is erased when both sit in a homopolymer >= 3 and are <= 2 bp apart; the
         * left member then meets the following variant. */
        int cur = 0, nxt = 1;
        while (cur < v->n && nxt < v->n) {
            if (hom[cur] >= 3 && hom[nxt] >= 3 && abs(v->pos[cur] - v->pos[nxt]) <= 2) {
                filtered[nxt] = 1;
                nxt++;
                continue;
            }
            cur = nxt;
            nxt++;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* allele calling                                                                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct { lps_call *a; uint64_t n, cap; } callvec;
static void cv_push(callvec *c, lps_call x) {
    if (c->n == c->cap) { c->cap = c->cap * 2 + 1024; c->a = (lps_call *)realloc(c->a, sizeof(lps_call) * c->cap); }
    c->a[c->n++] = x;
}
typedef struct { int32_t pos, side; } clip_ev;
typedef struct { clip_ev *a; size_t n, cap; } clipvec;
static void clip_push(clipvec *c, int32_t pos, int32_t side) {
    if (c->n == c->cap) { c->cap = c->cap * 2 + 256; c->a = (clip_ev *)realloc(c->a, sizeof(clip_ev) * c->cap); }
    c->a[c->n].pos = pos; c->a[c->n].side = side; c->n++;
}
static int cmp_clip(const void *a, const void *b) {
    const clip_ev *x = (const clip_ev *)a, *y = (const clip_ev *)b;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

static inline char base_at(const lps_read_batch *b, int r, int q) {
    const uint8_t *s = b->seq4 + b->seq_off[r];
    return NT16[(s[q >> 1] >> ((~q & 1) << 2)) & 0xf]; /* bam_seqi, htslib/sam.h:325 */
}

int orc_call_alleles(const lps_read_batch *b, const lps_variants *v, const uint8_t *hom, const uint8_t *danger,
                     const uint8_t *filtered, int apply_filter, const lps_phase_params *p, orc_calls *out) {
    memset(out, 0, sizeof(*out));
    callvec cv = {0, 0, 0};
    clipvec cl = {0, 0, 0};
    int n = b->n_reads, nv = v->n;
    out->n_reads = n;
    out->call_off = (uint64_t *)calloc((size_t)n + 1, sizeof(uint64_t));
    out->read_status = (uint8_t *)calloc((size_t)n + 1, 1);
    int last_var_pos = nv ? v->pos[nv - 1] : -1;
    int rc = 0;
    for (int r = 0; r < n; r++) {
        out->call_off[r] = cv.n;
        /* region "chr:1-lastSNP" of the iterator (ParsingBam.cpp:1273) + read filter (:1282-1291) */
        int flag = b->flag[r];
        if (b->ref_start[r] >= last_var_pos || (int)b->mapq[r] < p->mapping_quality || (flag & 0x4) || (flag & 0x100) || (flag & 0x400)) {
            out->read_status[r] = LPS_READ_FILTERED;
            continue;
        }
        uint64_t first_call = cv.n;
        size_t first_clip = cl.n;
        int ref_pos = b->ref_start[r], qpos = 0, lq = b->l_qseq[r];
        /* the stateful firstVariantIter equals lower_bound for a coordinate-sorted BAM (:1318-1330) */
        int lo = 0, hi = nv;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (v->pos[mid] < ref_pos) lo = mid + 1; else hi = mid; }
        int cur = lo;
        const uint32_t *cig = b->cigar + b->cigar_off[r];
        int ncig = (int)b->n_cigar[r];
        int aborted = 0;
        for (int i = 0; i < ncig && !aborted; i++) {
            int op = (int)(cig[i] & 15), len = (int)(cig[i] >> 4);
            while (cur < nv && v->pos[cur] < ref_pos) cur++;                      /* :1361-1364 */
            if (op == 0 || op == 7 || op == 8) {
                while (cur < nv && v->pos[cur] < ref_pos + len) {                  /* :1368-1370 */
                    int vp = v->pos[cur], off = vp - ref_pos;
                    if (qpos + off + 1 > lq) { aborted = 1; break; }               /* :1453-1455 */
                    int rl = v->ref_len[cur], al = v->alt_len[cur];
                    int allele = -1, q = 0;
                    if (rl == 1 && al == 1) {                                      /* :1458-1466 */
                        char c = base_at(b, r, qpos + off);
                        if (c == (char)v->ref0[cur]) allele = 0;
                        else if (c == (char)v->alt0[cur]) allele = 1;
                        q = b->qual[b->qual_off[r] + (uint64_t)(qpos + off)];
                    }
                    if (rl == 1 && al != 1 && i + 1 < ncig) {                      /* :1470-1491 */
                        allele = (ref_pos + len - 1 == vp && (cig[i + 1] & 15) == 1) ? 1 : 0;
                        q = danger[cur] ? -5 : -4;
                    }
                    if (rl != 1 && al == 1 && i + 1 < ncig) {                      /* :1495-1510 */
                        allele = (ref_pos + len - 1 == vp && (cig[i + 1] & 15) == 2) ? 1 : 0;
                        q = danger[cur] ? -5 : -4;
                    }
                    if (allele != -1) {
                        lps_call c = {cur, (int16_t)q, (int8_t)allele, 0};
                        cv_push(&cv, c);
                    }
                    cur++;
                }
                if (aborted) break;
                qpos += len; ref_pos += len;
            } else if (op == 1) {
                qpos += len;
            } else if (op == 2) {                                                  /* :1539-1607 */
                if (p->have_reference && cur < nv) {
                    int vp = v->pos[cur];
                    if (ref_pos + len + 1 == vp) {
                        /* nothing */
                    } else if (vp >= ref_pos && vp < ref_pos + len && hom[cur] >= 3) {
                        if (qpos + 1 > lq) { aborted = 1; break; }                 /* :1559-1561 */
                        int rl = v->ref_len[cur], al = v->alt_len[cur];
                        int allele = -1, q = 0;
                        if (rl == 1 && al == 1) {
                            char c = base_at(b, r, qpos);
                            if (c == (char)v->ref0[cur]) allele = 0;
                            else if (c == (char)v->alt0[cur]) allele = 1;
                            q = b->qual[b->qual_off[r] + (uint64_t)qpos];
                        } else if (rl != 1 && al == 1) {
                            allele = 1; q = -4;
                        }
                        if (allele != -1) {
                            lps_call c = {cur, (int16_t)q, (int8_t)allele, 1};
                            cv_push(&cv, c);
                            cur++;
                        }
                    }
                }
                ref_pos += len;
            } else if (op == 3) {
                ref_pos += len;
            } else if (op == 4) {
                qpos += len;
                if (len > 5) clip_push(&cl, ref_pos, i == 0 ? 0 : 1);             /* :1636-1645 */
            } else if (op == 5) {
                if (len > 5) clip_push(&cl, ref_pos, i == 0 ? 0 : 1);
            } else if (op == 6) {
            } else {
                rc = LPS_E_CIGAR;                                                 /* :1625-1628 exit(1) */
                aborted = 1;
            }
        }
        (void)first_clip; /* clips recorded before an abort stay counted (the map is updated in place) */
        if (aborted) {
            cv.n = first_call;                                                     /* `return` drops the read */
            out->read_status[r] = LPS_READ_ABORTED;
        } else if (apply_filter) {
            uint64_t w = first_call;
            for (uint64_t k = first_call; k < cv.n; k++) if (!filtered[cv.a[k].var]) cv.a[w++] = cv.a[k];
            cv.n = w;
        }
        if (rc) break;
    }
    out->call_off[n] = cv.n;
    for (int r = n - 1; r >= 0; r--) if (out->call_off[r] > out->call_off[r + 1]) out->call_off[r] = out->call_off[r + 1];
    out->n_calls = cv.n;
    out->calls = cv.a ? cv.a : (lps_call *)calloc(1, sizeof(lps_call));
    /* clipCount map: pos -> {FRONT, BACK} */
    qsort(cl.a, cl.n, sizeof(clip_ev), cmp_clip);
    out->clip_pos = (int32_t *)calloc(cl.n + 1, 4); out->clip_front = (int32_t *)calloc(cl.n + 1, 4); out->clip_back = (int32_t *)calloc(cl.n + 1, 4);
    int m = 0;
    for (size_t k = 0; k < cl.n; k++) {
        if (m == 0 || out->clip_pos[m - 1] != cl.a[k].pos) { out->clip_pos[m] = cl.a[k].pos; m++; }
        if (cl.a[k].side == 0) out->clip_front[m - 1]++; else out->clip_back[m - 1]++;
    }
    out->n_clips = m;
    free(cl.a);
    return rc;
}

void orc_calls_free(orc_calls *c) {
    free(c->call_off); free(c->calls); free(c->read_status); free(c->clip_pos); free(c->clip_front); free(c->clip_back);
    memset(c, 0, sizeof(*c));
}

/* ------------------------------------------------------------------------------------------ */
/* Clip::getCNVInterval — src/phase/PhasingGraph.cpp:1112-1227                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int push, slow_up, slow_down, curr, reject, pull_down, slow_down_cnt, cand_start, cand_end; } cnv_state;
static void cnv_reset(cnv_state *s) { memset(s, 0, sizeof(*s)); s->cand_start = -1; s->cand_end = -1; }
static void cnv_threshold(cnv_state *s, int up) {
    s->reject = up;
    if (up >= 20) { s->pull_down = up / 2; s->slow_down_cnt = 5; }
    else if (up >= 10) { s->pull_down = up / 2; s->slow_down_cnt = up / 4; }
    else { s->pull_down = 5; s->slow_down_cnt = 2; }
}
typedef struct { int32_t *s, *e; int n, cap; } ivec;
static void iv_push(ivec *v, int s, int e) {
    if (v->n == v->cap) { v->cap = v->cap * 2 + 16; v->s = (int32_t *)realloc(v->s, 4 * (size_t)v->cap); v->e = (int32_t *)realloc(v->e, 4 * (size_t)v->cap); }
    v->s[v->n] = s; v->e[v->n] = e; v->n++;
}
static void cnv_intervals(const orc_calls *c, ivec *out) {
    const int area = 30000;
    int n = c->n_clips;
    if (n == 0) return; /* the reference dereferences rbegin() of an empty map here (segfault) */
    cnv_state st; cnv_reset(&st);
    for (int k = 0; k <= n; k++) {
        /* the sentinel entry copies the last real one, AreaSize further right (:1134) */
        int pos = k < n ? c->clip_pos[k] : c->clip_pos[n - 1] + area;
        int up = k < n ? c->clip_front[k] : c->clip_front[n - 1];
        int down = k < n ? c->clip_back[k] : c->clip_back[n - 1];
        if (!st.push && !st.slow_down && !st.slow_up) {
            if (up >= 5 && st.curr == 0) {
                st.push = 1; st.slow_up = 0; st.slow_down = 1; st.curr = up - down; st.cand_start = pos; st.cand_end = pos + area;
                cnv_threshold(&st, up);
            } else if (up > down && st.curr == 0) {
                st.push = 0; st.slow_up = 1; st.slow_down = 0; st.curr = up - down; st.cand_start = pos; st.cand_end = pos + area;
            }
        } else if (st.push && st.slow_down) {
            if (up > st.reject) {
                st.push = 1; st.slow_up = 0; st.slow_down = 1; cnv_threshold(&st, up); st.cand_start = pos; st.cand_end = pos + area;
            }
            st.curr = st.curr + up - down;
            if (st.curr > 30) st.cand_end = pos + area;
            if (down >= st.pull_down) { iv_push(out, st.cand_start, pos); cnv_reset(&st); }
            else if (st.curr <= st.slow_down_cnt && pos <= st.cand_end) { iv_push(out, st.cand_start, pos); cnv_reset(&st); }
            if (pos > st.cand_end || st.curr <= 0 || pos - st.cand_start >= 200000) cnv_reset(&st);
        } else if (st.slow_up) {
            if (st.curr > 20 ? down >= st.curr / 4 : down >= 5) { iv_push(out, st.cand_start, pos); cnv_reset(&st); }
            else if (up >= 5) {
                st.push = 1; st.slow_up = 0; st.slow_down = 1; st.curr = up - down; st.cand_start = pos; st.cand_end = pos + area;
                cnv_threshold(&st, up);
            } else {
                st.curr = st.curr + up - down;
                if (st.curr > 30) st.cand_end = pos + area;
                if (pos > st.cand_end || st.curr <= 0 || pos - st.cand_start >= 200000) cnv_reset(&st);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* graph construction                                                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t key, cnt; } kc;
typedef struct { kc *a; int n, cap; } kcvec;
static void kc_inc(kcvec *m, int key) {
    for (int i = 0; i < m->n; i++) if (m->a[i].key == key) { m->a[i].cnt++; return; }
    if (m->n == m->cap) { m->cap = m->cap * 2 + 4; m->a = (kc *)realloc(m->a, sizeof(kc) * (size_t)m->cap); }
    m->a[m->n].key = key; m->a[m->n].cnt = 1; m->n++;
}
static int kc_find(const kcvec *m, int key, int *cnt) {
    for (int i = 0; i < m->n; i++) if (m->a[i].key == key) { *cnt = m->a[i].cnt; return 1; }
    return 0;
}
typedef struct { int32_t *a; int n, cap; } intvec;
static void int_push(intvec *v, int x) {
    if (v->n == v->cap) { v->cap = v->cap * 2 + 8; v->a = (int32_t *)realloc(v->a, 4 * (size_t)v->cap); }
    v->a[v->n++] = x;
}
static double int_mean(const intvec *v) {
    if (v->n == 0) return 0;
    double s = 0.0;
    for (int i = 0; i < v->n; i++) s += v->a[i];
    return s / v->n;
}
static inline int in_range(int p, int s, int e) { return p >= s && p <= e; }

typedef struct { int32_t rank, order; } rk;
static int cmp_rk(const void *a, const void *b) {
    const rk *x = (const rk *)a, *y = (const rk *)b;
    if (x->rank != y->rank) return x->rank < y->rank ? -1 : 1;
    return x->order < y->order ? -1 : (x->order > y->order ? 1 : 0);
}

/* far-cell hash (cells beyond the sweep window; written by the reference, never read) */
typedef struct { uint64_t *keys; float *vals; size_t cap, n; } fhash;
static float *fh_get(fhash *h, uint64_t key) {
    if ((h->n + 1) * 2 > h->cap) {
        size_t ncap = h->cap ? h->cap * 2 : 1024;
        uint64_t *nk = (uint64_t *)malloc(8 * ncap); float *nv = (float *)calloc(ncap, 4);
        memset(nk, 0xff, 8 * ncap);
        for (size_t i = 0; i < h->cap; i++) if (h->keys[i] != UINT64_MAX) {
            size_t j = (size_t)((h->keys[i] * 0x9e3779b97f4a7c15ULL) >> 20) & (ncap - 1);
            while (nk[j] != UINT64_MAX) j = (j + 1) & (ncap - 1);
            nk[j] = h->keys[i]; nv[j] = h->vals[i];
        }
        free(h->keys); free(h->vals); h->keys = nk; h->vals = nv; h->cap = ncap;
    }
    size_t j = (size_t)((key * 0x9e3779b97f4a7c15ULL) >> 20) & (h->cap - 1);
    while (h->keys[j] != UINT64_MAX && h->keys[j] != key) j = (j + 1) & (h->cap - 1);
    if (h->keys[j] == UINT64_MAX) { h->keys[j] = key; h->n++; }
    return &h->vals[j];
}

int orc_build_graph(const lps_read_batch *b, const lps_variants *v, const uint8_t *danger, const orc_calls *calls,
                    const lps_phase_params *p, orc_graph *out) {
    (void)danger;
    memset(out, 0, sizeof(*out));
    int nr = calls->n_reads, nv = v->n;
    /* alignments with >= 1 call (the reference keeps alignments emptied by filterSNP and then reads
     * front()/back() of an empty vector — undefined; they are dropped here) */
    int na = 0;
    for (int r = 0; r < nr; r++) if (calls->call_off[r + 1] > calls->call_off[r]) na++;
    int32_t *aln = (int32_t *)malloc(4 * (size_t)(na + 1));
    na = 0;
    for (int r = 0; r < nr; r++) if (calls->call_off[r + 1] > calls->call_off[r]) aln[na++] = r;

    /* ---- overlap filter: PhasingGraph.cpp:707-781 ---- */
    int max_rank = 0;
    for (int r = 0; r < nr; r++) if (b->name_rank[r] > max_rank) max_rank = b->name_rank[r];
    int32_t *range2 = (int32_t *)calloc((size_t)max_rank + 1, 4);   /* alignRange[name].second, .first stays 0 */
    int32_t *top = (int32_t *)malloc(4 * ((size_t)max_rank + 1));   /* readIdxVec[name].back() */
    int32_t *below = (int32_t *)malloc(4 * (size_t)(na + 1));
    uint8_t *dead = (uint8_t *)calloc((size_t)na + 1, 1);
    for (int i = 0; i <= max_rank; i++) top[i] = -1;
#define FIRSTPOS(k) (v->pos[calls->calls[calls->call_off[aln[k]]].var])
#define LASTPOS(k) (v->pos[calls->calls[calls->call_off[aln[k] + 1] - 1].var])
    for (int k = 0; k < na; k++) {
        int rank = b->name_rank[aln[k]];
        int first = FIRSTPOS(k), last = LASTPOS(k);
        int del_cur = 0;
        while (0 <= first && first <= range2[rank]) {
            if (last < range2[rank]) { del_cur = 1; break; }
            int pv = top[rank];
            if (pv < 0) break;
            int ps = FIRSTPOS(pv), pe = LASTPOS(pv);
            double os = ps > first ? ps : first, oe = pe < last ? pe : last;
            if (os > oe) break;
            double olen = oe - os + 1;
            double as = pe > last ? pe : last, ae = ps < first ? ps : first;
            double ratio = olen / (as - ae + 1);
            if (ratio >= p->overlap_threshold) {
                int len1 = pe - ps + 1, len2 = last - first + 1;
                if (len2 <= len1) { del_cur = 1; break; }
                dead[pv] = 1;
                top[rank] = below[pv];
                range2[rank] = top[rank] >= 0 ? LASTPOS(top[rank]) : first;
            } else break;
        }
        range2[rank] = last;
        if (del_cur) dead[k] = 1;
        else { below[k] = top[rank]; top[rank] = k; }
    }
    /* surviving alignments with their own copy of the calls */
    int ns = 0;
    uint64_t ncall = 0;
    for (int k = 0; k < na; k++) if (!dead[k]) { ns++; ncall += calls->call_off[aln[k] + 1] - calls->call_off[aln[k]]; }
    out->n_aln = ns;
    out->aln_read = (int32_t *)malloc(4 * (size_t)(ns + 1));
    out->aln_off = (uint64_t *)malloc(8 * (size_t)(ns + 1));
    out->aln_calls = (lps_call *)malloc(sizeof(lps_call) * (ncall + 1));
    ns = 0; ncall = 0;
    for (int k = 0; k < na; k++) if (!dead[k]) {
        out->aln_read[ns] = aln[k];
        out->aln_off[ns] = ncall;
        uint64_t c0 = calls->call_off[aln[k]], c1 = calls->call_off[aln[k] + 1];
        memcpy(out->aln_calls + ncall, calls->calls + c0, sizeof(lps_call) * (c1 - c0));
        ncall += c1 - c0;
        ns++;
    }
    out->aln_off[ns] = ncall;
    free(range2); free(top); free(below); free(dead); free(aln);

    /* ---- CNV intervals: the state machine runs twice (Clip ctor + PhasingProcess.cpp:148) ---- */
    ivec cnv = {0, 0, 0, 0};
    cnv_intervals(calls, &cnv);
    cnv_intervals(calls, &cnv);
    out->n_cnv = cnv.n;
    out->cnv_start = (int32_t *)malloc(4 * (size_t)(cnv.n + 1)); out->cnv_end = (int32_t *)malloc(4 * (size_t)(cnv.n + 1));
    for (int i = 0; i < cnv.n; i++) { out->cnv_start[i] = cnv.s[i]; out->cnv_end[i] = cnv.e[i]; }

    /* ---- CNV mismatch filter: PhasingGraph.cpp:520-692 (literal index walking, the interval
     *      vector is NOT sorted because of the duplication) ---- */
    if (ns > 0 && cnv.n > 0) {
        kcvec *mm = (kcvec *)calloc((size_t)ns, sizeof(kcvec));
        size_t ci = 0;
        for (int k = 0; k < ns; k++) {                                      /* calculateCnvMismatchRate */
            uint64_t c0 = out->aln_off[k], c1 = out->aln_off[k + 1];
            if (c0 == c1) continue;
            int rs = v->pos[out->aln_calls[c0].var], re = v->pos[out->aln_calls[c1 - 1].var];
            while (ci > 0 && cnv.s[ci] > rs) ci--;
            size_t i = ci;
            while (i < (size_t)cnv.n && cnv.s[i] <= re) {
                for (uint64_t c = c0; c < c1; c++) {
                    int vp = v->pos[out->aln_calls[c].var];
                    if (vp > cnv.e[i]) break;
                    if (in_range(vp, cnv.s[i], cnv.e[i]) && out->aln_calls[c].allele == 1) kc_inc(&mm[k], cnv.s[i]);
                }
                i++;
            }
            ci = i > 0 ? i - 1 : 0;
        }
        intvec *agg = (intvec *)calloc((size_t)nv * 2, sizeof(intvec));     /* aggregateCnvReadMismatchRate */
        ci = 0;
        for (int k = 0; k < ns; k++) {
            uint64_t c0 = out->aln_off[k], c1 = out->aln_off[k + 1];
            if (c0 == c1) continue;
            int rs = v->pos[out->aln_calls[c0].var], re = v->pos[out->aln_calls[c1 - 1].var];
            while (ci > 0 && cnv.s[ci] > rs) ci--;
            size_t i = ci;
            while (i < (size_t)cnv.n && cnv.s[i] <= re) {
                for (uint64_t c = c0; c < c1; c++) {
                    int vi = out->aln_calls[c].var, vp = v->pos[vi], cnt;
                    if (vp > cnv.e[i]) break;
                    if (in_range(vp, cnv.s[i], cnv.e[i]) && kc_find(&mm[k], cnv.s[i], &cnt)) int_push(&agg[2 * vi + out->aln_calls[c].allele], cnt);
                }
                i++;
            }
            ci = i > 0 ? i - 1 : 0;
        }
        double *miss = (double *)malloc(8 * (size_t)nv);                    /* calculateAverageMismatchRate */
        uint8_t *has = (uint8_t *)calloc((size_t)nv, 1);
        int any = 0;
        for (int vi = 0; vi < nv; vi++) {
            if (agg[2 * vi].n == 0 && agg[2 * vi + 1].n == 0) continue;     /* not a key of cnvReadMmrate */
            for (size_t i = 0; i < (size_t)cnv.n; i++) {                    /* cnvIndex stays 0 in this function */
                if (cnv.s[i] > v->pos[vi]) break;
                if (in_range(v->pos[vi], cnv.s[i], cnv.e[i]) && agg[2 * vi].n && agg[2 * vi + 1].n) {
                    double mr = int_mean(&agg[2 * vi]), ma = int_mean(&agg[2 * vi + 1]);
                    if (mr != 0 && ma != 0) { miss[vi] = ma / (mr + ma); has[vi] = 1; any = 1; }
                }
            }
        }
        if (any) {                                                          /* filterHighMismatchVariants */
            ci = 0;
            uint64_t w = 0;
            uint64_t *noff = (uint64_t *)malloc(8 * (size_t)(ns + 1));
            for (int k = 0; k < ns; k++) {
                uint64_t c0 = out->aln_off[k], c1 = out->aln_off[k + 1];
                noff[k] = w;
                if (c0 == c1) continue;
                int rs = v->pos[out->aln_calls[c0].var];
                while (ci > 0 && cnv.s[ci] > rs) ci--;
                for (uint64_t c = c0; c < c1; c++) {
                    int vi = out->aln_calls[c].var, vp = v->pos[vi];
                    int erase = 0;
                    size_t i = ci;
                    while (i < (size_t)cnv.n && cnv.s[i] <= vp) {
                        if (in_range(vp, cnv.s[i], cnv.e[i]) && has[vi] && miss[vi] >= 0.7) { erase = 1; break; }
                        i++;
                    }
                    if (!erase) out->aln_calls[w++] = out->aln_calls[c];
                    ci = i > 0 ? i - 1 : 0;
                }
            }
            noff[ns] = w;
            memcpy(out->aln_off, noff, 8 * (size_t)(ns + 1));
            free(noff);
        }
        for (int k = 0; k < ns; k++) free(mm[k].a);
        for (int i = 0; i < 2 * nv; i++) free(agg[i].a);
        free(mm); free(agg); free(miss); free(has);
    }
    free(cnv.s); free(cnv.e);

    /* ---- node set, node types (last writer wins, :803-832) ---- */
    int32_t *node_of = (int32_t *)malloc(4 * (size_t)(nv + 1));
    int8_t *vtype = (int8_t *)malloc((size_t)nv + 1);
    memset(vtype, -1, (size_t)nv + 1);
    for (uint64_t c = 0; c < out->aln_off[ns]; c++) {
        int q = out->aln_calls[c].quality;
        vtype[out->aln_calls[c].var] = (int8_t)(q == -4 ? 3 : (q == -5 ? 4 : 0));
    }
    int nn = 0;
    for (int i = 0; i < nv; i++) node_of[i] = vtype[i] >= 0 ? nn++ : -1;
    out->n_nodes = nn;
    out->node_var = (int32_t *)malloc(4 * (size_t)(nn + 1));
    out->node_type = (uint8_t *)malloc((size_t)nn + 1);
    for (int i = 0; i < nv; i++) if (node_of[i] >= 0) { out->node_var[node_of[i]] = i; out->node_type[node_of[i]] = (uint8_t)vtype[i]; }
    int W = p->connect_adjacent;
    out->window = W;
    out->weights = (float *)calloc((size_t)nn * (size_t)W * 4 + 4, sizeof(float));

    /* ---- merge by name (lexicographic order == name_rank order), sort, fan out: :795-888 ---- */
    rk *order = (rk *)malloc(sizeof(rk) * (size_t)(ns + 1));
    for (int k = 0; k < ns; k++) { order[k].rank = b->name_rank[out->aln_read[k]]; order[k].order = k; }
    qsort(order, (size_t)ns, sizeof(rk), cmp_rk);
    fhash far = {0, 0, 0, 0};
    int32_t *gpos = 0, *gperm = 0; lps_call *gcall = 0; size_t gcap = 0;
    for (int g0 = 0; g0 < ns;) {
        int g1 = g0;
        size_t m = 0;
        while (g1 < ns && order[g1].rank == order[g0].rank) { int k = order[g1].order; m += (size_t)(out->aln_off[k + 1] - out->aln_off[k]); g1++; }
        if (m > gcap) { gcap = m * 2 + 64; gpos = (int32_t *)realloc(gpos, 4 * gcap); gperm = (int32_t *)realloc(gperm, 4 * gcap); gcall = (lps_call *)realloc(gcall, sizeof(lps_call) * gcap); }
        m = 0;
        for (int g = g0; g < g1; g++) {
            int k = order[g].order;
            for (uint64_t c = out->aln_off[k]; c < out->aln_off[k + 1]; c++) { gcall[m] = out->aln_calls[c]; gpos[m] = v->pos[gcall[m].var]; gperm[m] = (int32_t)m; m++; }
        }
        if (g1 - g0 > 1) orc_std_sort_by_pos(gpos, gperm, (int32_t)m);   /* ReadVariant::sort(), Util.cpp:3-5 */
        for (size_t a = 0; a + 1 < m; a++) {
            const lps_call *ca = &gcall[gperm[a]];
            int qa = ca->quality < 0 ? 60 : ca->quality;                 /* -4/-5 -> 60 (:820-828) */
            int na_ = node_of[ca->var];
            for (size_t d = 1; d <= (size_t)W && a + d < m; d++) {
                const lps_call *cb = &gcall[gperm[a + d]];
                int qb = cb->quality < 0 ? 60 : cb->quality;
                int nb = node_of[cb->var];
                int which = ca->allele * 2 + cb->allele;
                float *cell;
                int dist = nb - na_;
                if (dist >= 1 && dist <= W) { cell = &out->weights[((size_t)na_ * (size_t)W + (size_t)(dist - 1)) * 4 + (size_t)which]; out->n_contrib++; }
                else { cell = fh_get(&far, ((uint64_t)(uint32_t)na_ << 34) | ((uint64_t)(uint32_t)nb << 2) | (uint64_t)which); out->n_contrib_far++; }
                /* SubEdge::addSubEdge (:40-43, :62-65): float ++ or float = float + double */
                if (qa >= p->base_quality && qb >= p->base_quality) *cell = *cell + 1.0f;
                else *cell = (float)((double)*cell + p->edge_weight);
            }
        }
        g0 = g1;
    }
    out->n_far_cells = far.n;
    free(far.keys); free(far.vals);
    free(order); free(gpos); free(gperm); free(gcall); free(node_of); free(vtype);
    return 0;
}

void orc_graph_free(orc_graph *g) {
    free(g->aln_read); free(g->aln_off); free(g->aln_calls); free(g->cnv_start); free(g->cnv_end);
    free(g->node_var); free(g->node_type); free(g->weights);
    memset(g, 0, sizeof(*g));
}

/* ------------------------------------------------------------------------------------------ */
/* sweep + read correction                                                                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t voter; float para, cross, weight; int hap; double esr; } vote_t;
typedef struct { vote_t *a; int n, cap; } votevec;
static void vote_push(votevec *v, vote_t x) {
    if (v->n == v->cap) { v->cap = v->cap * 2 + 8; v->a = (vote_t *)realloc(v->a, sizeof(vote_t) * (size_t)v->cap); }
    v->a[v->n++] = x;
}

int orc_solve(const lps_variants *v, const orc_graph *g, const lps_phase_params *p, orc_solution *out) {
    memset(out, 0, sizeof(*out));
    int nv = v->n, N = g->n_nodes, W = g->window;
    out->n_variants = nv;
    out->ps_sweep = (int32_t *)calloc((size_t)nv + 1, 4); out->ps = (int32_t *)calloc((size_t)nv + 1, 4);
    out->hap_ref_sweep = (int8_t *)malloc((size_t)nv + 1); out->hap_ref = (int8_t *)malloc((size_t)nv + 1);
    memset(out->hap_ref_sweep, -1, (size_t)nv + 1); memset(out->hap_ref, -1, (size_t)nv + 1);
    out->hp_counts = (int32_t *)calloc((size_t)nv * 4 + 4, 4);
    out->n_aln = g->n_aln;
    out->read_hp = (int8_t *)malloc((size_t)g->n_aln + 1);

    /* ---- edgeConnectResult: PhasingGraph.cpp:286-474 ---- */
    int8_t *hp = (int8_t *)calloc((size_t)N + 1, 1);          /* hpResult */
    float *w1 = (float *)calloc((size_t)N + 1, 4), *w2 = (float *)calloc((size_t)N + 1, 4); /* hpCountMap2 */
    votevec *votes = (votevec *)calloc((size_t)N + 1, sizeof(votevec));                      /* hpCountMap3 */
    int32_t *blk_of = (int32_t *)malloc(4 * (size_t)(N + 1));   /* block start node of each pushed member, -2 = not pushed */
    for (int k = 0; k < N; k++) blk_of[k] = -2;
    int block_start = -1, last_connect = -1;
#define NPOS(k) (v->pos[g->node_var[k]])
    for (int k = 0; k + 1 < N; k++) {
        if (abs(NPOS(k + 1) - NPOS(k)) > p->distance) continue;                          /* :318-320 */
        float h1 = w1[k], h2 = w2[k];
        {   /* Onelongcase: :251-283 */
            int counter = 0; float s1 = 0, s2 = 0;
            for (int i = 0; i < votes[k].n; i++) {
                const vote_t *t = &votes[k].a[i];
                if ((t->para + t->cross) <= 1) counter++;
                else if (t->esr < 0.2 && t->weight >= 1 && g->node_type[t->voter] != 3) {
                    if (t->hap == 1) s1 += t->weight; else if (t->hap == 2) s2 += t->weight;
                }
            }
            if (!(counter <= 3 || (s1 == 0 && s2 == 0))) { h1 = s1; h2 = s2; }
        }
        if (h1 == h2) {
            if (last_connect >= 0 && NPOS(k) < NPOS(last_connect)) continue;             /* :340-342 */
            block_start = k; blk_of[k] = k; hp[k] = 1;
        } else {
            hp[k] = (int8_t)(h1 > h2 ? 1 : 2);
            blk_of[k] = block_start;
        }
        int t = k + 1;
        for (int i = 0; i < W; i++) {
            vote_t vt; memset(&vt, 0, sizeof(vt));
            vt.voter = k; vt.weight = 1;
            /* findBestEdgePair: :166-228 */
            const float *c = &g->weights[((size_t)k * (size_t)W + (size_t)(t - k - 1)) * 4];
            float rr = c[0], ra = c[1], ar = c[2], aa = c[3];
            float para = rr + aa, cross = ar + ra;
            double esr = (double)(para < cross ? para : cross) / (double)(para < cross ? cross : para);
            int conn = -1;
            if (rr + aa > ra + ar) conn = 1; else if (rr + aa < ra + ar) conn = 2;
            if (esr > p->edge_threshold) conn = -1;
            if ((esr <= 0.1 && (rr + aa + ra + ar) >= 1) || ((rr + aa) < 1 && (ra + ar) >= 1) || ((rr + aa) >= 1 && (ra + ar) < 1)) vt.weight = 20;
            vt.para = rr + aa; vt.cross = ra + ar; vt.esr = esr;
            if (g->node_type[k] == 4) vt.weight = (float)0.1;                            /* :367-369 */
            if (conn != -1) {
                int same = (conn == 1);
                if (hp[k] == 1) { if (same) { w1[t] += vt.weight; vt.hap = 1; } else { w2[t] += vt.weight; vt.hap = 2; } }
                if (hp[k] == 2) { if (same) { w2[t] += vt.weight; vt.hap = 2; } else { w1[t] += vt.weight; vt.hap = 1; } }
                vote_push(&votes[t], vt);
                last_connect = t;
            }
            t++;
            if (t == N) break;
        }
    }
    /* blocks -> PS / haplotype of the REF allele: :423-467 */
    int8_t *hr = (int8_t *)malloc((size_t)N + 1);
    int32_t *nps = (int32_t *)calloc((size_t)N + 1, 4);
    memset(hr, -1, (size_t)N + 1);
    {
        int prev = -1, prev_blk = -3, size = 0, first = -1;
        for (int k = 0; k <= N; k++) {
            int bk = k < N ? blk_of[k] : -3;
            if (k < N && bk == -2) continue;
            if (bk != prev_blk) { prev_blk = bk; prev = -1; size = 0; first = k; }
            if (k == N) break;
            size++;
            if (prev >= 0) {
                int psv = NPOS(bk) + 1;
                nps[prev] = psv; nps[k] = psv;
                if (prev == first) hr[prev] = 0;
                hr[k] = (int8_t)(hp[prev] == hp[k] ? hr[prev] : 1 - hr[prev]);
            }
            prev = k;
        }
        (void)size;
    }
    for (int k = 0; k < N; k++) { out->ps_sweep[g->node_var[k]] = nps[k]; out->hap_ref_sweep[g->node_var[k]] = nps[k] ? hr[k] : -1; }

    /* ---- readCorrection: :891-1029 ---- */
    int32_t *node_of = (int32_t *)malloc(4 * (size_t)(nv + 1));
    for (int i = 0; i < nv; i++) node_of[i] = -1;
    for (int k = 0; k < N; k++) node_of[g->node_var[k]] = k;
    for (int a = 0; a < g->n_aln; a++) {
        double rc = 0, ac = 0;
        for (uint64_t c = g->aln_off[a]; c < g->aln_off[a + 1]; c++) {
            const lps_call *cl = &g->aln_calls[c];
            int k = node_of[cl->var];
            if (nps[k] == 0) continue;                           /* not in bkResult */
            int h = cl->allele == 0 ? hr[k] : 1 - hr[k];         /* subNodeHP[(pos, allele+1)] */
            double wgt = g->node_type[k] == 0 ? 1.0 : 0.1;       /* types 3,4 add 0.1 (:928-935) */
            if (g->node_type[k] == 0) { if (h == 0) rc++; else ac++; }
            else { if (h == 0) rc += wgt; else ac += wgt; }
        }
        double mx = rc > ac ? rc : ac;
        if (mx / (rc + ac) > p->read_confidence && (rc + ac) > 1) {
            int bh = rc > ac ? 0 : 1;
            out->read_hp[a] = (int8_t)bh;
            for (uint64_t c = g->aln_off[a]; c < g->aln_off[a + 1]; c++) {
                const lps_call *cl = &g->aln_calls[c];
                out->hp_counts[(size_t)cl->var * 4 + (size_t)bh * 2 + (size_t)cl->allele]++;
            }
        } else out->read_hp[a] = -1;
    }
    for (int k = 0; k < N; k++) {
        int vi = g->node_var[k];
        const int32_t *c = &out->hp_counts[(size_t)vi * 4];
        double r1 = (double)c[0] + (double)c[3], r2 = (double)c[2] + (double)c[1];
        double conf = (r1 > r2 ? r1 : r2) / (r1 + r2);
        int h = -1;
        if (conf > p->snp_confidence) { if (r1 > r2) h = 0; else if (r1 < r2) h = 1; }
        out->hap_ref[vi] = (int8_t)h;
        out->ps[vi] = h >= 0 ? nps[k] : 0;
    }
    for (int k = 0; k <= N; k++) free(votes[k].a);
    free(votes); free(hp); free(w1); free(w2); free(blk_of); free(hr); free(nps); free(node_of);
    return 0;
}

void orc_solution_free(orc_solution *s) {
    free(s->ps_sweep); free(s->hap_ref_sweep); free(s->ps); free(s->hap_ref); free(s->read_hp); free(s->hp_counts);
    memset(s, 0, sizeof(*s));
}
