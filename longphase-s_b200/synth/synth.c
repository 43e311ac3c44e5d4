/*
 * synth.c — deterministic ONT-like synthetic workload generator.
 *
 * Produces, for ONE contig: a reference string (iid ACGT with planted homopolymer runs so the
 * D-op/homopolymer rules fire, and planted 2-mer tandem repeats behind some indels so the
 * "danger indel" rule fires), a sorted het SNP/indel table with a random phase, and a
 * coordinate-sorted batch of alignments in the SoA layout of include/lps.h (CIGAR, 4-bit seq,
 * quals, MAPQ, flags) plus uuid-like read names whose lexicographic order differs from the
 * coordinate order.  Everything is a pure function of (seed, parameters): each read draws from
 * its own counter-seeded stream, so the output does not depend on the OpenMP thread count.
 *
 * Built as libsynth.so; used by tests/ and bench.py to make inputs.  It contains no part of the
 * hot path.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    uint64_t seed;
    int64_t contig_len;
    double variant_rate;   /* variants per bp (1e-3 = 1/kb)                              */
    double indel_frac;     /* fraction of variants that are 1-5 bp indels                */
    double danger_frac;    /* fraction of indels planted before a 2-mer x5 tandem repeat */
    double homopolymer_frac; /* fraction of reference bases inside planted homopolymer runs */
    double depth;          /* mean coverage                                              */
    double mean_len;       /* mean read length (ref span)                                */
    double sigma;          /* log-normal sigma                                           */
    double sub_rate, ins_rate, del_rate;
    double clip_frac;      /* reads soft-clipped 6-200 bp (each end independently)        */
    double supp_frac;      /* reads that get a supplementary partner alignment            */
    double sec_frac, dup_frac;
    double noseq_frac;     /* supplementary alignments stored with SEQ '*' (l_qseq = 0)   */
    double lowq_mapq_frac; /* MAPQ 1-59                                                   */
    double zero_mapq_frac; /* MAPQ 0                                                      */
    int32_t tumor;         /* 0: germline reads only                                      */
    double somatic_rate;   /* somatic variants per bp (tumor-only, on one haplotype)      */
    double purity;         /* fraction of reads that come from tumor cells               */
    uint64_t read_seed;    /* seed of the read streams (0: same as seed); reference and variants depend on `seed` only */
} synth_params;

typedef struct {
    /* reference */
    int64_t ref_len;
    char *ref;
    /* variants */
    int32_t n_var;
    int32_t *var_pos;
    uint8_t *var_ref0, *var_alt0;
    uint16_t *var_ref_len, *var_alt_len;
    uint8_t *var_hp1_is_alt;    /* phase: haplotype 0 carries ALT                         */
    uint8_t *var_is_somatic;    /* 1: tumor-only variant, present on haplotype (hp1_is_alt ? 0 : 1) of tumor cells */
    uint32_t *var_str_off;      /* [n_var+1] offsets into var_str: REF '\0' ALT '\0'       */
    char *var_str;
    /* reads */
    int32_t n_reads;
    int32_t *ref_start, *l_qseq;
    uint32_t *n_cigar;
    uint64_t *cigar_off, *seq_off, *qual_off;
    uint16_t *flag;
    uint8_t *mapq;
    int32_t *name_rank;
    uint8_t *hap;               /* truth haplotype of each alignment                      */
    uint32_t *cigar;
    uint64_t cigar_len;
    uint8_t *seq4;
    uint64_t seq_bytes;
    uint8_t *qual;
    uint64_t qual_bytes;
    char *names;                /* n_reads * 40 bytes, NUL terminated 36-char names        */
} synth_out;

/* ---- rng ------------------------------------------------------------------------------- */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t splitmix(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static inline void rng_seed(rng_t *r, uint64_t seed, uint64_t stream) {
    uint64_t x = seed * 0xd1342543de82ef95ULL + stream * 0x2545f4914f6cdd1dULL + 1;
    for (int i = 0; i < 4; i++) r->s[i] = splitmix(&x);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t *r) {
    uint64_t *s = r->s, result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static inline double rng_u(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }
static inline double rng_normal(rng_t *r) {
    double u1 = rng_u(r), u2 = rng_u(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static inline int rng_geom(rng_t *r, double p) { /* >=1, mean 1/p */
    double u = rng_u(r);
    if (u < 1e-300) u = 1e-300;
    int k = 1 + (int)(log(u) / log(1.0 - p));
    return k < 1 ? 1 : k;
}

static const char ACGT[4] = {'A', 'C', 'G', 'T'};
static inline uint8_t nt16(char c) {
    switch (c) { case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; default: return 15; }
}

/* ---- growable per-read buffers ----------------------------------------------------------- */
typedef struct {
    uint32_t *cig; int ncig, ccap;
    uint8_t *bases; uint8_t *quals; int nq, qcap;   /* bases as nt16 codes, one per byte */
} rbuf;
static void rb_reserve_q(rbuf *b, int extra) {
    if (b->nq + extra > b->qcap) {
        b->qcap = (b->nq + extra) * 2 + 64;
        b->bases = (uint8_t *)realloc(b->bases, (size_t)b->qcap);
        b->quals = (uint8_t *)realloc(b->quals, (size_t)b->qcap);
    }
}
static void rb_op(rbuf *b, int op, int len) {
    if (len <= 0) return;
    if (b->ncig && (int)(b->cig[b->ncig - 1] & 15) == op) { b->cig[b->ncig - 1] += (uint32_t)len << 4; return; }
    if (b->ncig == b->ccap) { b->ccap = b->ccap * 2 + 64; b->cig = (uint32_t *)realloc(b->cig, sizeof(uint32_t) * (size_t)b->ccap); }
    b->cig[b->ncig++] = ((uint32_t)len << 4) | (uint32_t)op;
}
/* skewed base-quality table: index by a random byte */
static uint8_t QTAB[256];
static void init_qtab(void) {
    for (int i = 0; i < 256; i++) {
        double u = (i + 0.5) / 256.0;
        /* ~12 % of bases below Q12, bulk between 15 and 40 */
        double q = u < 0.12 ? 3.0 + u / 0.12 * 9.0 : 12.0 + pow((u - 0.12) / 0.88, 0.7) * 28.0;
        QTAB[i] = (uint8_t)q;
    }
}
static inline void rb_base(rbuf *b, rng_t *r, char c) {
    rb_reserve_q(b, 1);
    b->bases[b->nq] = nt16(c);
    b->quals[b->nq] = QTAB[rng_next(r) & 255];
    b->nq++;
}
static inline char other_base(rng_t *r, char c) {
    char o;
    do { o = ACGT[rng_below(r, 4)]; } while (o == c);
    return o;
}

/* walk the reference from `start` for `span` bases on haplotype `hap`, appending M/I/D ops */
static void gen_aligned(const synth_params *p, const synth_out *o, rng_t *r, rbuf *b, int64_t start, int64_t span, int hap, int tumor_cell) {
    int64_t end = start + span;
    if (end > o->ref_len) end = o->ref_len;
    /* first variant >= start */
    int lo = 0, hi = o->n_var;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (o->var_pos[mid] < start) lo = mid + 1; else hi = mid; }
    int vi = lo;
    double perr = p->sub_rate + p->ins_rate + p->del_rate;
    int64_t pos = start;
    int64_t next_err = perr > 0 ? pos + rng_geom(r, perr) - 1 : INT64_MAX;
    while (pos < end) {
        int64_t next_var = vi < o->n_var ? o->var_pos[vi] : INT64_MAX;
        int64_t stop = end;
        if (next_var < stop) stop = next_var;
        if (next_err < stop) stop = next_err;
        /* clean run [pos, stop) */
        if (stop > pos) {
            int n = (int)(stop - pos);
            rb_reserve_q(b, n);
            for (int i = 0; i < n; i++) {
                b->bases[b->nq + i] = nt16(o->ref[pos + i]);
            }
            int i = 0;
            for (; i + 8 <= n; i += 8) {
                uint64_t x = rng_next(r);
                for (int k = 0; k < 8; k++) b->quals[b->nq + i + k] = QTAB[(x >> (8 * k)) & 255];
            }
            if (i < n) { uint64_t x = rng_next(r); for (; i < n; i++) { b->quals[b->nq + i] = QTAB[x & 255]; x >>= 8; } }
            b->nq += n;
            rb_op(b, 0, n);
            pos = stop;
            if (pos >= end) break;
        }
        if (pos == next_var) {
            int carries_alt = (o->var_hp1_is_alt[vi] != 0) == (hap == 0);
            if (o->var_is_somatic[vi] && !tumor_cell) carries_alt = 0;
            int rl = o->var_ref_len[vi], al = o->var_alt_len[vi];
            const char *rs = o->var_str + o->var_str_off[vi];
            const char *as = rs + rl + 1;
            if (rl == 1 && al == 1) {
                char c = carries_alt ? as[0] : rs[0];
                /* sequencing substitution on top of the variant base */
                if (rng_u(r) < p->sub_rate) c = other_base(r, c);
                rb_base(b, r, c); rb_op(b, 0, 1); pos += 1;
            } else if (rl == 1) { /* insertion after the anchor */
                rb_base(b, r, rs[0]); rb_op(b, 0, 1); pos += 1;
                if (carries_alt && pos < end) {
                    for (int k = 1; k < al; k++) rb_base(b, r, as[k]);
                    rb_op(b, 1, al - 1);
                }
            } else { /* deletion after the anchor */
                rb_base(b, r, rs[0]); rb_op(b, 0, 1); pos += 1;
                if (carries_alt && pos + (rl - 1) < end) { rb_op(b, 2, rl - 1); pos += rl - 1; }
            }
            vi++;
            while (vi < o->n_var && o->var_pos[vi] < pos) vi++;   /* variants swallowed by a deletion */
            if (next_err < pos) next_err = pos + rng_geom(r, perr) - 1;
            continue;
        }
        /* error event at pos */
        double u = rng_u(r) * perr;
        if (u < p->sub_rate) {
            rb_base(b, r, other_base(r, o->ref[pos])); rb_op(b, 0, 1); pos += 1;
        } else if (u < p->sub_rate + p->ins_rate) {
            int len = rng_geom(r, 0.6);
            for (int k = 0; k < len; k++) rb_base(b, r, ACGT[rng_below(r, 4)]);
            rb_op(b, 1, len);
            /* an insertion must be followed by an aligned base to stay a sane CIGAR */
            rb_base(b, r, o->ref[pos]); rb_op(b, 0, 1); pos += 1;
        } else {
            int len = rng_geom(r, 0.55);
            if (pos + len >= end) len = (int)(end - pos - 1);
            if (len > 0 && b->ncig > 0) {
                rb_op(b, 2, len); pos += len;
                while (vi < o->n_var && o->var_pos[vi] < pos) vi++;
            }
            if (pos < end) { rb_base(b, r, o->ref[pos]); rb_op(b, 0, 1); pos += 1; }
        }
        while (vi < o->n_var && o->var_pos[vi] < pos) vi++;
        next_err = pos + rng_geom(r, perr) - 1;
    }
}

typedef struct {
    int64_t start; int64_t span; int hap; uint16_t flag; uint8_t mapq; int clip_front, clip_back; int hard; int noseq;
    uint64_t name_id; uint64_t stream; int tumor_cell;
} aln_plan;

static int cmp_plan(const void *a, const void *b) {
    const aln_plan *x = (const aln_plan *)a, *y = (const aln_plan *)b;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    if (x->stream != y->stream) return x->stream < y->stream ? -1 : 1;
    return 0;
}
typedef struct { char name[40]; int32_t idx; } name_ent;
static int cmp_name(const void *a, const void *b) { return strcmp(((const name_ent *)a)->name, ((const name_ent *)b)->name); }

static void make_name(uint64_t seed, uint64_t id, char *out) {
    uint64_t x = seed ^ (id * 0x9e3779b97f4a7c15ULL) ^ 0x5bf03635d1a7c3e1ULL;
    uint64_t a = splitmix(&x), b = splitmix(&x);
    snprintf(out, 40, "%08x-%04x-%04x-%04x-%012llx", (uint32_t)(a >> 32), (uint32_t)((a >> 16) & 0xffff), (uint32_t)(a & 0xffff),
             (uint32_t)(b >> 48), (unsigned long long)(b & 0xffffffffffffULL));
}

void synth_default_params(synth_params *p) {
    memset(p, 0, sizeof(*p));
    p->seed = 1; p->contig_len = 1000000; p->variant_rate = 1e-3; p->indel_frac = 0.0; p->danger_frac = 0.2;
    p->homopolymer_frac = 0.05; p->depth = 30; p->mean_len = 20000; p->sigma = 0.5;
    p->sub_rate = 0.02; p->ins_rate = 0.015; p->del_rate = 0.02;
    p->clip_frac = 0.15; p->supp_frac = 0.03; p->sec_frac = 0.01; p->dup_frac = 0.005; p->noseq_frac = 0.0;
    p->lowq_mapq_frac = 0.07; p->zero_mapq_frac = 0.03;
}

void synth_free(synth_out *o) {
    free(o->ref); free(o->var_pos); free(o->var_ref0); free(o->var_alt0); free(o->var_ref_len); free(o->var_alt_len);
    free(o->var_hp1_is_alt); free(o->var_is_somatic); free(o->var_str_off); free(o->var_str);
    free(o->ref_start); free(o->l_qseq); free(o->n_cigar); free(o->cigar_off); free(o->seq_off); free(o->qual_off);
    free(o->flag); free(o->mapq); free(o->name_rank); free(o->hap); free(o->cigar); free(o->seq4); free(o->qual); free(o->names);
    memset(o, 0, sizeof(*o));
}

int synth_generate(const synth_params *p, synth_out *o) {
    memset(o, 0, sizeof(*o));
    if (!QTAB[255]) init_qtab();
    rng_t r;
    /* ---- reference (stream 1) ---- */
    int64_t L = p->contig_len;
    o->ref_len = L;
    o->ref = (char *)malloc((size_t)L + 1);
    rng_seed(&r, p->seed, 1);
    {
        int64_t i = 0;
        /* a homopolymer run of mean length 6.5 every `gap` bases gives the requested fraction */
        double gap = p->homopolymer_frac > 0 ? 6.5 / p->homopolymer_frac : 1e18;
        while (i < L) {
            int64_t run = (int64_t)(-log(1.0 - rng_u(&r)) * gap) + 1;
            for (int64_t k = 0; k < run && i < L; k++) o->ref[i++] = ACGT[rng_next(&r) >> 62];
            if (i < L) {
                int hl = 3 + (int)rng_below(&r, 8);
                char c = ACGT[rng_next(&r) >> 62];
                for (int k = 0; k < hl && i < L; k++) o->ref[i++] = c;
            }
        }
        o->ref[L] = 0;
    }
    /* ---- variants (stream 2) ---- */
    rng_seed(&r, p->seed, 2);
    {
        int cap = (int)(L * (p->variant_rate + p->somatic_rate) * 1.3) + 64;
        const double total_rate = p->variant_rate + p->somatic_rate;
        o->var_pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
        o->var_ref0 = (uint8_t *)malloc((size_t)cap); o->var_alt0 = (uint8_t *)malloc((size_t)cap);
        o->var_ref_len = (uint16_t *)malloc(2 * (size_t)cap); o->var_alt_len = (uint16_t *)malloc(2 * (size_t)cap);
        o->var_hp1_is_alt = (uint8_t *)malloc((size_t)cap);
        o->var_is_somatic = (uint8_t *)calloc((size_t)cap, 1);
        o->var_str_off = (uint32_t *)malloc(4 * ((size_t)cap + 1));
        o->var_str = (char *)malloc((size_t)cap * 16);
        uint32_t so = 0;
        int n = 0;
        int64_t pos = 50, last_end = 0;
        while (n + 1 < cap) {
            pos += (int64_t)(-log(1.0 - rng_u(&r)) / total_rate) + 1;
            if (pos + 40 >= L) break;
            int is_indel = rng_u(&r) < p->indel_frac;
            o->var_pos[n] = (int32_t)pos;
            o->var_hp1_is_alt[n] = (uint8_t)(rng_next(&r) >> 63);
            o->var_is_somatic[n] = (uint8_t)(p->somatic_rate > 0 && rng_u(&r) < p->somatic_rate / total_rate);
            o->var_str_off[n] = so;
            if (!is_indel) {
                char rc = o->ref[pos], ac = other_base(&r, rc);
                o->var_ref0[n] = (uint8_t)rc; o->var_alt0[n] = (uint8_t)ac; o->var_ref_len[n] = 1; o->var_alt_len[n] = 1;
                o->var_str[so++] = rc; o->var_str[so++] = 0; o->var_str[so++] = ac; o->var_str[so++] = 0;
                n++;
                /* occasionally put a second SNP 1-2 bp away inside a planted homopolymer so that
                   SnpParser::filterSNP has something to erase */
                if (rng_u(&r) < 0.02 && n < cap && pos - 3 > last_end) {
                    for (int k = -2; k <= 5; k++) o->ref[pos + k] = rc;
                    int d = 1 + (int)rng_below(&r, 2);
                    char ac2 = other_base(&r, rc);
                    o->var_pos[n] = (int32_t)(pos + d);
                    o->var_hp1_is_alt[n] = (uint8_t)(rng_next(&r) >> 63);
                    o->var_str_off[n] = so;
                    o->var_ref0[n] = (uint8_t)rc; o->var_alt0[n] = (uint8_t)ac2; o->var_ref_len[n] = 1; o->var_alt_len[n] = 1;
                    o->var_str[so++] = rc; o->var_str[so++] = 0; o->var_str[so++] = ac2; o->var_str[so++] = 0;
                    n++;
                    pos += 6;
                }
                last_end = pos;
                continue;
            }
            int len = 1 + (int)rng_below(&r, 5);
            int danger = rng_u(&r) < p->danger_frac;
            if (danger) { /* plant XYXYXYXYXY behind the anchor */
                char x = ACGT[rng_below(&r, 4)], y = other_base(&r, x);
                for (int k = 0; k < 12; k++) o->ref[pos + 1 + k] = (k & 1) ? y : x;
            }
            if (rng_next(&r) >> 63) { /* insertion */
                o->var_ref0[n] = (uint8_t)o->ref[pos]; o->var_alt0[n] = (uint8_t)o->ref[pos];
                o->var_ref_len[n] = 1; o->var_alt_len[n] = (uint16_t)(1 + len);
                o->var_str[so++] = o->ref[pos]; o->var_str[so++] = 0;
                o->var_str[so++] = o->ref[pos];
                for (int k = 0; k < len; k++) o->var_str[so++] = danger ? o->ref[pos + 1 + (k & 1)] : ACGT[rng_below(&r, 4)];
                o->var_str[so++] = 0;
            } else {          /* deletion */
                o->var_ref0[n] = (uint8_t)o->ref[pos]; o->var_alt0[n] = (uint8_t)o->ref[pos];
                o->var_ref_len[n] = (uint16_t)(1 + len); o->var_alt_len[n] = 1;
                for (int k = 0; k <= len; k++) o->var_str[so++] = o->ref[pos + k];
                o->var_str[so++] = 0; o->var_str[so++] = o->ref[pos]; o->var_str[so++] = 0;
                pos += len;   /* keep the deleted bases free of other variants */
            }
            n++;
            last_end = pos + 13;
        }
        /* refresh REF chars of SNPs that a later tandem-repeat plant may have overwritten */
        for (int i = 0; i < n; i++) {
            if (o->var_ref_len[i] == 1 && o->var_alt_len[i] == 1) {
                char rc = o->ref[o->var_pos[i]];
                char *s = o->var_str + o->var_str_off[i];
                if (s[0] != rc) { s[0] = rc; o->var_ref0[i] = (uint8_t)rc; if (s[2] == rc) { s[2] = (rc == 'A') ? 'C' : 'A'; o->var_alt0[i] = (uint8_t)s[2]; } }
            } else {
                char *s = o->var_str + o->var_str_off[i];
                int rl = o->var_ref_len[i];
                for (int k = 0; k < rl; k++) s[k] = o->ref[o->var_pos[i] + k];
                s[rl + 1] = s[0];
                o->var_ref0[i] = (uint8_t)s[0]; o->var_alt0[i] = (uint8_t)s[0];
            }
        }
        o->var_str_off[n] = so;
        o->n_var = n;
    }
    /* ---- alignment plan (stream 3) ---- */
    const uint64_t rseed = p->read_seed ? p->read_seed : p->seed;
    rng_seed(&r, rseed, 3);
    double mu = log(p->mean_len) - 0.5 * p->sigma * p->sigma;
    int64_t n_primary = (int64_t)(p->depth * (double)L / p->mean_len);
    if (n_primary < 1) n_primary = 1;
    int64_t cap = n_primary + (int64_t)(n_primary * (p->supp_frac + 0.01)) + 16;
    aln_plan *plan = (aln_plan *)malloc(sizeof(aln_plan) * (size_t)cap);
    int64_t na = 0;
    for (int64_t i = 0; i < n_primary; i++) {
        aln_plan a; memset(&a, 0, sizeof(a));
        double len = exp(mu + p->sigma * rng_normal(&r));
        if (len < 500) len = 500;
        if (len > L / 2) len = (double)(L / 2);
        a.span = (int64_t)len;
        a.start = (int64_t)(rng_u(&r) * (double)(L - a.span));
        a.hap = (int)(rng_next(&r) >> 63);
        a.tumor_cell = p->purity > 0 && rng_u(&r) < p->purity;
        double u = rng_u(&r);
        a.mapq = u < p->zero_mapq_frac ? 0 : (u < p->zero_mapq_frac + p->lowq_mapq_frac ? (uint8_t)(1 + rng_below(&r, 59)) : 60);
        a.flag = (rng_next(&r) >> 63) ? 16 : 0;
        u = rng_u(&r);
        if (u < p->sec_frac) a.flag |= 0x100;
        else if (u < p->sec_frac + p->dup_frac) a.flag |= 0x400;
        if (rng_u(&r) < p->clip_frac) a.clip_front = 6 + (int)rng_below(&r, 195);
        if (rng_u(&r) < p->clip_frac) a.clip_back = 6 + (int)rng_below(&r, 195);
        /* short clips (<=5) that must NOT be counted */
        if (!a.clip_front && rng_u(&r) < 0.05) a.clip_front = 1 + (int)rng_below(&r, 5);
        a.name_id = (uint64_t)i; a.stream = 16 + (uint64_t)na;
        plan[na++] = a;
        if (rng_u(&r) < p->supp_frac && na < cap) {
            /* chimeric partner with the same name: adjacent, partially overlapping or contained */
            aln_plan s = a;
            s.flag = (uint16_t)((a.flag & 16) | 0x800);
            double m = rng_u(&r);
            if (m < 0.4) { s.start = a.start + a.span + (int64_t)rng_below(&r, 3000); s.span = a.span / 3 + 500; }
            else if (m < 0.7) { s.start = a.start + (int64_t)((double)a.span * (0.6 + 0.35 * rng_u(&r))); s.span = a.span / 2 + 500; }
            else if (m < 0.85) { s.start = a.start + a.span / 4; s.span = a.span / 3 + 200; }
            else { s.start = a.start + (int64_t)rng_below(&r, 2000); s.span = a.span + 4000; }
            if (s.start + s.span >= L) s.span = L - s.start - 1;
            if (s.span > 300) {
                s.hard = 1; s.clip_front = 50 + (int)rng_below(&r, 3000); s.clip_back = rng_u(&r) < 0.5 ? 20 + (int)rng_below(&r, 500) : 0;
                s.noseq = rng_u(&r) < p->noseq_frac;
                s.stream = 16 + (uint64_t)na;
                plan[na++] = s;
            }
        }
    }
    qsort(plan, (size_t)na, sizeof(aln_plan), cmp_plan);
    int32_t n = (int32_t)na;
    o->n_reads = n;
    o->ref_start = (int32_t *)malloc(4 * (size_t)n); o->l_qseq = (int32_t *)malloc(4 * (size_t)n);
    o->n_cigar = (uint32_t *)malloc(4 * (size_t)n);
    o->cigar_off = (uint64_t *)malloc(8 * (size_t)n); o->seq_off = (uint64_t *)malloc(8 * (size_t)n); o->qual_off = (uint64_t *)malloc(8 * (size_t)n);
    o->flag = (uint16_t *)malloc(2 * (size_t)n); o->mapq = (uint8_t *)malloc((size_t)n);
    o->name_rank = (int32_t *)malloc(4 * (size_t)n); o->hap = (uint8_t *)malloc((size_t)n);
    o->names = (char *)calloc((size_t)n, 40);

    /* ---- generate alignments in parallel into per-read buffers ---- */
    rbuf *bufs = (rbuf *)calloc((size_t)n, sizeof(rbuf));
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t i = 0; i < n; i++) {
        const aln_plan *a = &plan[i];
        rng_t rr; rng_seed(&rr, rseed, a->stream);
        rbuf *b = &bufs[i];
        int clip_op = a->hard ? 5 : 4;
        if (a->clip_front) {
            rb_op(b, clip_op, a->clip_front);
            if (!a->hard) for (int k = 0; k < a->clip_front; k++) rb_base(b, &rr, ACGT[rng_below(&rr, 4)]);
        }
        gen_aligned(p, o, &rr, b, a->start, a->span, a->hap, a->tumor_cell);
        if (a->clip_back) {
            rb_op(b, clip_op, a->clip_back);
            if (!a->hard) for (int k = 0; k < a->clip_back; k++) rb_base(b, &rr, ACGT[rng_below(&rr, 4)]);
        }
        if (a->noseq) b->nq = 0;
    }
    uint64_t co = 0, so = 0, qo = 0;
    for (int32_t i = 0; i < n; i++) {
        o->cigar_off[i] = co; o->seq_off[i] = so; o->qual_off[i] = qo;
        co += (uint64_t)bufs[i].ncig; so += (uint64_t)(bufs[i].nq + 1) / 2; qo += (uint64_t)bufs[i].nq;
    }
    o->cigar_len = co; o->seq_bytes = so; o->qual_bytes = qo;
    o->cigar = (uint32_t *)malloc(4 * (size_t)co + 64); o->seq4 = (uint8_t *)calloc((size_t)so + 64, 1); o->qual = (uint8_t *)malloc((size_t)qo + 64);
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t i = 0; i < n; i++) {
        const aln_plan *a = &plan[i];
        rbuf *b = &bufs[i];
        o->ref_start[i] = (int32_t)a->start; o->l_qseq[i] = b->nq; o->n_cigar[i] = (uint32_t)b->ncig;
        o->flag[i] = a->flag; o->mapq[i] = a->mapq; o->hap[i] = (uint8_t)a->hap;
        memcpy(o->cigar + o->cigar_off[i], b->cig, 4 * (size_t)b->ncig);
        memcpy(o->qual + o->qual_off[i], b->quals, (size_t)b->nq);
        uint8_t *s = o->seq4 + o->seq_off[i];
        for (int k = 0; k + 1 < b->nq; k += 2) s[k >> 1] = (uint8_t)((b->bases[k] << 4) | b->bases[k + 1]);
        if (b->nq & 1) s[b->nq >> 1] = (uint8_t)(b->bases[b->nq - 1] << 4);
        make_name(rseed, a->name_id, o->names + (size_t)i * 40);
        free(b->cig); free(b->bases); free(b->quals);
    }
    free(bufs);
    /* ---- name ranks (std::string operator< == strcmp order for these ASCII names) ---- */
    name_ent *ne = (name_ent *)malloc(sizeof(name_ent) * (size_t)n);
    for (int32_t i = 0; i < n; i++) { memcpy(ne[i].name, o->names + (size_t)i * 40, 40); ne[i].idx = i; }
    qsort(ne, (size_t)n, sizeof(name_ent), cmp_name);
    int32_t rank = -1;
    for (int32_t i = 0; i < n; i++) {
        if (i == 0 || strcmp(ne[i].name, ne[i - 1].name) != 0) rank++;
        o->name_rank[ne[i].idx] = rank;
    }
    free(ne);
    free(plan);
    return 0;
}
