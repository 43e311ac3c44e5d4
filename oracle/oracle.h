/*
 * oracle.h — CPU restatement of the LongPhase-S read-to-variant hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load liboracle.so; the product (longphase-s_b200/) never does.
 *
 * Parity status: PINNED against the unmodified reference compiled in oracle/_ref
 * (oracle/ref_tap.cpp drives the reference's own get_snp / filterSNP / Clip / addEdge /
 * edgeConnectResult / readCorrection / exportResult on the same inputs; see
 * tests/test_oracle_vs_reference.py and the fixtures in tests/golden/).  The reference ships no
 * golden vectors of its own.
 *
 * Every function cites the reference file:line it restates (paths relative to the reference
 * tree).  The code is written from the behavioural rule sheets of SURVEY.md Appendix A on the
 * SoA layout of include/lps.h; it is not a copy of the reference.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#include "../include/lps.h"

#ifdef __cplusplus
extern "C" {
#endif

/* per-variant notes: homopolymerLength (src/shared/Util.cpp:21-54), is_danger
 * (src/phase/ParsingBam.cpp:378-417), filterSNP erase set (src/phase/ParsingBam.cpp:866-888) */
int orc_annotate(const char *ref, int64_t ref_len, const lps_variants *v, int is_ont,
                 uint8_t *hom, uint8_t *danger, uint8_t *filtered);

/* growable result of orc_call_alleles */
typedef struct {
    int32_t n_reads;
    uint64_t n_calls;
    uint64_t *call_off;     /* [n_reads+1] */
    lps_call *calls;
    uint8_t *read_status;
    int32_t n_clips;
    int32_t *clip_pos, *clip_front, *clip_back;
} orc_calls;

/* BamParser::direct_detect_alleles filter + get_snp + getClip (src/phase/ParsingBam.cpp:1243-1645);
 * apply_filter!=0 also drops calls at `filtered` variants (filterSNP, :891-911).               */
int orc_call_alleles(const lps_read_batch *b, const lps_variants *v, const uint8_t *hom, const uint8_t *danger,
                     const uint8_t *filtered, int apply_filter, const lps_phase_params *p, orc_calls *out);
void orc_calls_free(orc_calls *c);

typedef struct {
    /* stage C: alignments surviving the overlap filter, calls surviving the CNV filter */
    int32_t n_aln;
    int32_t *aln_read;      /* batch index of each surviving alignment */
    uint64_t *aln_off;      /* [n_aln+1] */
    lps_call *aln_calls;
    int32_t n_cnv;
    int32_t *cnv_start, *cnv_end;
    /* graph */
    int32_t n_nodes;
    int32_t *node_var;
    uint8_t *node_type;
    int32_t window;
    float *weights;         /* [n_nodes][window][4] */
    uint64_t n_contrib, n_contrib_far;
    uint64_t n_far_cells;   /* distinct (a,b,allele pair) cells beyond the window */
} orc_graph;

/* VairiantGraph::addEdge (src/phase/PhasingGraph.cpp:694-889) incl. Clip::getCNVInterval run twice
 * (:1103-1106, :1128-1227; src/phase/PhasingProcess.cpp:147-148) and SubEdge::addSubEdge (:25-70) */
int orc_build_graph(const lps_read_batch *b, const lps_variants *v, const uint8_t *danger, const orc_calls *calls,
                    const lps_phase_params *p, orc_graph *out);
void orc_graph_free(orc_graph *g);

typedef struct {
    int32_t n_variants;
    int32_t *ps_sweep;      /* after edgeConnectResult */
    int8_t *hap_ref_sweep;
    int32_t *ps;            /* after readCorrection */
    int8_t *hap_ref;
    int32_t n_aln;
    int8_t *read_hp;        /* per stage-C alignment */
    int32_t *hp_counts;     /* [n_variants][4] */
} orc_solution;

/* VairiantGraph::edgeConnectResult (:286-474), findBestEdgePair (:166-228), Onelongcase (:251-283),
 * readCorrection (:891-1029), exportResult (:1049-1077)                                          */
int orc_solve(const lps_variants *v, const orc_graph *g, const lps_phase_params *p, orc_solution *out);
void orc_solution_free(orc_solution *s);

/* germline haplotag: dispatch + CigarParser::parsingCigar + GermlineHaplotagStrategy (see oracle_tag.c) */
typedef struct {
    int32_t n_reads;
    uint8_t *category;
    int8_t *hp;
    int32_t *ps, *pq, *h1, *h2;
    uint64_t n_calls;
    uint64_t *call_off;
    lps_call *calls;       /* variants that touched countPS: allele = variantsHP (0/1) or -1, origin 0 M-SNP, 1 D-SNP, 2 indel */
} orc_tags;
int orc_tag_reads(const lps_read_batch *b, const lps_variants *v, const uint8_t *hom, const lps_tag_params *p, orc_tags *out);
void orc_tags_free(orc_tags *t);

/* somatic family (oracle_somatic.c): the two extract passes of SomaticVarCaller::extractSomaticData and somatic tagging.
 * Per-position arrays are indexed by tumor slot (ascending tumor-present variants), as in include/lps.h.            */
enum { ORC_SOM_EXTRACT_NORMAL = 0, ORC_SOM_EXTRACT_TUMOR = 1, ORC_SOM_TAG = 2 };
typedef struct {
    int32_t n_reads, n_tum;
    int32_t *tum_var;
    uint8_t *category;
    int8_t *read_hp, *hp_before;
    int32_t *ps, *pq, *h1, *h2, *h3;
    uint8_t *n_ps;
    int32_t *end_pos, *read_len;
    float *derive_similarity;
    int32_t *pos_base, *read_hp_count, *somatic_read_hp_count, *case_count, *allele_count, *window_hist;
    int32_t *hp_before_count, *hp_after_count, *h3_before_count, *h3_after_count, *cover_start, *cover_end;
    float *ratios_f;         /* [n_tum][LPS_RF_FIELDS] postProcess (extract passes) */
    double *ratios_d;        /* [n_tum][LPS_RD_FIELDS] */
    int32_t *case_read_count;
    uint64_t n_window_items;
    uint64_t n_calls;
    uint64_t *call_off;
    lps_call *calls;
} orc_somatic_out;
int orc_somatic(int mode, const lps_read_batch *b, const lps_variants *v, const lps_tumor_variants *t, const uint8_t *hom,
                const char *ref, int64_t ref_len, const lps_tag_params *p, orc_somatic_out *out);
void orc_somatic_free(orc_somatic_out *o);

/* std::sort with the reference's comparator (src/shared/Util.h:100-106, Util.cpp:3-5): the order of
 * equal positions inside a merged read is whatever libstdc++'s introsort leaves, so the oracle calls
 * the same std::sort.  perm[] is permuted alongside pos[].                                        */
void orc_std_sort_by_pos(int32_t *pos, int32_t *perm, int32_t n);

#ifdef __cplusplus
}
#endif
#endif
