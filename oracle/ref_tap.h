/* ref_tap.h — C view of the reference tap library (TEST INFRASTRUCTURE ONLY, see ref_tap.cpp). */
#ifndef REF_TAP_H
#define REF_TAP_H
#include <stdint.h>
#include "../include/lps.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const char *chr;
    const char *ref;
    int64_t ref_len;
    int32_t n_var;
    const int32_t *var_pos;
    const uint32_t *var_str_off;   /* REF '\0' ALT '\0' per variant                       */
    const char *var_str;
    lps_read_batch batch;
    const char *names;             /* NUL-terminated, one per read, name_stride apart      */
    int32_t name_stride;
    int32_t stop_after_calls;      /* 1: only get_snp + filterSNP                          */
    lps_phase_params p;
} tap_phase_in;

/* a std::vector<ReadVariant> flattened: alignment k holds calls [off[k], off[k+1]) */
typedef struct {
    int32_t n_aln;
    int32_t *read_idx;             /* index of the alignment in the input batch            */
    uint64_t *off;
    int32_t *pos, *allele, *quality;
} tap_calls;

/* one entry per key of VairiantGraph::totalVariantInfo */
typedef struct {
    int32_t n;
    int32_t *pos, *type;
    int32_t *ps;                   /* bkResult[(pos,1)] or 0                               */
    int32_t *hap_ref, *hap_alt;    /* subNodeHP[(pos,1)], [(pos,2)], -1 when absent        */
} tap_nodes;

typedef struct {
    tap_calls stage_a;             /* after get_snp                                        */
    tap_calls stage_b;             /* after SnpParser::filterSNP                           */
    tap_calls stage_c;             /* after the filters at the head of addEdge             */
    int32_t n_clips;
    int32_t *clip_pos, *clip_front, *clip_back;
    int32_t n_cnv;
    int32_t *cnv_start, *cnv_end;
    int32_t n_empty_after_filter;  /* alignments whose calls were all erased by filterSNP  */
    int32_t n_edge_nodes;
    int64_t n_cells;
    uint64_t n_contrib;            /* sum of SubEdge::readCount                            */
    int32_t *cell_a, *cell_b;      /* positions                                            */
    uint8_t *cell_which;           /* 0 rr, 1 ra, 2 ar, 3 aa                               */
    float *cell_val;
    tap_nodes nodes_sweep;         /* after edgeConnectResult                              */
    tap_nodes nodes_final;         /* after readCorrection                                 */
    int32_t *read_hp;              /* per stage_c alignment: readHpMap[name]               */
    int32_t n_result;              /* exportResult                                         */
    int32_t *res_pos, *res_block, *res_hap_ref, *res_hap_alt;
    double t_get_snp, t_filter_snp, t_clip, t_add_edge, t_sweep, t_read_correction;
    double t_export;               /* exportResult                                          */
    int64_t n_result_calls;        /* timing mode (stop_after_calls == 2): calls after get_snp */
} tap_phase_out;

/* ---- germline haplotag tap (ref_tap_tag.cpp) ---- */
typedef struct {
    const char *chr;
    const char *ref;
    int64_t ref_len;
    int32_t n_var;
    const int32_t *var_pos;
    const uint32_t *var_str_off;
    const char *var_str;
    const uint8_t *var_hp1_is_alt;
    const int32_t *var_ps;
    lps_read_batch batch;
    const char *names;
    int32_t name_stride;
    lps_tag_params p;
} tap_tag_in;

typedef struct {
    int32_t n_reads;
    uint8_t *category;
    int32_t *hp, *ps, *pq, *h1, *h2, *n_ps;   /* per read: judgeHaplotype result and hpCount / countPS.size() */
    uint64_t *var_off;                        /* per read CSR over variantsHP: (position, 0|1) */
    int32_t *var_pos, *var_hp;
    uint64_t *ps_off;                         /* per read CSR over countPS: (PS id, count) */
    int32_t *ps_id, *ps_count;
    int64_t stats[14];                        /* ReadStatistics in the order of lps_tag_result's counters */
    double t_total;
} tap_tag_out;

int ref_tap_tag(const tap_tag_in *in, tap_tag_out *out);
void ref_tap_tag_free(tap_tag_out *out);

/* ---- somatic family tap (ref_tap_somatic.cpp); the output is oracle.h's orc_somatic_out ---- */
#define TAP_SOM_STATS 22   /* the 13 ReadStatistics counters of lps_somatic_tag_result followed by totalHpCount[0..8] */
typedef struct {
    const char *chr;
    const char *ref;
    int64_t ref_len;
    int32_t n_var;
    const int32_t *var_pos;
    const uint32_t *var_str_off;     /* NORMAL record: REF '\0' ALT '\0' */
    const char *var_str;
    const uint8_t *var_hp1_is_alt;
    const int32_t *var_ps;
    const uint8_t *nor_gt;           /* NULL: every NORMAL record is PHASED_HETERO */
    const uint8_t *nor_present;      /* NULL: everywhere */
    const uint8_t *tum_present;
    const uint32_t *tum_str_off;     /* TUMOR record strings, same layout */
    const char *tum_str;
    const uint8_t *tum_gt, *tum_hp1_is_alt;
    const int32_t *tum_ps;
    const uint8_t *is_somatic;
    const int8_t *derive_hp;
    lps_read_batch batch;
    const char *names;
    int32_t name_stride;
    lps_tag_params p;
    int64_t *stats_out;              /* [TAP_SOM_STATS], written in mode 2 */
    void *keep;                      /* optional ref_tap_state: modes 0 / 1 leave the reference's chrPosNorBase / chrPosSomaticInfo
                                        (and, mode 1, readHpResultSet / tumorPosReadCorrBaseHP with the key of every alignment) there */
} tap_som_in;

/* the reference's TumorPurityEstimator on the maps two extract passes left in a state object */
typedef struct {
    double purity;                   /* TumorPurityEstimator::estimateTumorPurity() */
    int32_t threshold, n_after_lcvf, n_used;
    double median, q1, q3, iqr, lower_whisker, upper_whisker;
} tap_purity_out;
/* the reference's calling stage between the extract passes and the tagging pass (SomaticVarCaller::variantCalling,
 * src/somatic_haplotag/SomaticVarCaller.cpp:816-866 + getSomaticFlag :2397-2412) on the maps a state object holds for `chr`;
 * per tumor slot (same slot order as ref_tap_somatic), per alignment of the tumor batch                                   */
typedef struct {
    int32_t n_tum, n_reads;
    uint8_t *touched;                /* the position has a SomaticData entry                                               */
    float *mean_alt, *z_score;       /* SomaticData::meanAltCountPerVarRead, zScore                                        */
    int32_t *interval_snp_count, *min_distance, *dense_alt_same;
    uint8_t *in_dense, *filtered_by; /* filtered_by: [n_tum][6] TINC, MessyRead, ReadCount, HapConsistency, VariantCluster, DenseAlt */
    uint8_t *is_filter_out, *high_con;
    int32_t *derive_hp;              /* SomaticData::somaticReadDeriveByHP                                                 */
    uint8_t *is_somatic; int32_t *flag_derive_hp;   /* what getSomaticFlag writes into the variant map                     */
    int8_t *read_hp;                 /* ReadVarHpCount::hpResult after calculateReadSetHP, -1 for alignments without a record */
    int32_t *read_h3;                /* ReadVarHpCount::HP3 after calibrateReadHP                                          */
    int32_t tier;
} tap_call_out;
int ref_tap_somatic_call(void *state, const tap_som_in *in, double purity, int enable_filter, tap_call_out *out);
void ref_tap_somatic_call_free(tap_call_out *out);
void *ref_tap_state_new(void);
void ref_tap_state_free(void *state);
int ref_tap_purity(void *state, const char *chr, tap_purity_out *out);

int ref_tap_phase(const tap_phase_in *in, tap_phase_out *out);
void ref_tap_phase_free(tap_phase_out *out);
int ref_tap_homopolymer(const char *ref, int64_t len, int pos);

#ifdef __cplusplus
}
#endif
#endif
