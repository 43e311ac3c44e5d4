// host_somatic_call.cpp — the calling stage between the two extract passes and the tagging pass of `somatic_haplotag`
// (SURVEY §8f rank 3): SomaticVarCaller::variantCalling minus extraction and logs (reference
// src/somatic_haplotag/SomaticVarCaller.cpp:816-866) and getSomaticFlag (:2397-2412).  O(tumor positions + their reads), host
// arithmetic in the reference's own float / double / int types; it consumes the per-position counters and per-alignment
// products the extract kernels produced (lps_extract_result) and yields the isSomaticVariant / somaticReadDeriveByHP flags
// lps_contig_set_tumor_variants takes for lps_somatic_tag_reads.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <vector>
#include "lps_ctx.cuh"

namespace {

enum { HP_UNTAG = 0, HP_H1 = 1, HP_H2 = 2, HP_H3 = 3, HP_H4 = 4, HP_H1_1 = 5, HP_H1_2 = 6, HP_H2_1 = 7, HP_H2_2 = 8 };   // ReadHP
enum { SNP_NONE = 0, SNP_H1 = 1, SNP_H2 = 2, SNP_H3 = 3, SNP_H4 = 4 };                                                       // SnpHP

// SomaticVarFilterParams after setFilterParamsWithPurity (:951-1060): the reference stores the tier constants in float / int
// members, so the double literals are narrowed exactly like there
struct FilterParams {
    float norVAF_maxThr; int norDepth_minThr;
    float MessyReadRatioThreshold; int ReadCount_minThr;
    float HapConsistency_VAF_maxThr; int HapConsistency_ReadCount_maxThr, HapConsistency_somaticRead_minThr;
    float IntervalSnpCount_VAF_maxThr; int IntervalSnpCount_ReadCount_maxThr, IntervalSnpCount_minThr;
    float zScore_maxThr;
    float DenseAlt_condition1_thr = 0.5, DenseAlt_condition2_thr = 0.6; int DenseAlt_sameCount_minThr = 3;   // SomaticVarCaller.h:101-103
};

int tier_of(double purity) {
    if (purity >= 0.9 && purity <= 1.0) return 1;
    if (purity >= 0.7 && purity < 0.9) return 2;
    if (purity >= 0.5 && purity < 0.7) return 3;
    if (purity >= 0.3 && purity < 0.5) return 4;
    return 5;
}

FilterParams params_of(int tier) {
    FilterParams p;
    auto set = [&](double nv, double rc, double hc_rc, double hc_vaf, double hc_sr, double is_rc, double is_vaf, double is_min, double z) {
        p.norVAF_maxThr = nv; p.norDepth_minThr = 1; p.MessyReadRatioThreshold = 1.0; p.ReadCount_minThr = rc;
        p.HapConsistency_ReadCount_maxThr = hc_rc; p.HapConsistency_VAF_maxThr = hc_vaf; p.HapConsistency_somaticRead_minThr = hc_sr;
        p.IntervalSnpCount_ReadCount_maxThr = is_rc; p.IntervalSnpCount_VAF_maxThr = is_vaf; p.IntervalSnpCount_minThr = is_min;
        p.zScore_maxThr = z;
    };
    switch (tier) {
        case 1: set(0.13, 3.0, 12.0, 0.144, 0.0, 12.0, 0.189, 4.0, 5.233); break;
        case 2: set(0.13, 3.0, 10.0, 0.130, 1.0, 10.0, 0.133, 4.0, 2.676); break;
        case 3: set(0.105, 1.0, 10.0, 0.071, 0.0, 10.0, 0.105, 4.0, 5.683); break;
        case 4: set(0.117, 1.0, 8.0, 0.035, 1.0, 8.0, 0.049, 4.0, 3.043); break;
        default: set(0.130, 1.0, 8.0, 0.020, 1.0, 8.0, 0.025, 8.0, 1.953); break;
    }
    return p;
}

// SomaticJudgeHapStrategy::judgeSomaticReadHap (src/haplotag/HaplotagStrategy.cpp:452-602), the haplotype only
int judge_somatic_read_hap(int hp1, int hp2, int hp3, int hp4, int n_ps, double pct) {
    double tumMin, tumMax, norMin, norMax;
    int maxTum, maxNor;
    if (hp3 > hp4) { tumMin = hp4; tumMax = hp3; maxTum = SNP_H3; } else { tumMin = hp3; tumMax = hp4; maxTum = SNP_H4; }
    if (hp1 > hp2) { norMin = hp2; norMax = hp1; maxNor = SNP_H1; } else { norMin = hp1; norMax = hp2; maxNor = SNP_H2; }
    const double tumSim = tumMax == 0 ? 0.0 : tumMax / (tumMax + tumMin), norSim = norMax == 0 ? 0.0 : norMax / (norMax + norMin);
    int hp = HP_UNTAG;
    if (tumMax != 0) {
        if (tumSim >= pct) {
            if (norSim >= pct) hp = maxTum == SNP_H3 ? (maxNor == SNP_H1 ? HP_H1_1 : HP_H2_1) : (maxNor == SNP_H1 ? HP_H1_2 : HP_H2_2);
            else hp = maxTum == SNP_H3 ? HP_H3 : HP_H4;
        }
    } else if (norMax != 0) {
        if (norSim >= pct) hp = maxNor;
    }
    if (n_ps > 1) hp = HP_UNTAG;                       // the read crosses two phase sets (:561-574)
    return hp;
}

}  // namespace

int lps_somatic_filter_params_of(double purity, lps_somatic_filter_params *out) {
    if (!out) return LPS_E_ARG;
    const int tier = tier_of(purity);
    const FilterParams p = params_of(tier);
    out->tier = tier; out->tumor_purity = (float)purity; out->nor_vaf_max = p.norVAF_maxThr; out->nor_depth_min = p.norDepth_minThr;
    out->messy_read_ratio = p.MessyReadRatioThreshold; out->read_count_min = p.ReadCount_minThr;
    out->hap_consistency_vaf_max = p.HapConsistency_VAF_maxThr; out->hap_consistency_read_count_max = p.HapConsistency_ReadCount_maxThr;
    out->hap_consistency_somatic_read_min = p.HapConsistency_somaticRead_minThr;
    out->interval_snp_count_vaf_max = p.IntervalSnpCount_VAF_maxThr; out->interval_snp_count_read_count_max = p.IntervalSnpCount_ReadCount_maxThr;
    out->interval_snp_count_min = p.IntervalSnpCount_minThr; out->z_score_max = p.zScore_maxThr;
    out->dense_alt_condition1 = p.DenseAlt_condition1_thr; out->dense_alt_condition2 = p.DenseAlt_condition2_thr;
    out->dense_alt_same_count_min = p.DenseAlt_sameCount_minThr;
    return LPS_OK;
}

int lps_somatic_call(const lps_somatic_call_input *in, lps_somatic_call_result *out) {
    if (!in || !out || !in->normal || !in->tumor || in->n_tum < 0) return LPS_E_ARG;
    const lps_extract_result &N = *in->normal, &T = *in->tumor;
    const int nt = in->n_tum, nr = T.reads.n_reads;
    if (N.n_tum != nt || T.n_tum != nt || (nt && (!in->pos || !in->callable || !T.tum_var || !T.pos_base || !T.read_hp_count ||
                                                 !T.somatic_read_hp_count || !T.ratios_f || !T.case_read_count || !T.window_hist ||
                                                 !N.pos_base || !N.ratios_f)) ||
        (nr && (!T.reads.h1 || !T.reads.h2 || !T.reads.h3 || !T.reads.n_ps || !T.call_off || (T.n_calls && !T.calls))))
        return LPS_E_ARG;
    out->tier = tier_of(in->purity);
    const FilterParams fp = params_of(out->tier);

    // ---- readHpResultSet / tumorPosReadCorrBaseHP (:407-459) rebuilt from the per-alignment call lists ----
    std::vector<int32_t> slot_of_var;                                  // variant index -> tumor slot
    {
        int32_t maxv = -1;
        for (int k = 0; k < nt; k++) maxv = std::max(maxv, T.tum_var[k]);
        slot_of_var.assign((size_t)maxv + 2, -1);
        for (int k = 0; k < nt; k++) slot_of_var[(size_t)T.tum_var[k]] = k;
    }
    std::vector<uint8_t> in_set((size_t)nr, 0);                        // the alignment has a ReadVarHpCount record
    std::vector<int32_t> hp3((size_t)nr, 0);
    std::vector<uint32_t> cnt((size_t)nt + 1, 0);
    for (int r = 0; r < nr; r++)
        for (uint64_t c = T.call_off[r]; c < T.call_off[r + 1]; c++)
            if (T.calls[c].quality & 1) {                              // the position was in the read's tumorSnpPosVec
                const int32_t v = T.calls[c].var;
                if (v < 0 || (size_t)v >= slot_of_var.size() || slot_of_var[(size_t)v] < 0) return LPS_E_ARG;
                in_set[(size_t)r] = 1; cnt[(size_t)slot_of_var[(size_t)v] + 1]++;
            }
    for (int k = 0; k < nt; k++) cnt[(size_t)k + 1] += cnt[(size_t)k];
    std::vector<int32_t> pr_read(cnt[(size_t)nt]);                     // CSR by slot: alignments, in batch order
    std::vector<int8_t> pr_hp(cnt[(size_t)nt]);                        //              their base haplotype at the position
    {
        std::vector<uint32_t> fill(cnt.begin(), cnt.end() - 1);
        for (int r = 0; r < nr; r++) {
            if (in_set[(size_t)r]) hp3[(size_t)r] = T.reads.h3[r];
            for (uint64_t c = T.call_off[r]; c < T.call_off[r + 1]; c++)
                if (T.calls[c].quality & 1) {
                    const uint32_t at = fill[(size_t)slot_of_var[(size_t)T.calls[c].var]]++;
                    pr_read[at] = r; pr_hp[at] = T.calls[c].allele;
                }
        }
    }
    auto sum9 = [](const int32_t *p) { int s = 0; for (int i = 0; i < 9; i++) s += p[i]; return s; };
    std::vector<uint8_t> touched((size_t)nt, 0);
    for (int k = 0; k < nt; k++)
        touched[(size_t)k] = T.pos_base[(size_t)k * LPS_PB_FIELDS + LPS_PB_DEPTH] > 0 || sum9(T.read_hp_count + (size_t)k * 9) > 0 ||
                             sum9(T.somatic_read_hp_count + (size_t)k * 9) > 0;

    // ---- getDenseTumorSnpInterval (:1243-1351) ----
    std::vector<float> mean_alt((size_t)nt, 0.f), z_score((size_t)nt, 0.f);
    std::vector<int32_t> interval_cnt((size_t)nt, 0), min_dist((size_t)nt, 0), same_cnt((size_t)nt, 0);
    std::vector<uint8_t> in_dense((size_t)nt, 0), filt((size_t)nt * 6, 0), filter_out((size_t)nt, 0), high_con((size_t)nt, 0);
    std::vector<int8_t> derive((size_t)nt, 0);
    for (int k = 0; k < nt; k++) {
        if (cnt[(size_t)k] == cnt[(size_t)k + 1]) continue;
        float readCount = 0.0, altMean = 0.0;
        for (uint32_t i = cnt[(size_t)k]; i < cnt[(size_t)k + 1]; i++) {
            if (pr_hp[i] != SNP_H3) continue;
            readCount++;
            altMean += hp3[(size_t)pr_read[i]];
        }
        if (altMean != 0) altMean /= readCount;
        mean_alt[(size_t)k] = altMean;
    }
    {
        struct Interval { std::map<int, double> altMean, z; std::map<int, int> minDistance; int snpCount = 0; };
        std::vector<int> ts;                                           // touched slots, ascending position
        for (int k = 0; k < nt; k++) if (touched[(size_t)k]) ts.push_back(k);
        std::vector<Interval> done;
        Interval cur;
        bool rec = false;
        int startPos = 0;
        const int dense_distance = 5000;                               // INTERVAL_SNP_MAX_DISTANCE (SomaticVarCaller.h:462)
        auto close = [&]() {
            const double size = (double)cur.altMean.size();
            double sum = 0.0;
            for (auto &kv : cur.altMean) sum += kv.second;
            const double mean = size == 0 ? 0.0 : sum / size;          // statisticsUtils::calculateMean (:43-52)
            double var = 0.0;
            for (auto &kv : cur.altMean) var += (kv.second - mean) * (kv.second - mean);
            const double sd = std::sqrt(var / (double)cur.altMean.size());
            for (auto &kv : cur.altMean) cur.z[kv.first] = sd == 0 ? 0.0 : (kv.second - mean) / sd;
            done.push_back(cur);
        };
        for (size_t i = 0; i + 1 < ts.size(); i++) {
            const int curSlot = ts[i], nextSlot = ts[i + 1], curPos = in->pos[curSlot], nextPos = in->pos[nextSlot];
            const int d = nextPos - curPos;
            if (d <= dense_distance) {
                if (!rec) {
                    rec = true; startPos = curPos;
                    cur.altMean[curSlot] = mean_alt[(size_t)curSlot]; cur.minDistance[curSlot] = d; cur.snpCount++;
                }
                if (d < cur.minDistance[curSlot]) cur.minDistance[curSlot] = d;
                cur.altMean[nextSlot] = mean_alt[(size_t)nextSlot]; cur.minDistance[nextSlot] = d;
                cur.snpCount++;
            } else if (rec) {
                close();
                rec = false; startPos = 0; cur = Interval();
            }
        }
        if (rec && !ts.empty() && in->pos[ts.back()] - startPos <= dense_distance) close();     // a longer trailing run is dropped (:1327-1332)
        for (const Interval &iv : done) {
            if (iv.snpCount <= 1) continue;
            for (auto &kv : iv.z) {
                in_dense[(size_t)kv.first] = 1;
                z_score[(size_t)kv.first] = (float)std::abs(kv.second);
                interval_cnt[(size_t)kv.first] = iv.snpCount;
            }
            for (auto &kv : iv.minDistance) min_dist[(size_t)kv.first] = kv.second;
        }
    }

    // ---- somaticFeatureFilter (:1062-1230) ----
    for (int k = 0; k < nt; k++) {
        if (!touched[(size_t)k] || !in->callable[k]) continue;
        const float norVAF = N.ratios_f[(size_t)k * LPS_RF_FIELDS + LPS_RF_VAF];
        const float norDepth = N.pos_base[(size_t)k * LPS_PB_FIELDS + LPS_PB_DEPTH];
        const float tumVAF = T.ratios_f[(size_t)k * LPS_RF_FIELDS + LPS_RF_VAF], mixed = T.ratios_f[(size_t)k * LPS_RF_FIELDS + LPS_RF_MIXED_RATIO];
        const int caseReads = T.case_read_count[k];
        uint8_t *f = &filt[(size_t)k * 6];
        f[0] = !(norVAF <= fp.norVAF_maxThr && norDepth > fp.norDepth_minThr);                                     // TINC
        f[1] = mixed >= fp.MessyReadRatioThreshold;
        f[2] = caseReads <= fp.ReadCount_minThr;
        const int h11 = T.somatic_read_hp_count[(size_t)k * 9 + HP_H1_1], h21 = T.somatic_read_hp_count[(size_t)k * 9 + HP_H2_1];
        f[3] = caseReads <= fp.HapConsistency_ReadCount_maxThr && tumVAF <= fp.HapConsistency_VAF_maxThr &&
               h11 > fp.HapConsistency_somaticRead_minThr && h21 > fp.HapConsistency_somaticRead_minThr;
        f[4] = caseReads <= fp.IntervalSnpCount_ReadCount_maxThr && tumVAF <= fp.IntervalSnpCount_VAF_maxThr &&
               interval_cnt[(size_t)k] > fp.IntervalSnpCount_minThr && z_score[(size_t)k] <= fp.zScore_maxThr && z_score[(size_t)k] >= 0.0;
        // DenseAlt (:1160-1203): offsets at which the reads carrying ALT differ from the reference far more often than the others
        const int32_t *refh = T.window_hist + ((size_t)k * 2 + 0) * LPS_WINDOW_BINS, *alth = refh + LPS_WINDOW_BINS;
        const int altCount = T.pos_base[(size_t)k * LPS_PB_FIELDS + LPS_PB_ALT];
        int same = 0;
        for (int b = 0; b < LPS_WINDOW_BINS; b++) {
            const int aa = alth[b], ra = refh[b];
            if (aa == 0) continue;
            const double c1 = (double)aa / altCount, c2 = (double)aa / (ra + aa);
            if (c1 >= fp.DenseAlt_condition1_thr && c2 >= fp.DenseAlt_condition2_thr && ++same == fp.DenseAlt_sameCount_minThr) break;
        }
        same_cnt[(size_t)k] = same;
        f[5] = same >= fp.DenseAlt_sameCount_minThr;
        filter_out[(size_t)k] = f[0] || f[1] || f[2] || f[3] || f[4] || f[5];
        if (in->enable_filter && filter_out[(size_t)k]) continue;
        high_con[(size_t)k] = 1;
    }

    // ---- calibrateReadHP (:1366-1403): reads lose the H3 votes of positions that did not pass ----
    for (int k = 0; k < nt; k++) {
        if (!touched[(size_t)k] || high_con[(size_t)k]) continue;
        if (cnt[(size_t)k] == cnt[(size_t)k + 1]) return LPS_E_DATA;     // reference: "[ERROR](calibrate read HP) => can't find pos", exit(1)
        for (uint32_t i = cnt[(size_t)k]; i < cnt[(size_t)k + 1]; i++)
            if (pr_hp[i] == SNP_H3 && --hp3[(size_t)pr_read[i]] < 0) return LPS_E_DATA;
    }
    // ---- calculateReadSetHP (:1418-1439) ----
    std::vector<int8_t> read_hp((size_t)nr, -1);
    for (int r = 0; r < nr; r++)
        if (in_set[(size_t)r])
            read_hp[(size_t)r] = (int8_t)judge_somatic_read_hap(T.reads.h1[r], T.reads.h2[r], hp3[(size_t)r], 0, T.reads.n_ps[r], in->percentage_threshold);
    // ---- statisticSomaticPosReadHP (:1441-1518): which germline haplotype the somatic reads of a position derive from ----
    for (int k = 0; k < nt; k++) {
        if (!touched[(size_t)k] || !high_con[(size_t)k]) continue;
        if (cnt[(size_t)k] == cnt[(size_t)k + 1]) return LPS_E_DATA;     // reference: "[ERROR](statistic all read HP) => can't find pos", exit(1)
        int d11 = 0, d21 = 0;
        for (uint32_t i = cnt[(size_t)k]; i < cnt[(size_t)k + 1]; i++) {
            if (pr_hp[i] != SNP_H3) continue;
            const int hp = read_hp[(size_t)pr_read[i]];
            d11 += hp == HP_H1_1; d21 += hp == HP_H2_1;
        }
        const int tot = d11 + d21;
        float r11 = 0.0, r21 = 0.0;
        if (tot > 0) {
            if (d11 > 0) r11 = (float)d11 / (float)tot;
            if (d21 > 0) r21 = (float)d21 / (float)tot;
        }
        derive[(size_t)k] = r11 >= 1.0 ? SNP_H1 : r21 >= 1.0 ? SNP_H2 : SNP_NONE;
    }

    auto put = [&](auto *dst, const auto &src) { if (dst) std::copy(src.begin(), src.end(), dst); };
    put(out->touched, touched); put(out->mean_alt_per_var_read, mean_alt); put(out->z_score, z_score); put(out->interval_snp_count, interval_cnt);
    put(out->min_distance, min_dist); put(out->dense_alt_same_count, same_cnt); put(out->in_dense_interval, in_dense); put(out->filtered_by, filt);
    put(out->is_filter_out, filter_out); put(out->is_somatic, high_con); put(out->derive_hp, derive); put(out->read_hp, read_hp);
    if (out->read_h3) for (int r = 0; r < nr; r++) out->read_h3[r] = in_set[(size_t)r] ? hp3[(size_t)r] : -1;
    out->n_somatic = 0;
    for (int k = 0; k < nt; k++) out->n_somatic += high_con[(size_t)k];
    return LPS_OK;
}
