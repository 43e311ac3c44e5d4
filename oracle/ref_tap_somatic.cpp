/*
 * ref_tap_somatic.cpp — TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference's somatic objects on in-memory bam1_t
 * records built from the SoA batch of include/lps.h:
 *   mode 0  ExtractNorDataChrProcessor::processRead   (src/somatic_haplotag/SomaticVarCaller.cpp:123-173)
 *   mode 1  ExtractTumDataChrProcessor::processRead   (:334-459)
 *   mode 2  SomaticHaplotagChrProcessor::judgeHaplotype (src/somatic_haplotag/SomaticHaplotagProcess.cpp:310-459)
 * and flattens the reference's own maps (chrPosNorBase, chrPosSomaticInfo, readHpResultSet, tumorPosReadCorrBaseHP,
 * chrReadHpResult) into the slot-indexed arrays of oracle.h's orc_somatic_out.  The dispatch of
 * ChromosomeProcessor::processSingleChrom (HaplotagParsingBam.cpp:457-486) can only run from a BAM file, so the driver
 * applies the same seven-way test before calling the reference's processRead path.  Values the reference keeps in
 * locals only (per-read counts of reads without a tumor position) are reported as -1 = "not observable".
 */
#include <bits/stdc++.h>
#include <htslib/sam.h>
#include <htslib/faidx.h>
#include <htslib/khash.h>
#include <htslib/kbitset.h>
#include <htslib/thread_pool.h>
#include <htslib/vcf.h>
#include <htslib/vcfutils.h>
#include <zlib.h>
#include <omp.h>

#define private public
#define protected public
#include "somatic_haplotag/SomaticHaplotagProcess.h"
#include "somatic_haplotag/SomaticVarCaller.h"
#undef private
#undef protected

#include "../include/lps.h"
#include "oracle.h"
#include "ref_tap.h"

namespace {
struct TapState {
    std::map<std::string, std::map<int, PosBase>> nor;
    std::map<std::string, std::map<int, SomaticData>> tum;
    std::map<std::string, std::map<std::string, ReadVarHpCount>> readSet;
    std::map<std::string, std::map<int, std::map<std::string, int>>> posRead;
    std::map<std::string, std::vector<std::string>> readKey;    // per alignment of the tumor batch: its key in readSet, "" if none
};
template <typename T> T *zalloc(size_t n) { return (T *)calloc(n + 1, sizeof(T)); }

VarData make_var(const char *s, int gt, int hp1_is_alt, int ps) {
    VarData vd;
    vd.allele.Ref = s;
    vd.allele.Alt = s + vd.allele.Ref.size() + 1;
    vd.GT = (GenomeType)gt;
    vd.setVariantType();
    if (gt == GenomeType::PHASED_HETERO) {
        vd.PhasedSet = ps;
        if (hp1_is_alt) { vd.HP1 = vd.allele.Alt; vd.HP2 = vd.allele.Ref; }
        else { vd.HP1 = vd.allele.Ref; vd.HP2 = vd.allele.Alt; }
    }
    return vd;
}

void fill_pos_base(int32_t *pb, const PosBase &b) {
    pb[LPS_PB_ALT] = b.altCount; pb[LPS_PB_A] = b.A_count; pb[LPS_PB_C] = b.C_count; pb[LPS_PB_G] = b.G_count; pb[LPS_PB_T] = b.T_count;
    pb[LPS_PB_UNKNOWN] = b.unknow; pb[LPS_PB_DEPTH] = b.depth; pb[LPS_PB_DEL] = b.delCount;
    pb[LPS_PB_MPQ_ALT] = b.MPQ_altCount; pb[LPS_PB_MPQ_A] = b.MPQ_A_count; pb[LPS_PB_MPQ_C] = b.MPQ_C_count;
    pb[LPS_PB_MPQ_G] = b.MPQ_G_count; pb[LPS_PB_MPQ_T] = b.MPQ_T_count; pb[LPS_PB_MPQ_UNKNOWN] = b.MPQ_unknow;
    pb[LPS_PB_MPQ_DEPTH] = b.filteredMpqDepth;
}
void fill_hp9(int32_t *dst, const std::map<int, int> &m) {
    for (auto &kv : m) if (kv.first >= 0 && kv.first < 9) dst[kv.first] = kv.second;
}
}  // namespace

extern "C" int ref_tap_somatic(int mode, const tap_som_in *in, orc_somatic_out *out) {
    memset(out, 0, sizeof(*out));
    const lps_read_batch &b = in->batch;
    std::string chr = in->chr;
    std::string ref_string(in->ref, (size_t)in->ref_len);
    if (!in->p.have_reference) ref_string = "";

    // the union map exactly as VcfParser stores NORMAL / TUMOR records (HaplotagVcfParser.cpp:336-402)
    std::map<int, MultiGenomeVar> currentVariants;
    std::map<int, int> slot_of_pos;
    std::vector<int> tum_var;
    for (int i = 0; i < in->n_var; i++) {
        MultiGenomeVar &mv = currentVariants[in->var_pos[i]];
        if (!in->nor_present || in->nor_present[i])
            mv.Variant[NORMAL] = make_var(in->var_str + in->var_str_off[i], in->nor_gt ? in->nor_gt[i] : 1, in->var_hp1_is_alt[i], in->var_ps[i]);
        if (in->tum_present[i]) {
            mv.Variant[TUMOR] = make_var(in->tum_str + in->tum_str_off[i], in->tum_gt[i], in->tum_hp1_is_alt[i], in->tum_ps[i]);
            slot_of_pos[in->var_pos[i]] = (int)tum_var.size();
            tum_var.push_back(i);
        }
        mv.isSomaticVariant = in->is_somatic[i] != 0;
        mv.somaticReadDeriveByHP = in->derive_hp[i];
    }
    const int n = b.n_reads, nt = (int)tum_var.size();
    out->n_reads = n; out->n_tum = nt;
    out->tum_var = zalloc<int32_t>(nt);
    for (int i = 0; i < nt; i++) out->tum_var[i] = tum_var[i];
    out->category = zalloc<uint8_t>(n); out->read_hp = zalloc<int8_t>(n); out->hp_before = zalloc<int8_t>(n);
    out->ps = zalloc<int32_t>(n); out->pq = zalloc<int32_t>(n); out->h1 = zalloc<int32_t>(n); out->h2 = zalloc<int32_t>(n);
    out->h3 = zalloc<int32_t>(n); out->n_ps = zalloc<uint8_t>(n); out->end_pos = zalloc<int32_t>(n); out->read_len = zalloc<int32_t>(n);
    out->derive_similarity = zalloc<float>(n);
    out->pos_base = zalloc<int32_t>((size_t)nt * LPS_PB_FIELDS); out->read_hp_count = zalloc<int32_t>((size_t)nt * 9);
    out->somatic_read_hp_count = zalloc<int32_t>((size_t)nt * 9); out->case_count = zalloc<int32_t>((size_t)nt * LPS_CASE_FIELDS);
    out->allele_count = zalloc<int32_t>((size_t)nt * 2); out->window_hist = zalloc<int32_t>((size_t)nt * 2 * LPS_WINDOW_BINS);
    out->hp_before_count = zalloc<int32_t>((size_t)nt * 9); out->hp_after_count = zalloc<int32_t>((size_t)nt * 9);
    out->h3_before_count = zalloc<int32_t>((size_t)nt * 9); out->h3_after_count = zalloc<int32_t>((size_t)nt * 9);
    out->cover_start = zalloc<int32_t>(nt); out->cover_end = zalloc<int32_t>(nt);
    out->ratios_f = zalloc<float>((size_t)nt * LPS_RF_FIELDS); out->ratios_d = zalloc<double>((size_t)nt * LPS_RD_FIELDS);
    out->case_read_count = zalloc<int32_t>(nt);
    for (int i = 0; i < nt; i++) { out->cover_start[i] = INT_MAX; out->cover_end[i] = INT_MIN; }
    out->call_off = zalloc<uint64_t>((size_t)n + 1);
    std::vector<lps_call> calls;

    ParsingBamConfig cfg;
    cfg.numThreads = 1; cfg.qualityThreshold = in->p.mapping_quality; cfg.percentageThreshold = in->p.percentage_threshold;
    cfg.resultPrefix = "/tmp/ref_tap_somatic"; cfg.region = ""; cfg.command = ""; cfg.version = ""; cfg.outputFormat = "bam";
    cfg.tagSupplementary = in->p.tag_supplementary != 0; cfg.writeReadLog = false;
    std::map<Genome, VCF_Info> vcfSet;
    int chrLength = (int)in->ref_len;
    ChrProcContext pctx(chr, chrLength, cfg, mode == 0 ? NORMAL : TUMOR, vcfSet);

    std::map<std::string, std::map<int, PosBase>> chrPosNorBase;
    std::map<std::string, std::map<int, SomaticData>> chrPosSomaticInfo;
    std::map<std::string, std::map<std::string, ReadVarHpCount>> chrReadHpResultSet;
    std::map<std::string, std::map<int, std::map<std::string, int>>> chrTumorPosReadCorrBaseHP;
    chrPosNorBase[chr]; chrPosSomaticInfo[chr]; chrReadHpResultSet[chr]; chrTumorPosReadCorrBaseHP[chr];
    ReadStatistics readStats;
    SomaticReadBenchmark bench("", "", cfg.qualityThreshold);
    ReadHpDistriLog before, after;
    ExtractNorDataChrProcessor *norProc = nullptr;
    ExtractTumDataChrProcessor *tumProc = nullptr;
    SomaticHaplotagChrProcessor *tagProc = nullptr;
    if (mode == 0) norProc = new ExtractNorDataChrProcessor(chrPosNorBase, chr);
    else if (mode == 1) tumProc = new ExtractTumDataChrProcessor(chrPosSomaticInfo, chrReadHpResultSet, chrTumorPosReadCorrBaseHP, chr);
    else tagProc = new SomaticHaplotagChrProcessor(false, in->p.mapq_filter != 0, readStats, nullptr, bench, before, after, chr);
    const bool mapq_filter = in->p.mapq_filter != 0;

    bam_hdr_t hdr;
    memset(&hdr, 0, sizeof(hdr));
    char *tname = strdup(in->chr);
    hdr.n_targets = 1; hdr.target_name = &tname;
    std::map<int, MultiGenomeVar>::iterator firstVariantIter = currentVariants.begin(), firstVariantIter2 = currentVariants.begin();
    std::map<int, MultiGenomeVar>::reverse_iterator last = currentVariants.rbegin();
    std::vector<uint8_t> data;
    bam1_t aln;
    memset(&aln, 0, sizeof(aln));
    auto &readSet = chrReadHpResultSet[chr];
    auto &posRead = chrTumorPosReadCorrBaseHP[chr];
    std::vector<std::string> readKeys((size_t)n);

    for (int32_t r = 0; r < n; r++) {
        out->call_off[r] = calls.size();
        out->h1[r] = out->h2[r] = out->h3[r] = -1; out->end_pos[r] = out->read_len[r] = -1; out->n_ps[r] = 255;
        out->hp_before[r] = -1; out->derive_similarity[r] = -1.f;
        int flag = b.flag[r];
        int category;
        if (b.mapq[r] < cfg.qualityThreshold && mapq_filter) category = LPS_TAG_LOW_MAPQ;
        else if ((flag & 0x4) != 0) category = LPS_TAG_UNMAPPED;
        else if ((flag & 0x100) != 0) category = LPS_TAG_SECONDARY;
        else if ((flag & 0x800) != 0 && cfg.tagSupplementary == false) category = LPS_TAG_SUPPLEMENTARY;
        else if (last == currentVariants.rend()) category = LPS_TAG_EMPTY_VARIANTS;
        else if (int(b.ref_start[r]) <= (*last).first) category = LPS_TAG_PROCESSED;
        else category = LPS_TAG_OTHER;
        out->category[r] = (uint8_t)category;
        if (tagProc) {
            switch (category) {
                case LPS_TAG_LOW_MAPQ: tagProc->processLowMappingQuality(); break;
                case LPS_TAG_UNMAPPED: tagProc->processUnmappedRead(); break;
                case LPS_TAG_SECONDARY: tagProc->processSecondaryAlignment(); break;
                case LPS_TAG_SUPPLEMENTARY: tagProc->processSupplementaryAlignment(); break;
                case LPS_TAG_EMPTY_VARIANTS: tagProc->processEmptyVariants(); break;
                case LPS_TAG_OTHER: tagProc->processOtherCase(); break;
                default: break;
            }
        }
        if (category != LPS_TAG_PROCESSED) continue;
        const char *name = in->names + (size_t)r * in->name_stride;
        size_t ln = strlen(name) + 1, lnp = (ln + 3) & ~(size_t)3;
        size_t nbytes = lnp + 4 * (size_t)b.n_cigar[r] + ((size_t)b.l_qseq[r] + 1) / 2 + (size_t)b.l_qseq[r];
        data.assign(nbytes + 64, 0);
        memcpy(data.data(), name, ln);
        memcpy(data.data() + lnp, b.cigar + b.cigar_off[r], 4 * (size_t)b.n_cigar[r]);
        memcpy(data.data() + lnp + 4 * (size_t)b.n_cigar[r], b.seq4 + b.seq_off[r], ((size_t)b.l_qseq[r] + 1) / 2);
        memcpy(data.data() + lnp + 4 * (size_t)b.n_cigar[r] + ((size_t)b.l_qseq[r] + 1) / 2, b.qual + b.qual_off[r], (size_t)b.l_qseq[r]);
        aln.data = data.data(); aln.l_data = (int)nbytes; aln.m_data = (uint32_t)data.size();
        aln.core.pos = b.ref_start[r]; aln.core.tid = 0; aln.core.qual = b.mapq[r]; aln.core.flag = b.flag[r];
        aln.core.l_qname = (uint16_t)lnp; aln.core.l_extranul = (uint8_t)(lnp - ln);
        aln.core.n_cigar = b.n_cigar[r]; aln.core.l_qseq = b.l_qseq[r];

        if (mode == 0) {
            norProc->processRead(aln, hdr, ref_string, currentVariants, firstVariantIter, pctx);
            // observe the read's germline haplotype with the reference's own parser + strategy on scratch maps
            std::map<int, int> hpCount, variantsHP, norCountPS;
            hpCount[SnpHP::GERMLINE_H1] = 0; hpCount[SnpHP::GERMLINE_H2] = 0;
            std::map<int, PosBase> scratchBase;
            std::vector<int> scratchPos;
            int ref_pos = aln.core.pos, query_pos = 0;
            CigarParserContext cctx(aln, hdr, chr, cfg, firstVariantIter2, currentVariants, ref_string);
            CigarParser *parser = new ExtractNorDataCigarParser(cctx, scratchBase, scratchPos, ref_pos, query_pos, cfg.qualityThreshold);
            parser->parsingCigar(hpCount, variantsHP, norCountPS);
            delete parser;
            GermlineHaplotagStrategy judger;
            double mn = 0, mx = 0; int pq = 0, ps = 0;
            int hp = judger.judgeReadHap(hpCount, mn, mx, cfg.percentageThreshold, pq, ps, norCountPS, nullptr, nullptr);
            out->read_hp[r] = (int8_t)hp; out->pq[r] = pq; out->ps[r] = hp ? ps : 0;
            out->h1[r] = hpCount[1]; out->h2[r] = hpCount[2]; out->h3[r] = 0;
            out->n_ps[r] = (uint8_t)std::min<size_t>(norCountPS.size(), 2);
            out->end_pos[r] = ref_pos; out->read_len[r] = query_pos;
        } else if (mode == 1) {
            std::string key = name;
            auto it = readSet.find(key);
            if (it != readSet.end()) key = key + "-" + std::to_string(it->second.readIDcount + 1);
            const size_t before_n = readSet.size();
            tumProc->processRead(aln, hdr, ref_string, currentVariants, firstVariantIter, pctx);
            if (readSet.size() != before_n) {
                auto rec = readSet.find(key);
                if (rec == readSet.end()) { fprintf(stderr, "ref_tap_somatic: record %s not found\n", key.c_str()); return -1; }
                const ReadVarHpCount &rv = rec->second;
                readKeys[(size_t)r] = key;
                out->read_hp[r] = (int8_t)rv.hpResult; out->h1[r] = rv.HP1; out->h2[r] = rv.HP2; out->h3[r] = rv.HP3;
                out->n_ps[r] = (uint8_t)std::min<size_t>(rv.norCountPS.size(), 2);
                out->end_pos[r] = rv.endPos; out->read_len[r] = rv.readLength;
                out->ps[r] = rv.hpResult ? (rv.norCountPS.empty() ? -1 : rv.norCountPS.begin()->first) : 0;
                // merged view of posHpPairs (variantsHP of the read, 1-based) and tumorPosReadCorrBaseHP (tumorSnpPosVec)
                std::map<int, std::pair<int, int>> merged;   // pos0 -> (vhp, flags)
                for (auto &ph : rv.posHpPairs) merged[ph.first - 1].first = ph.second;
                out->pq[r] = rv.posHpPairs.empty() ? 0 : 1;   // 1: posHpPairs was recorded for this read
                for (int k = 0; k < nt; k++) {
                    int pos = in->var_pos[tum_var[k]];
                    if (pos < b.ref_start[r] || pos > rv.endPos) continue;
                    auto pr = posRead.find(pos);
                    if (pr == posRead.end()) continue;
                    auto rr = pr->second.find(key);
                    if (rr == pr->second.end()) continue;
                    merged[pos].second |= 1;
                    if (merged[pos].first == 0) merged[pos].first = rr->second;
                    else if (rr->second != 0 && merged[pos].first != rr->second) { fprintf(stderr, "ref_tap_somatic: baseHP mismatch\n"); return -1; }
                }
                for (auto &kv : merged) {
                    auto vit = std::lower_bound(in->var_pos, in->var_pos + in->n_var, kv.first);
                    lps_call c = {(int32_t)(vit - in->var_pos), (int16_t)kv.second.second, (int8_t)kv.second.first, 0};
                    calls.push_back(c);
                }
            } else {
                out->read_hp[r] = -1;
            }
        } else {
            int pq = 0, ps = 0;
            if ((aln.core.flag & 0x800) != 0) tagProc->localReadStats.totalSupplementary++;
            int hp = tagProc->judgeHaplotype(hdr, aln, chr, cfg.percentageThreshold, nullptr, pq, ps, TUMOR, ref_string, cfg, firstVariantIter,
                                             currentVariants, vcfSet);
            if (hp != ReadHP::unTag) { tagProc->localReadStats.totalHpCount[hp]++; tagProc->localReadStats.totalTagCount++; }
            else { tagProc->localReadStats.totalHpCount[ReadHP::unTag]++; tagProc->localReadStats.totalUnTagCount++; }
            tagProc->localReadStats.totalAlignment++;
            out->read_hp[r] = (int8_t)hp; out->pq[r] = pq; out->ps[r] = hp ? ps : 0;
        }
    }
    out->call_off[n] = calls.size();
    out->n_calls = calls.size();
    out->calls = zalloc<lps_call>(calls.size());
    if (!calls.empty()) memcpy(out->calls, calls.data(), sizeof(lps_call) * calls.size());

    // the reference's own postProcess (ratios of every touched position)
    if (mode == 0) norProc->postProcess(chr, currentVariants);
    else if (mode == 1) tumProc->postProcess(chr, currentVariants);
    auto fill_ratios = [&](size_t sl, const PosBase &pb) {
        float *f = out->ratios_f + sl * LPS_RF_FIELDS;
        double *d = out->ratios_d + sl * LPS_RD_FIELDS;
        f[LPS_RF_VAF] = pb.VAF; f[LPS_RF_NONDEL_VAF] = pb.nonDelVAF; f[LPS_RF_MPQ_VAF] = pb.filteredMpqVAF;
        f[LPS_RF_LOW_MPQ_RATIO] = pb.lowMpqReadRatio; f[LPS_RF_DEL_RATIO] = pb.delRatio;
        d[LPS_RD_GERMLINE_IMBALANCE] = pb.germlineHaplotypeImbalanceRatio; d[LPS_RD_PCT_GERMLINE_HP] = pb.percentageOfGermlineHp;
    };
    if (mode == 0) {
        for (auto &kv : chrPosNorBase[chr]) {
            auto s = slot_of_pos.find(kv.first);
            if (s == slot_of_pos.end()) { fprintf(stderr, "ref_tap_somatic: PosBase at a non-tumor position\n"); return -1; }
            fill_pos_base(out->pos_base + (size_t)s->second * LPS_PB_FIELDS, kv.second);
            fill_hp9(out->read_hp_count + (size_t)s->second * 9, kv.second.ReadHpCount);
            fill_ratios((size_t)s->second, kv.second);
        }
    } else if (mode == 1) {
        for (auto &kv : chrPosSomaticInfo[chr]) {
            auto s = slot_of_pos.find(kv.first);
            if (s == slot_of_pos.end()) { fprintf(stderr, "ref_tap_somatic: SomaticData at a non-tumor position\n"); return -1; }
            const size_t sl = (size_t)s->second;
            const SomaticData &sd = kv.second;
            fill_pos_base(out->pos_base + sl * LPS_PB_FIELDS, sd.base);
            fill_hp9(out->read_hp_count + sl * 9, sd.base.ReadHpCount);
            fill_hp9(out->somatic_read_hp_count + sl * 9, sd.somaticReadHpCount);
            int32_t *cc = out->case_count + sl * LPS_CASE_FIELDS;
            cc[LPS_CASE_CLEAN_HP3] = sd.totalCleanHP3Read; cc[LPS_CASE_PURE_H1_1] = sd.pure_H1_1_read; cc[LPS_CASE_PURE_H2_1] = sd.pure_H2_1_read;
            cc[LPS_CASE_PURE_H3] = sd.pure_H3_read; cc[LPS_CASE_MIXED] = sd.Mixed_HP_read; cc[LPS_CASE_UNTAG] = sd.unTag;
            out->allele_count[sl * 2] = sd.alleleCount[0]; out->allele_count[sl * 2 + 1] = sd.alleleCount[1];
            fill_ratios(sl, sd.base);
            out->ratios_f[sl * LPS_RF_FIELDS + LPS_RF_MIXED_RATIO] = sd.Mixed_HP_readRatio;
            out->ratios_f[sl * LPS_RF_FIELDS + LPS_RF_PURE_H1_1_RATIO] = sd.pure_H1_1_readRatio;
            out->ratios_f[sl * LPS_RF_FIELDS + LPS_RF_PURE_H2_1_RATIO] = sd.pure_H2_1_readRatio;
            out->ratios_f[sl * LPS_RF_FIELDS + LPS_RF_PURE_H3_RATIO] = sd.pure_H3_readRatio;
            out->ratios_d[sl * LPS_RD_FIELDS + LPS_RD_ALLELIC_IMBALANCE] = sd.allelicImbalanceRatio;
            out->ratios_d[sl * LPS_RD_FIELDS + LPS_RD_SOMATIC_IMBALANCE] = sd.somaticHaplotypeImbalanceRatio;
            out->case_read_count[sl] = sd.CaseReadCount;
            for (int a = 0; a < 2; a++)
                for (auto &ob : sd.PosSomaticOffsetBase[a]) {
                    if (ob.first < -LPS_WINDOW || ob.first > LPS_WINDOW) { fprintf(stderr, "ref_tap_somatic: offset out of range\n"); return -1; }
                    out->window_hist[(sl * 2 + a) * LPS_WINDOW_BINS + ob.first + LPS_WINDOW]++;
                }
        }
    } else {
        chrReadHpResult *bf = before.getChrHpResultsPtr(chr), *af = after.getChrHpResultsPtr(chr);
        for (auto &kv : bf->posReadHpResult) {
            auto s = slot_of_pos.find(kv.first);
            if (s == slot_of_pos.end()) { fprintf(stderr, "ref_tap_somatic: somatic position without a tumor record\n"); return -1; }
            fill_hp9(out->hp_before_count + (size_t)s->second * 9, kv.second.readHpCounter);
            fill_hp9(out->h3_before_count + (size_t)s->second * 9, kv.second.somaticBaseReadHpCounter);
        }
        for (auto &kv : af->posReadHpResult) {
            auto s = slot_of_pos.find(kv.first);
            if (s == slot_of_pos.end()) return -1;
            fill_hp9(out->hp_after_count + (size_t)s->second * 9, kv.second.readHpCounter);
            fill_hp9(out->h3_after_count + (size_t)s->second * 9, kv.second.somaticBaseReadHpCounter);
            out->cover_start[s->second] = kv.second.coverRegionStartPos; out->cover_end[s->second] = kv.second.coverRegionEndPos;
        }
        const ReadStatistics &s = tagProc->localReadStats;
        int64_t st[TAP_SOM_STATS] = {s.totalAlignment, s.totalSupplementary, s.totalSecondary, s.totalUnmapped, s.totalTagCount, s.totalUnTagCount,
                                     s.totalLowerQuality, s.totalOtherCase, s.totalEmptyVariant, s.totalHighSimilarity, s.totalCrossTwoBlock,
                                     s.totalWithOutVaraint, s.totalreadOnlyH3Snp};
        for (int k = 0; k < 9; k++) { auto it = s.totalHpCount.find(k); st[13 + k] = it == s.totalHpCount.end() ? 0 : it->second; }
        memcpy(in->stats_out, st, sizeof(st));
    }
    if (in->keep) {
        TapState *st = (TapState *)in->keep;
        if (mode == 0) st->nor[chr] = chrPosNorBase[chr];
        else if (mode == 1) {
            st->tum[chr] = chrPosSomaticInfo[chr];
            st->readSet[chr] = chrReadHpResultSet[chr];
            st->posRead[chr] = chrTumorPosReadCorrBaseHP[chr];
            st->readKey[chr] = readKeys;
        }
    }
    delete norProc; delete tumProc; delete tagProc;
    free(tname);
    return 0;
}

extern "C" void *ref_tap_state_new(void) { return new TapState(); }
extern "C" void ref_tap_state_free(void *state) { delete (TapState *)state; }

// The reference's own estimator (TumorPurityEstimator.cpp:31-84) on the maps of the two passes; the intermediate values come from
// running its private stages once more in the order estimateTumorPurity runs them.
extern "C" int ref_tap_purity(void *state, const char *chr, tap_purity_out *out) {
    TapState *st = (TapState *)state;
    memset(out, 0, sizeof(*out));
    std::vector<std::string> chrVec{std::string(chr)};
    // the estimator keeps a REFERENCE to the prefix (TumorPurityEstimator.h:292) and writes its report there: it must outlive both objects
    const std::string prefix = "/tmp/ref_tap_purity";
    {
        TumorPurityEstimator est(chrVec, st->nor, st->tum, false, prefix);
        out->purity = est.estimateTumorPurity();
    }
    try {
        TumorPurityEstimator e2(chrVec, st->nor, st->tum, false, prefix);
        std::vector<PurityData> v;
        e2.buildPurityFeatureValueVec(v);
        out->n_after_lcvf = (int32_t)v.size();
        int thr = e2.findBimodalValleyThreshold(v);
        e2.bimodalValleyFilter(v, thr);
        BoxPlotValue pv = e2.statisticPurityData(v);
        e2.removeOutliers(v, pv);
        pv = e2.statisticPurityData(v);
        out->threshold = thr; out->n_used = (int32_t)v.size();
        out->median = pv.median; out->q1 = pv.q1; out->q3 = pv.q3; out->iqr = pv.iqr; out->lower_whisker = pv.lowerWhisker; out->upper_whisker = pv.upperWhisker;
    } catch (const std::exception &) {
        out->n_used = -1;
    }
    return 0;
}

// The reference's own calling stage (SomaticVarCaller::variantCalling :816-866, minus extraction and logs) on the state of one
// contig, then getSomaticFlag (:2397-2412).  The private stages are called in the order variantCalling calls them.
extern "C" int ref_tap_somatic_call(void *state, const tap_som_in *in, double purity, int enable_filter, tap_call_out *out) {
    TapState *st = (TapState *)state;
    memset(out, 0, sizeof(*out));
    const std::string chr = in->chr;
    std::map<int, MultiGenomeVar> currentVariants;
    std::vector<int> tum_pos;
    for (int i = 0; i < in->n_var; i++) {
        MultiGenomeVar &mv = currentVariants[in->var_pos[i]];
        if (!in->nor_present || in->nor_present[i])
            mv.Variant[NORMAL] = make_var(in->var_str + in->var_str_off[i], in->nor_gt ? in->nor_gt[i] : 1, in->var_hp1_is_alt[i], in->var_ps[i]);
        if (in->tum_present[i]) {
            mv.Variant[TUMOR] = make_var(in->tum_str + in->tum_str_off[i], in->tum_gt[i], in->tum_hp1_is_alt[i], in->tum_ps[i]);
            tum_pos.push_back(in->var_pos[i]);
        }
    }
    const int nt = (int)tum_pos.size(), n = in->batch.n_reads;
    out->n_tum = nt; out->n_reads = n;
    out->touched = zalloc<uint8_t>(nt); out->mean_alt = zalloc<float>(nt); out->z_score = zalloc<float>(nt);
    out->interval_snp_count = zalloc<int32_t>(nt); out->min_distance = zalloc<int32_t>(nt); out->dense_alt_same = zalloc<int32_t>(nt);
    out->in_dense = zalloc<uint8_t>(nt); out->filtered_by = zalloc<uint8_t>((size_t)nt * 6); out->is_filter_out = zalloc<uint8_t>(nt);
    out->high_con = zalloc<uint8_t>(nt); out->derive_hp = zalloc<int32_t>(nt); out->is_somatic = zalloc<uint8_t>(nt);
    out->flag_derive_hp = zalloc<int32_t>(nt); out->read_hp = zalloc<int8_t>(n); out->read_h3 = zalloc<int32_t>(n);

    CallerConfig callerCfg(enable_filter != 0, false, purity, false);
    ParsingBamConfig cfg;
    cfg.numThreads = 1; cfg.qualityThreshold = in->p.mapping_quality; cfg.percentageThreshold = in->p.percentage_threshold;
    cfg.resultPrefix = "/tmp/ref_tap_somatic"; cfg.region = ""; cfg.command = ""; cfg.version = ""; cfg.outputFormat = "bam";
    cfg.tagSupplementary = in->p.tag_supplementary != 0; cfg.writeReadLog = false;
    std::vector<std::string> chrVec{chr};
    SomaticVarCaller caller(callerCfg, cfg, chrVec);
    (*caller.chrPosNorBase)[chr] = st->nor[chr];
    (*caller.chrPosSomaticInfo)[chr] = st->tum[chr];
    (*caller.chrReadHpResultSet)[chr] = st->readSet[chr];
    (*caller.chrTumorPosReadCorrBaseHP)[chr] = st->posRead[chr];

    double tumorPurity = purity;
    caller.setFilterParamsWithPurity(caller.somaticParams, tumorPurity);
    const SomaticVarFilterParams &sp = caller.somaticParams;
    out->tier = sp.zScore_maxThr == 5.233f ? 1 : sp.zScore_maxThr == 2.676f ? 2 : sp.zScore_maxThr == 5.683f ? 3 : sp.zScore_maxThr == 3.043f ? 4 : 5;
    std::map<int, SomaticData> &somaticPosInfo = (*caller.chrPosSomaticInfo)[chr];
    std::map<std::string, ReadVarHpCount> &readHpResultSet = (*caller.chrReadHpResultSet)[chr];
    std::map<int, std::map<std::string, int>> &tumorPosReadCorrBaseHP = (*caller.chrTumorPosReadCorrBaseHP)[chr];
    chrReadHpResult *distri = caller.callerReadHpDistri->getChrHpResultsPtr(chr);
    caller.getDenseTumorSnpInterval(somaticPosInfo, readHpResultSet, tumorPosReadCorrBaseHP, (*caller.denseTumorSnpInterval)[chr]);
    caller.somaticFeatureFilter(caller.somaticParams, currentVariants, chr, somaticPosInfo, tumorPurity);
    caller.calibrateReadHP(chr, somaticPosInfo, readHpResultSet, tumorPosReadCorrBaseHP);
    caller.calculateReadSetHP(chr, readHpResultSet, tumorPosReadCorrBaseHP, cfg.percentageThreshold);
    caller.statisticSomaticPosReadHP(chr, somaticPosInfo, tumorPosReadCorrBaseHP, readHpResultSet, *distri);
    std::map<std::string, std::map<int, MultiGenomeVar>> chrMultiVariants;
    chrMultiVariants[chr] = currentVariants;
    caller.getSomaticFlag(chrVec, chrMultiVariants);

    for (int k = 0; k < nt; k++) {
        auto it = somaticPosInfo.find(tum_pos[k]);
        const MultiGenomeVar &mv = chrMultiVariants[chr][tum_pos[k]];
        out->is_somatic[k] = mv.isSomaticVariant; out->flag_derive_hp[k] = mv.somaticReadDeriveByHP;
        if (it == somaticPosInfo.end()) continue;
        const SomaticData &sd = it->second;
        out->touched[k] = 1; out->mean_alt[k] = sd.meanAltCountPerVarRead; out->z_score[k] = sd.zScore;
        out->interval_snp_count[k] = sd.intervalSnpCount; out->min_distance[k] = sd.minDistance; out->dense_alt_same[k] = sd.denseAltSameCount;
        out->in_dense[k] = sd.inDenseTumorInterval;
        uint8_t *f = out->filtered_by + (size_t)k * 6;
        f[0] = sd.filteredByTINC; f[1] = sd.filteredByMessyRead; f[2] = sd.filteredByReadCount; f[3] = sd.filteredByHapConsistency;
        f[4] = sd.filteredByVariantCluster; f[5] = sd.filteredByDenseAlt;
        out->is_filter_out[k] = sd.isFilterOut; out->high_con[k] = sd.isHighConSomaticSNP; out->derive_hp[k] = sd.somaticReadDeriveByHP;
    }
    const std::vector<std::string> &keys = st->readKey[chr];
    for (int r = 0; r < n; r++) {
        out->read_hp[r] = -1; out->read_h3[r] = -1;
        if ((size_t)r >= keys.size() || keys[(size_t)r].empty()) continue;
        auto it = readHpResultSet.find(keys[(size_t)r]);
        if (it == readHpResultSet.end()) return -1;
        out->read_hp[r] = (int8_t)it->second.hpResult; out->read_h3[r] = it->second.HP3;
    }
    return 0;
}

extern "C" void ref_tap_somatic_call_free(tap_call_out *o) {
    free(o->touched); free(o->mean_alt); free(o->z_score); free(o->interval_snp_count); free(o->min_distance); free(o->dense_alt_same);
    free(o->in_dense); free(o->filtered_by); free(o->is_filter_out); free(o->high_con); free(o->derive_hp); free(o->is_somatic);
    free(o->flag_derive_hp); free(o->read_hp); free(o->read_h3);
    memset(o, 0, sizeof(*o));
}

extern "C" void ref_tap_somatic_free(orc_somatic_out *o) {
    free(o->tum_var); free(o->category); free(o->read_hp); free(o->hp_before); free(o->ps); free(o->pq); free(o->h1); free(o->h2);
    free(o->h3); free(o->n_ps); free(o->end_pos); free(o->read_len); free(o->derive_similarity); free(o->pos_base);
    free(o->read_hp_count); free(o->somatic_read_hp_count); free(o->case_count); free(o->allele_count); free(o->window_hist);
    free(o->hp_before_count); free(o->hp_after_count); free(o->h3_before_count); free(o->h3_after_count); free(o->cover_start);
    free(o->cover_end); free(o->call_off); free(o->calls); free(o->ratios_f); free(o->ratios_d); free(o->case_read_count);
    memset(o, 0, sizeof(*o));
}
