/*
 * ref_tap_tag.cpp — TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference's germline haplotag objects
 * (GermlineHaplotagChrProcessor::judgeHaplotype, GermlineHaplotagCigarParser / CigarParser::parsingCigar,
 * GermlineHaplotagStrategy) on in-memory bam1_t records built from the SoA batch of include/lps.h.
 * The dispatch of ChromosomeProcessor::processSingleChrom (HaplotagParsingBam.cpp:457-486) can only run from a
 * BAM file, so the driver applies the same seven-way test before calling the reference's processRead path.
 * For every processed read it records BOTH the reference's own judgeHaplotype result (hp, PS, PQ) and the maps
 * (hpCount, variantsHP, countPS) obtained by running the reference's parser + strategy objects a second time.
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
#include "haplotag/HaplotagProcess.h"
#undef private
#undef protected

#include "../include/lps.h"
#include "ref_tap.h"

namespace {
template <typename T> T *dupv(const std::vector<T> &v) {
    T *p = (T *)malloc(sizeof(T) * (v.size() + 1));
    if (!v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}

extern "C" int ref_tap_tag(const tap_tag_in *in, tap_tag_out *out) {
    memset(out, 0, sizeof(*out));
    const lps_read_batch &b = in->batch;
    std::string chr = in->chr;
    std::string ref_string(in->ref, (size_t)in->ref_len);
    if (!in->p.have_reference) ref_string = "";

    // variant map exactly as VcfParser stores a phased het NORMAL record (HaplotagVcfParser.cpp:336-402)
    std::map<int, MultiGenomeVar> currentVariants;
    for (int i = 0; i < in->n_var; i++) {
        VarData vd;
        const char *s = in->var_str + in->var_str_off[i];
        vd.allele.Ref = s;
        vd.allele.Alt = s + vd.allele.Ref.size() + 1;
        vd.GT = GenomeType::PHASED_HETERO;
        vd.setVariantType();
        vd.PhasedSet = in->var_ps[i];
        if (in->var_hp1_is_alt[i]) { vd.HP1 = vd.allele.Alt; vd.HP2 = vd.allele.Ref; }
        else { vd.HP1 = vd.allele.Ref; vd.HP2 = vd.allele.Alt; }
        currentVariants[in->var_pos[i]].Variant[NORMAL] = vd;
    }
    ParsingBamConfig cfg;
    cfg.numThreads = 1; cfg.qualityThreshold = in->p.mapping_quality; cfg.percentageThreshold = in->p.percentage_threshold;
    cfg.resultPrefix = "/tmp/ref_tap_tag"; cfg.region = ""; cfg.command = ""; cfg.version = ""; cfg.outputFormat = "bam";
    cfg.tagSupplementary = in->p.tag_supplementary != 0; cfg.writeReadLog = false;
    std::map<Genome, VCF_Info> vcfSet;
    ReadStatistics readStats;
    GermlineHaplotagChrProcessor proc(false, in->p.mapq_filter != 0, readStats, nullptr);
    int chrLength = (int)in->ref_len;
    ChrProcContext pctx(chr, chrLength, cfg, NORMAL, vcfSet);

    bam_hdr_t hdr;
    memset(&hdr, 0, sizeof(hdr));
    char *tname = strdup(in->chr);
    hdr.n_targets = 1; hdr.target_name = &tname;

    std::map<int, MultiGenomeVar>::iterator firstVariantIter = currentVariants.begin();
    std::map<int, MultiGenomeVar>::iterator firstVariantIter2 = currentVariants.begin();
    std::map<int, MultiGenomeVar>::reverse_iterator last = currentVariants.rbegin();

    std::vector<uint8_t> cat;
    std::vector<int32_t> hp, ps, pq, h1, h2, nps, c_pos, c_hp, p_ps, p_cnt;
    std::vector<uint64_t> c_off(1, 0), p_off(1, 0);
    std::vector<uint8_t> data;
    bam1_t aln;
    memset(&aln, 0, sizeof(aln));
    double t0 = now_s();
    for (int32_t r = 0; r < b.n_reads; r++) {
        int flag = b.flag[r];
        int category;
        if (b.mapq[r] < cfg.qualityThreshold && proc.mappingQualityFilter) { category = LPS_TAG_LOW_MAPQ; proc.processLowMappingQuality(); }
        else if ((flag & 0x4) != 0) { category = LPS_TAG_UNMAPPED; proc.processUnmappedRead(); }
        else if ((flag & 0x100) != 0) { category = LPS_TAG_SECONDARY; proc.processSecondaryAlignment(); }
        else if ((flag & 0x800) != 0 && cfg.tagSupplementary == false) { category = LPS_TAG_SUPPLEMENTARY; proc.processSupplementaryAlignment(); }
        else if (last == currentVariants.rend()) { category = LPS_TAG_EMPTY_VARIANTS; proc.processEmptyVariants(); }
        else if (int(b.ref_start[r]) <= (*last).first) category = LPS_TAG_PROCESSED;
        else { category = LPS_TAG_OTHER; proc.processOtherCase(); }
        cat.push_back((uint8_t)category);
        int o_hp = 0, o_ps = 0, o_pq = 0, o_h1 = 0, o_h2 = 0, o_nps = 0;
        if (category == LPS_TAG_PROCESSED) {
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
            // (1) the reference's own judgeHaplotype (HaplotagProcess.cpp:363-438)
            int pqValue = 0, psValue = 0;
            int haplotype = proc.judgeHaplotype(hdr, aln, chr, cfg.percentageThreshold, nullptr, pqValue, psValue, NORMAL, ref_string, cfg,
                                                firstVariantIter, currentVariants, vcfSet);
            // the counters processRead keeps (HaplotagProcess.cpp:318-354); the aux-tag edits need a heap-owned record and are skipped
            if ((aln.core.flag & 0x800) != 0) proc.localReadStats.totalSupplementary++;
            if (haplotype != ReadHP::unTag) { proc.localReadStats.totalHpCount[haplotype]++; proc.localReadStats.totalTagCount++; }
            else { proc.localReadStats.totalHpCount[ReadHP::unTag]++; proc.localReadStats.totalUnTagCount++; }
            proc.localReadStats.totalAlignment++;
            // (2) the same reference objects once more, to observe hpCount / variantsHP / countPS
            std::map<int, int> hpCount, variantsHP, countPS;
            hpCount[SnpHP::GERMLINE_H1] = 0; hpCount[SnpHP::GERMLINE_H2] = 0;
            int ref_pos = aln.core.pos, query_pos = 0;
            CigarParserContext cctx(aln, hdr, chr, cfg, firstVariantIter2, currentVariants, ref_string);
            CigarParser *parser = new GermlineHaplotagCigarParser(cctx, ref_pos, query_pos);
            parser->parsingCigar(hpCount, variantsHP, countPS);
            delete parser;
            o_hp = haplotype; o_pq = pqValue; o_ps = haplotype != ReadHP::unTag ? psValue : 0;
            o_h1 = hpCount[SnpHP::GERMLINE_H1]; o_h2 = hpCount[SnpHP::GERMLINE_H2]; o_nps = (int)countPS.size();
            for (auto &kv : variantsHP) { c_pos.push_back(kv.first); c_hp.push_back(kv.second); }
            for (auto &kv : countPS) { p_ps.push_back(kv.first); p_cnt.push_back(kv.second); }
        }
        hp.push_back(o_hp); ps.push_back(o_ps); pq.push_back(o_pq); h1.push_back(o_h1); h2.push_back(o_h2); nps.push_back(o_nps);
        c_off.push_back(c_pos.size()); p_off.push_back(p_ps.size());
    }
    out->t_total = now_s() - t0;
    out->n_reads = b.n_reads;
    out->category = dupv(cat); out->hp = dupv(hp); out->ps = dupv(ps); out->pq = dupv(pq); out->h1 = dupv(h1); out->h2 = dupv(h2);
    out->n_ps = dupv(nps);
    out->var_off = dupv(c_off); out->var_pos = dupv(c_pos); out->var_hp = dupv(c_hp);
    out->ps_off = dupv(p_off); out->ps_id = dupv(p_ps); out->ps_count = dupv(p_cnt);
    const ReadStatistics &s = proc.localReadStats;
    int64_t st[14] = {s.totalAlignment, s.totalSupplementary, s.totalSecondary, s.totalUnmapped, s.totalTagCount, s.totalUnTagCount,
                      s.totalLowerQuality, s.totalOtherCase, s.totalEmptyVariant, s.totalHighSimilarity, s.totalWithOutVaraint, 0, 0, 0};
    auto get = [&](int k) { auto it = s.totalHpCount.find(k); return it == s.totalHpCount.end() ? 0 : it->second; };
    st[11] = get(ReadHP::H1); st[12] = get(ReadHP::H2); st[13] = get(ReadHP::unTag);
    memcpy(out->stats, st, sizeof(st));
    free(tname);
    return 0;
}

extern "C" void ref_tap_tag_free(tap_tag_out *o) {
    free(o->category); free(o->hp); free(o->ps); free(o->pq); free(o->h1); free(o->h2); free(o->n_ps);
    free(o->var_off); free(o->var_pos); free(o->var_hp); free(o->ps_off); free(o->ps_id); free(o->ps_count);
    memset(o, 0, sizeof(*o));
}
