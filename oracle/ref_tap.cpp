/*
 * ref_tap.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin driver of OUR OWN that links against the UNMODIFIED reference objects (compiled in
 * place from /root/reference/src by oracle/Makefile) and replays the per-contig body of
 * PhasingProcess (src/phase/PhasingProcess.cpp:128-158) on an in-memory batch of alignments in
 * the SoA layout of include/lps.h, so that the reference's own get_snp / filterSNP / Clip /
 * addEdge / edgeConnectResult / readCorrection / exportResult can be observed stage by stage
 * and timed without BGZF I/O.  The reference has no dump of per-read calls or edge floats
 * (SURVEY.md §4), so the private members are read through a test-only `#define private public`
 * taken AFTER the standard and htslib headers have been included.
 *
 * The in-memory bam1_t records are built from the SoA batch; the read filter of
 * BamParser::direct_detect_alleles (ParsingBam.cpp:1282-1291) is applied by the driver because
 * that function itself can only read from a BAM file.
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
#include <unistd.h>

#define private public
#define protected public
#include "phase/PhasingGraph.h"
#include "phase/ParsingBam.h"
#undef private
#undef protected

#include "../include/lps.h"
#include "ref_tap.h"

namespace {

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename T> T *dup_vec(const std::vector<T> &v) {
    T *p = (T *)malloc(sizeof(T) * (v.size() + 1));
    if (!v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

void dump_stage(const std::vector<ReadVariant> &rv, tap_calls *out) {
    std::vector<int32_t> idx, pos, allele, quality;
    std::vector<uint64_t> off;
    off.push_back(0);
    for (const auto &r : rv) {
        idx.push_back(r.mapping_quality);
        for (const auto &v : r.variantVec) {
            pos.push_back(v.position);
            allele.push_back(v.allele);
            quality.push_back(v.quality);
        }
        off.push_back(pos.size());
    }
    out->n_aln = (int32_t)rv.size();
    out->read_idx = dup_vec(idx);
    out->off = dup_vec(off);
    out->pos = dup_vec(pos);
    out->allele = dup_vec(allele);
    out->quality = dup_vec(quality);
}

void free_stage(tap_calls *c) {
    free(c->read_idx); free(c->off); free(c->pos); free(c->allele); free(c->quality);
}

void dump_nodes(VairiantGraph &g, tap_nodes *out) {
    std::vector<int32_t> pos, type, ps, hr, ha;
    for (auto &kv : *g.totalVariantInfo) {
        int p = kv.first;
        pos.push_back(p);
        auto t = g.variantType->find(p);
        type.push_back(t == g.variantType->end() ? -1 : t->second);
        auto b = g.bkResult->find(std::make_pair(p, 1));
        ps.push_back(b == g.bkResult->end() ? 0 : b->second);
        auto r = g.subNodeHP->find(std::make_pair(p, 1));
        auto a = g.subNodeHP->find(std::make_pair(p, 2));
        hr.push_back(r == g.subNodeHP->end() ? -1 : r->second);
        ha.push_back(a == g.subNodeHP->end() ? -1 : a->second);
    }
    out->n = (int32_t)pos.size();
    out->pos = dup_vec(pos); out->type = dup_vec(type); out->ps = dup_vec(ps);
    out->hap_ref = dup_vec(hr); out->hap_alt = dup_vec(ha);
}
void free_nodes(tap_nodes *n) { free(n->pos); free(n->type); free(n->ps); free(n->hap_ref); free(n->hap_alt); }

}  // namespace

extern "C" int ref_tap_phase(const tap_phase_in *in, tap_phase_out *out) {
    memset(out, 0, sizeof(*out));
    const lps_read_batch &b = in->batch;
    std::string chr = in->chr;

    // ---- parameters (defaults: src/phase/Phasing.cpp:88-116) ----
    PhasingParameters params;
    params.numThreads = 1;
    params.distance = in->p.distance;
    params.svFile = "";
    params.modFile = "";
    params.fastaFile = "";
    params.resultPrefix = "/tmp/ref_tap";
    params.generateDot = false;
    params.isONT = in->p.is_ont != 0;
    params.isPB = !params.isONT;
    params.phaseIndel = true;
    params.indelQuality = 0;
    params.connectAdjacent = in->p.connect_adjacent;
    params.mappingQuality = in->p.mapping_quality;
    params.mismatchRate = 3;
    params.baseQuality = in->p.base_quality;
    params.edgeWeight = in->p.edge_weight;
    params.snpConfidence = in->p.snp_confidence;
    params.readConfidence = in->p.read_confidence;
    params.edgeThreshold = in->p.edge_threshold;
    params.overlapThreshold = in->p.overlap_threshold;
    params.deepsomaticOutput = false;
    params.svWindow = 20;
    params.svThreshold = 0.1;

    // ---- a header-only VCF so that the reference's own SnpParser can be constructed ----
    char vcf_path[256];
    snprintf(vcf_path, sizeof(vcf_path), "/tmp/ref_tap_%d_%p.vcf", (int)getpid(), (const void *)in);
    {
        FILE *f = fopen(vcf_path, "w");
        if (!f) return -1;
        fprintf(f, "##fileformat=VCFv4.2\n##contig=<ID=%s,length=%lld>\n", in->chr, (long long)in->ref_len);
        fprintf(f, "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n");
        fprintf(f, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n");
        fclose(f);
    }
    params.snpFile = vcf_path;
    SnpParser snpFile(params);
    unlink(vcf_path);
    // variant table exactly as SnpParser would hold it after parsing (ParsingBam.cpp:284-296)
    {
        auto &m = (*snpFile.chrVariant)[chr];
        for (int i = 0; i < in->n_var; i++) {
            RefAlt ra;
            const char *s = in->var_str + in->var_str_off[i];
            ra.Ref = s;
            ra.Alt = s + ra.Ref.size() + 1;
            ra.is_reverse = false; ra.is_modify = false; ra.is_danger = false;
            m[in->var_pos[i]] = ra;
        }
    }
    SVParser svFile(params, snpFile);
    METHParser modFile(params, snpFile, svFile);

    std::string chr_reference(in->ref, (size_t)in->ref_len);
    if (!in->p.have_reference) chr_reference = "";

    bam_hdr_t hdr;
    memset(&hdr, 0, sizeof(hdr));
    char *tname = strdup(in->chr);
    hdr.n_targets = 1;
    hdr.target_name = &tname;

    int lastSNPpos = snpFile.getLastSNP(chr);
    if (lastSNPpos == -1) { free(tname); return -2; }
    // stop_after_calls == 2: timing mode (bench.py's reference arm) - the whole path, nothing flattened for the caller; only the
    // timers around the reference's own calls and the counts are returned
    const bool timing = in->stop_after_calls == 2;

    // ---- stage A: get_snp over the batch ----
    double t0 = now_s();
    BamParser *bamParser = new BamParser(chr, params.bamFile, snpFile, svFile, modFile, chr_reference);
    std::vector<ReadVariant> readVariantVec;
    ClipCount clipCount;
    std::vector<uint8_t> data;
    bam1_t aln;
    memset(&aln, 0, sizeof(aln));
    for (int32_t r = 0; r < b.n_reads; r++) {
        // sam_itr_querys region "chr:1-lastSNP" (ParsingBam.cpp:1273): alignments starting at or
        // beyond lastSNPpos (0-based) do not overlap [0, lastSNPpos) and are never returned.
        if (b.ref_start[r] >= lastSNPpos) continue;
        int flag = b.flag[r];
        if (b.mapq[r] < params.mappingQuality || (flag & 0x4) != 0 || (flag & 0x100) != 0 || (flag & 0x400) != 0) continue;
        const char *name = in->names + (size_t)r * in->name_stride;
        size_t ln = strlen(name) + 1;
        size_t lnp = (ln + 3) & ~(size_t)3;
        size_t nbytes = lnp + 4 * (size_t)b.n_cigar[r] + ((size_t)b.l_qseq[r] + 1) / 2 + (size_t)b.l_qseq[r];
        data.assign(nbytes + 8, 0);
        memcpy(data.data(), name, ln);
        memcpy(data.data() + lnp, b.cigar + b.cigar_off[r], 4 * (size_t)b.n_cigar[r]);
        memcpy(data.data() + lnp + 4 * (size_t)b.n_cigar[r], b.seq4 + b.seq_off[r], ((size_t)b.l_qseq[r] + 1) / 2);
        memcpy(data.data() + lnp + 4 * (size_t)b.n_cigar[r] + ((size_t)b.l_qseq[r] + 1) / 2, b.qual + b.qual_off[r], (size_t)b.l_qseq[r]);
        aln.data = data.data();
        aln.l_data = (int)nbytes;
        aln.m_data = (uint32_t)data.size();
        aln.core.pos = b.ref_start[r];
        aln.core.tid = 0;
        aln.core.qual = b.mapq[r];
        aln.core.flag = b.flag[r];
        aln.core.l_qname = (uint16_t)lnp;
        aln.core.l_extranul = (uint8_t)(lnp - ln);
        aln.core.n_cigar = b.n_cigar[r];
        aln.core.l_qseq = b.l_qseq[r];
        size_t before = readVariantVec.size();
        bamParser->get_snp(hdr, aln, readVariantVec, clipCount, chr_reference, params.isONT, params.svWindow, params.svThreshold);
        if (readVariantVec.size() > before) readVariantVec.back().mapping_quality = r;  // stash the batch index
    }
    delete bamParser;
    out->t_get_snp = now_s() - t0;
    if (timing) {
        out->stage_a.n_aln = (int32_t)readVariantVec.size();
        for (auto &r : readVariantVec) out->n_result_calls += (int64_t)r.variantVec.size();
    } else dump_stage(readVariantVec, &out->stage_a);
    if (!timing) {
        std::vector<int32_t> p, f, k;
        for (auto &kv : clipCount) {
            p.push_back(kv.first);
            auto fi = kv.second.find(FRONT), bi = kv.second.find(BACK);
            f.push_back(fi == kv.second.end() ? 0 : fi->second);
            k.push_back(bi == kv.second.end() ? 0 : bi->second);
        }
        out->n_clips = (int32_t)p.size();
        out->clip_pos = dup_vec(p); out->clip_front = dup_vec(f); out->clip_back = dup_vec(k);
    }

    // ---- stage B: filterSNP (ONT) ----
    t0 = now_s();
    if (params.isONT) snpFile.filterSNP(chr, readVariantVec, chr_reference);
    out->t_filter_snp = now_s() - t0;
    if (!timing) dump_stage(readVariantVec, &out->stage_b);
    if (readVariantVec.empty() || in->stop_after_calls == 1) { free(tname); return 0; }

    // the reference dereferences front()/back() of EMPTY variantVecs in addEdge (UB); report them
    for (auto &r : readVariantVec) if (r.variantVec.empty()) out->n_empty_after_filter++;

    // ---- Clip / CNV (reference segfaults on an empty clip map, PhasingGraph.cpp:1134) ----
    if (clipCount.empty()) { free(tname); return -3; }
    t0 = now_s();
    Clip *clip = new Clip(chr, clipCount);
    clip->getCNVInterval(clipCount, chr);
    out->t_clip = now_s() - t0;
    if (!timing) {
        std::vector<int32_t> s, e;
        for (auto &c : clip->cnvVec) { s.push_back(c.first); e.push_back(c.second); }
        out->n_cnv = (int32_t)s.size();
        out->cnv_start = dup_vec(s); out->cnv_end = dup_vec(e);
    }

    // ---- stage C: addEdge ----
    t0 = now_s();
    VairiantGraph *g = new VairiantGraph(chr_reference, params, chr);
    g->addEdge(readVariantVec, *clip);
    out->t_add_edge = now_s() - t0;
    if (!timing) dump_stage(readVariantVec, &out->stage_c);
    if (!timing) {
        std::vector<int32_t> ap, bp;
        std::vector<uint8_t> which;
        std::vector<float> val;
        uint64_t contrib = 0;
        for (auto &kv : *g->edgeList) {
            VariantEdge *e = kv.second;
            contrib += (uint64_t)e->ref->readCount + (uint64_t)e->alt->readCount;
            for (auto &c : *e->ref->refReadCount) { ap.push_back(kv.first); bp.push_back(c.first); which.push_back(0); val.push_back(c.second); }
            for (auto &c : *e->ref->altReadCount) { ap.push_back(kv.first); bp.push_back(c.first); which.push_back(1); val.push_back(c.second); }
            for (auto &c : *e->alt->refReadCount) { ap.push_back(kv.first); bp.push_back(c.first); which.push_back(2); val.push_back(c.second); }
            for (auto &c : *e->alt->altReadCount) { ap.push_back(kv.first); bp.push_back(c.first); which.push_back(3); val.push_back(c.second); }
        }
        out->n_edge_nodes = (int32_t)g->edgeList->size();
        out->n_cells = (int64_t)ap.size();
        out->n_contrib = contrib;
        out->cell_a = dup_vec(ap); out->cell_b = dup_vec(bp); out->cell_which = dup_vec(which); out->cell_val = dup_vec(val);
    }

    // ---- sweep, then read correction (VairiantGraph::phasingProcess split in two) ----
    t0 = now_s();
    g->edgeConnectResult();
    out->t_sweep = now_s() - t0;
    if (!timing) dump_nodes(*g, &out->nodes_sweep);
    t0 = now_s();
    g->readCorrection();
    out->t_read_correction = now_s() - t0;
    if (!timing) dump_nodes(*g, &out->nodes_final);
    if (!timing) {
        std::vector<int32_t> hp;
        for (auto &r : readVariantVec) {
            auto it = g->readHpMap->find(r.read_name);
            hp.push_back(it == g->readHpMap->end() ? -2 : it->second);
        }
        out->read_hp = dup_vec(hp);
    }
    PhasingResult result;
    t0 = now_s();
    g->exportResult(chr, result);
    out->t_export = now_s() - t0;
    out->n_result = (int32_t)result.size();
    if (!timing) {
        std::vector<int32_t> pos, blk, h1, h2;
        for (auto &kv : result) {
            pos.push_back(std::stoi(kv.first.substr(chr.size() + 1)));
            blk.push_back(kv.second.block);
            int a = -9, c = -9;
            sscanf(kv.second.RAstatus.c_str(), "%d|%d", &a, &c);
            h1.push_back(a); h2.push_back(c);
        }
        // PhasingResult is keyed by string; sort by position for the caller
        std::vector<int> order(pos.size());
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int x, int y) { return pos[x] < pos[y]; });
        std::vector<int32_t> p2, b2, a2, c2;
        for (int i : order) { p2.push_back(pos[i]); b2.push_back(blk[i]); a2.push_back(h1[i]); c2.push_back(h2[i]); }
        out->n_result = (int32_t)p2.size();
        out->res_pos = dup_vec(p2); out->res_block = dup_vec(b2); out->res_hap_ref = dup_vec(a2); out->res_hap_alt = dup_vec(c2);
    }
    g->destroy();
    delete g;
    delete clip;
    free(tname);
    return 0;
}

extern "C" void ref_tap_phase_free(tap_phase_out *o) {
    free_stage(&o->stage_a); free_stage(&o->stage_b); free_stage(&o->stage_c);
    free(o->clip_pos); free(o->clip_front); free(o->clip_back);
    free(o->cnv_start); free(o->cnv_end);
    free(o->cell_a); free(o->cell_b); free(o->cell_which); free(o->cell_val);
    free_nodes(&o->nodes_sweep); free_nodes(&o->nodes_final);
    free(o->read_hp);
    free(o->res_pos); free(o->res_block); free(o->res_hap_ref); free(o->res_hap_alt);
    memset(o, 0, sizeof(*o));
}

/* homopolymerLength of the reference (src/shared/Util.cpp:21-54), exposed for the oracle tests */
extern "C" int ref_tap_homopolymer(const char *ref, int64_t len, int pos) {
    std::string s(ref, (size_t)len);
    return homopolymerLength(pos, s);
}
