// somatic_host.cpp — the `somatic_haplotag` sub-command above the C ABI: options, the NORMAL (phased) and TUMOR VCF loaders and
// their union map, the two extract passes over the normal and the tumor BAM, tumor purity (+ <prefix>_purity.out), the somatic
// calling stage, and the tagging pass over the tumor BAM (HP:Z, optional PS:i, PQ:i) with the @PG line and the stderr report.
//
// Reference seams replaced (file:line relative to the reference tree):
//   SomaticHaplotagMain / option handling          src/somatic_haplotag/SomaticHaplotag.cpp:34-142
//   SomaticHaplotagProcess::pipelineProcess        src/somatic_haplotag/SomaticHaplotagProcess.cpp:51-103, 105-235
//   SomaticVarCaller::variantCalling               src/somatic_haplotag/SomaticVarCaller.cpp:816-949  (extract passes, purity, calling, getSomaticFlag :2397-2412)
//   TumorPurityEstimator::writePurityResult        src/somatic_haplotag/TumorPurityEstimator.cpp:375-424
//   SomaticHaplotagChrProcessor::processRead / addAuxiliaryTags   src/somatic_haplotag/SomaticHaplotagProcess.cpp:310-400, 464-472
// The device judges whole batches (lps_extract_normal / lps_extract_tumor / lps_somatic_tag_reads); purity and calling are the
// host stages of liblps_b200.so (lps_estimate_purity, lps_somatic_call).
//   SomaticVarCaller::writeSomaticVarCallingLog     src/somatic_haplotag/SomaticVarCaller.cpp:1576-1926  (<prefix>_somatic_var.out)
// Scope notes: --log, --truth-vcf, --truth-bed, --sv-file and --mod-file are parsed but rejected (benchmark tooling and text logs
// outside the rebuilt hot path); --somatic-calling-log writes <prefix>_somatic_var.out, the file of BASELINE.md's parity
// gate, and none of the other debugging logs that switch produces in the reference.
#include "host_common.h"

#include <getopt.h>
#include <omp.h>

#include <climits>
#include <cmath>
#include <ctime>

namespace {

const char *SOM_USAGE =
    "Usage:  somatic_haplotag [OPTION] ... READSFILE\n"
    "      --help                          display this help and exit.\n\n"
    "required arguments:\n"
    "      -s, --snp-file=NAME             input phased normal sample SNP VCF file.\n"
    "      -b, --bam-file=NAME             input normal sample BAM file.\n"
    "      --tumor-snv-file=NAME           input tumor sample SNV VCF file.\n"
    "      --tumor-bam-file=NAME           input tumor sample BAM file for somatic haplotag.\n"
    "      -r, --reference=NAME            reference FASTA.\n\n"
    "optional arguments:\n"
    "      --tagSupplementary              tag supplementary alignment. default:false\n"
    "      -q, --qualityThreshold=Num      not tag alignment if the mapping quality less than threshold. default:1\n"
    "      -p, --percentageThreshold=Num   share of the alleles the winning haplotype needs. default:0.6\n"
    "      -t, --threads=Num               number of thread. default:1\n"
    "      -o, --out-prefix=NAME           prefix of the tagged tumor BAM. default:result\n"
    "      --region=REGION                 chrom | chrom:start | chrom:start-end. default:\"\"(all regions)\n"
    "      --tumor-purity=Num              tumor purity (0.1~1.0). default: automatic estimation.\n"
    "      --disableFilter                 accept all tumor VCF variants as somatic. default: false.\n"
    "      --output-somatic-vcf            write <prefix>_sc.vcf: the tumor VCF's SNP / indel records, FILTER PASS for called somatic\n"
    "                                      variants and LowQual for the others. default: false.\n"
    "      --cram                          the output file will be in the cram format. default:bam\n"
    "not available in this build: --log, --truth-vcf, --truth-bed, --sv-file, --mod-file\n";

enum { S_HELP = 1, S_SUP, S_SV, S_MOD, S_REGION, S_CRAM, S_LOG, S_TUM_SNP, S_TUM_BAM, S_DISABLE_FILTER, S_PURITY, S_OUT_VCF, S_CALL_LOG,
       S_TRUTH_VCF, S_TRUTH_BED, S_BENCH_LOG };

const struct option SOM_LONG[] = {
    {"help", no_argument, NULL, S_HELP},
    {"snp-file", required_argument, NULL, 's'},
    {"bam-file", required_argument, NULL, 'b'},
    {"reference", required_argument, NULL, 'r'},
    {"sv-file", required_argument, NULL, S_SV},
    {"mod-file", required_argument, NULL, S_MOD},
    {"threads", required_argument, NULL, 't'},
    {"qualityThreshold", required_argument, NULL, 'q'},
    {"percentageThreshold", required_argument, NULL, 'p'},
    {"tagSupplementary", no_argument, NULL, S_SUP},
    {"out-prefix", required_argument, NULL, 'o'},
    {"region", required_argument, NULL, S_REGION},
    {"cram", no_argument, NULL, S_CRAM},
    {"log", no_argument, NULL, S_LOG},
    {"tumor-snv-file", required_argument, NULL, S_TUM_SNP},
    {"tumor-bam-file", required_argument, NULL, S_TUM_BAM},
    {"disableFilter", no_argument, NULL, S_DISABLE_FILTER},
    {"tumor-purity", required_argument, NULL, S_PURITY},
    {"output-somatic-vcf", no_argument, NULL, S_OUT_VCF},
    {"somatic-calling-log", no_argument, NULL, S_CALL_LOG},
    {"truth-vcf", required_argument, NULL, S_TRUTH_VCF},
    {"truth-bed", required_argument, NULL, S_TRUTH_BED},
    {"benchmark-log", no_argument, NULL, S_BENCH_LOG},
    {NULL, 0, NULL, 0}};

struct SomOptions {
    int threads = 1, quality = 1;
    double percentage = 0.6, purity = 0.2;
    bool tag_supplementary = false, estimate_purity = true, enable_filter = true, unsupported = false;
    bool cram = false;
    bool write_sc_vcf = false;  // --output-somatic-vcf: <prefix>_sc.vcf
    bool write_calling_log = false;   // --somatic-calling-log: <prefix>_somatic_var.out (the other debugging logs of that switch are not written)
    bool purity_only = false;   // the `estimate_purity` sub-command (PurityEstimation.cpp): extract passes + purity, no calling, no tagging
    std::string snp_file, bam, tumor_vcf, tumor_bam, fasta, prefix = "result", region, command = "longphase-s ";
};

// one position of the union map std::map<int, MultiGenomeVar> (HaplotagType.h:146-162)
struct UnionVar {
    bool has_nor = false, has_tum = false;
    lpsh::SampleRecord nor, tum;
    uint8_t is_somatic = 0;   // MultiGenomeVar::isSomaticVariant
    int8_t derive_hp = 0;     // MultiGenomeVar::somaticReadDeriveByHP
};

// deep copy of what lps_somatic_call and lps_estimate_purity read from an lps_extract_result (the context owns the original
// only until its next call)
struct ExtractCopy {
    bool set = false;
    std::vector<int32_t> tum_var, pos_base, read_hp_count, somatic_read_hp_count, window_hist, case_read_count, case_count, h1, h2, h3, end_pos;
    std::vector<float> ratios_f;
    std::vector<double> ratios_d;
    std::vector<uint8_t> category, n_ps;
    std::vector<int8_t> read_hp;
    std::vector<uint64_t> call_off;
    std::vector<lps_call> calls;
    template <class T>
    static void put(std::vector<T> &dst, const T *src, size_t n) { if (src) dst.assign(src, src + n); else dst.clear(); }
    void assign(const lps_extract_result &r) {
        const size_t nt = (size_t)r.n_tum, nr = (size_t)r.reads.n_reads;
        put(tum_var, r.tum_var, nt); put(pos_base, r.pos_base, nt * LPS_PB_FIELDS); put(read_hp_count, r.read_hp_count, nt * 9);
        put(somatic_read_hp_count, r.somatic_read_hp_count, nt * 9); put(window_hist, r.window_hist, nt * 2 * LPS_WINDOW_BINS);
        put(case_read_count, r.case_read_count, nt); put(case_count, r.case_count, nt * LPS_CASE_FIELDS); put(ratios_f, r.ratios_f, nt * LPS_RF_FIELDS); put(ratios_d, r.ratios_d, nt * LPS_RD_FIELDS);
        put(category, r.reads.category, nr); put(read_hp, r.reads.read_hp, nr); put(h1, r.reads.h1, nr); put(h2, r.reads.h2, nr);
        put(h3, r.reads.h3, nr); put(n_ps, r.reads.n_ps, nr); put(end_pos, r.reads.end_pos, nr);
        put(call_off, r.call_off, r.call_off ? nr + 1 : 0); put(calls, r.calls, r.calls ? (size_t)r.n_calls : 0);
        set = true;
    }
    template <class T>
    static const T *ptr(const std::vector<T> &v) { return v.empty() ? nullptr : v.data(); }
    lps_extract_result view() const {
        lps_extract_result r;
        memset(&r, 0, sizeof(r));
        r.n_tum = (int32_t)tum_var.size();
        r.tum_var = ptr(tum_var); r.pos_base = ptr(pos_base); r.read_hp_count = ptr(read_hp_count);
        r.reads.n_reads = (int32_t)category.size();
        r.reads.category = ptr(category); r.reads.read_hp = ptr(read_hp); r.reads.h1 = ptr(h1); r.reads.h2 = ptr(h2); r.reads.h3 = ptr(h3);
        r.reads.n_ps = ptr(n_ps); r.reads.end_pos = ptr(end_pos);
        r.somatic_read_hp_count = ptr(somatic_read_hp_count); r.window_hist = ptr(window_hist); r.case_read_count = ptr(case_read_count);
        r.ratios_f = ptr(ratios_f); r.ratios_d = ptr(ratios_d); r.case_count = ptr(case_count);
        r.n_calls = calls.size(); r.call_off = ptr(call_off); r.calls = ptr(calls);
        return r;
    }
};

struct TumorArrays {   // lps_tumor_variants of one contig
    std::vector<uint8_t> nor_present, tum_present, ref0, alt0, gt_kind, hp1_is_alt, is_somatic;
    std::vector<uint16_t> ref_len, alt_len;
    std::vector<int32_t> ps;
    std::vector<int8_t> derive_hp;
    void view(lps_tumor_variants *t) const {
        memset(t, 0, sizeof(*t));
        t->n = (int32_t)tum_present.size();
        t->nor_present = nor_present.data(); t->tum_present = tum_present.data(); t->ref0 = ref0.data(); t->alt0 = alt0.data();
        t->ref_len = ref_len.data(); t->alt_len = alt_len.data(); t->gt_kind = gt_kind.data(); t->hp1_is_alt = hp1_is_alt.data();
        t->ps = ps.data(); t->is_somatic = is_somatic.data(); t->derive_hp = derive_hp.data();
    }
};

struct ContigState {
    ExtractCopy normal, tumor;
    // per tumor slot, kept for the calling log: lps_somatic_call's products and SomaticData::statisticPurity
    std::vector<uint8_t> is_somatic, is_filter_out, in_dense, used_for_purity;
    std::vector<float> mean_alt, z_score;
    std::vector<int32_t> interval_snp_count, min_distance;
};

}  // namespace

struct lpsh_som {
    SomOptions opt;
    std::vector<std::string> chr_names;
    std::map<std::string, int> chr_length;
    std::map<std::string, std::map<int, UnionVar>> variants;
    std::map<std::string, std::string> ref_normal, ref_tumor;   // FastaParser strings of the NORMAL / TUMOR passes (different last positions)
    bool have_reference = false;
    std::vector<ContigState> contig;
    double purity = 0.0;
    lps_purity_result purity_result;
    int64_t n_somatic = 0;
    // whole-contig pack of an extract pass
    lpsh::PackedContig pack;
    TumorArrays tum;
    // tagging pass
    lpsh::TagBamIO io;          // tumor BAM in, tagged BAM out, index, thread pool, region reader
    lpsh::Chunk chunk;          // the chunk of the staged API (lpsh_som_tag_pack / lpsh_som_tag_emit)
    int chunk_contig = -1;
    size_t chunk_reads = 8192;
    // ReadStatistics
    int64_t st_alignment = 0, st_supplementary = 0, st_secondary = 0, st_unmapped = 0, st_tag = 0, st_untag = 0, st_low = 0, st_other = 0,
            st_empty = 0, st_similar = 0, st_cross = 0, st_no_variant = 0, st_only_h3 = 0, st_hp[LPS_READHP_FIELDS] = {0};
    std::time_t t_begin = time(NULL);
};

namespace {

const char *READ_HP_TEXT[LPS_READHP_FIELDS] = {".", "1", "2", "3", "4", "1-1", "1-2", "2-1", "2-2"};   // ReadHapUtil::readHapIntToString

int parse_som_options(int argc, char **argv, SomOptions &o) {
    optind = 1;
    bool bad = false;
    if (argc > 0 && std::string(argv[0]) == "estimate_purity") {   // ParamsHandler<PurityEstimParameters>::initialize (PurityEstimation.cpp:37-41)
        o.purity_only = true;
        o.quality = 20;
        o.tag_supplementary = true;
    }
    for (int c; (c = getopt_long(argc, argv, "s:b:o:t:q:p:r:", SOM_LONG, NULL)) != -1;) {
        switch (c) {
            case 't': lpsh::take(optarg, o.threads); break;
            case 'o': lpsh::take(optarg, o.prefix); break;
            case 'q': lpsh::take(optarg, o.quality); break;
            case 'p': lpsh::take(optarg, o.percentage); break;
            case S_SUP: o.tag_supplementary = true; break;
            case S_REGION: lpsh::take(optarg, o.region); break;
            case 's': lpsh::take(optarg, o.snp_file); break;
            case 'b': lpsh::take(optarg, o.bam); break;
            case 'r': lpsh::take(optarg, o.fasta); break;
            case S_TUM_SNP: lpsh::take(optarg, o.tumor_vcf); break;
            case S_TUM_BAM: lpsh::take(optarg, o.tumor_bam); break;
            case S_DISABLE_FILTER: if (o.purity_only) bad = true; else o.enable_filter = false; break;
            case S_PURITY: if (o.purity_only) bad = true; else { lpsh::take(optarg, o.purity); o.estimate_purity = false; } break;
            case S_OUT_VCF: if (o.purity_only) bad = true; else o.write_sc_vcf = true; break;
            case S_CRAM: if (o.purity_only) bad = true; else o.cram = true; break;
            case S_CALL_LOG: o.write_calling_log = true; break;
            case S_LOG: case S_SV: case S_MOD: case S_TRUTH_VCF: case S_TRUTH_BED: case S_BENCH_LOG:
                o.unsupported = true; break;
            case S_HELP: std::cout << SOM_USAGE << std::endl; return 2;
            default: bad = true;
        }
    }
    for (int i = 0; i < argc; i++) { o.command += argv[i]; o.command += " "; }
    const char *prog = o.purity_only ? "estimate_purity" : "somatic_haplotag";
    bad |= !lpsh::required_file(prog, o.snp_file, "SNP file");
    bad |= !lpsh::required_file(prog, o.bam, "BAM file");
    bad |= !lpsh::required_file(prog, o.fasta, "reference file");
    bad |= !lpsh::required_file(prog, o.tumor_vcf, "tumor SNV file");
    bad |= !lpsh::required_file(prog, o.tumor_bam, "tumor BAM file");
    if (o.threads < 1) { std::cerr << "[ERROR] " << prog << ": invalid threads. value: " << o.threads << "\nplease check -t, --threads=Num\n"; bad = true; }
    if (o.percentage > 1 || o.percentage < 0) {
        std::cerr << "[ERROR] " << prog << ": invalid percentage threshold. value: " << o.percentage
                  << "\nthis value need: 0~1, please check -p, --percentageThreshold=Num\n";
        bad = true;
    }
    if (o.purity < 0.1 || o.purity > 1.0) {
        std::cerr << "[ERROR] " << prog << ": invalid tumor purity. value: " << o.purity << "\nthis value need: 0.1~1.0, --tumor-purity=Number\n";
        bad = true;
    }
    if (o.unsupported) {
        std::cerr << "[ERROR] " << prog << ": --log, --truth-vcf, --truth-bed, --sv-file and --mod-file "
                     "are not available in this build.\n";
        bad = true;
    }
    if (bad) { std::cerr << "\n"; std::cout << SOM_USAGE << std::endl; return 1; }
    return 0;
}

void som_banner(const SomOptions &o) {   // SomaticHaplotagProcess::printParamsMessage (SomaticHaplotagProcess.cpp:13-49)
    std::ostream &e = std::cerr;
    if (o.purity_only) {                  // PurityEstimProcess::printParamsMessage (PurityEstimationProcess.cpp:9-29)
        e << "LongPhase-S v" << lpsh::REFERENCE_VERSION << " - Estimate Tumor Purity (" << lps_version() << ")\n\n[Input Files]\n";
        e << "phased normal SNP file       : " << o.snp_file << "\ntumor SNP file               : " << o.tumor_vcf << "\n";
        e << "normal BAM file              : " << o.bam << "\ntumor BAM file               : " << o.tumor_bam << "\n";
        e << "reference file               : " << o.fasta << "\n\n[Output Files]\npurity estimation file       :" << o.prefix + "_purity.out" << "\n";
        e << "-------------------------------------------\n[Purity Estimation Params] \n";
        e << "number of threads            : " << o.threads << "\nestimation region            : " << (!o.region.empty() ? o.region : "all") << "\n";
        e << "filter mapping quality below : " << o.quality << "\npercentage threshold         : " << o.percentage << "\n";
        e << "include supplementary reads  : " << (o.tag_supplementary ? "enabled" : "disabled") << "\n-------------------------------------------\n";
        return;
    }
    e << "LongPhase-S v" << lpsh::REFERENCE_VERSION << " - Somatic Haplotag (" << lps_version() << ")\n\n[Input Files]\n";
    e << "phased normal SNP file       : " << o.snp_file << "\ntumor SNP file               : " << o.tumor_vcf << "\n";
    e << "normal BAM file              : " << o.bam << "\ntumor BAM file               : " << o.tumor_bam << "\n";
    e << "reference file               : " << o.fasta << "\n\n[Output Files]\n";
    e << "tagged tumor BAM file        : " << o.prefix + (o.cram ? ".cram" : ".bam") << "\npurity estimation file       : " << (o.estimate_purity ? o.prefix + "_purity.out" : "") << "\n";
    e << "somatic calling VCF file     : " << (o.write_sc_vcf ? o.prefix + "_sc.vcf" : "") << "\n";
    e << "-------------------------------------------\n[Somatic Haplotagging Params] \n";
    e << "number of threads            : " << o.threads << "\ntag region                   : " << (!o.region.empty() ? o.region : "all") << "\n";
    e << "filter mapping quality below : " << o.quality << "\npercentage threshold         : " << o.percentage << "\n";
    e << "tag supplementary            : " << (o.tag_supplementary ? "enabled" : "disabled") << "\n\n[Somatic Variant Calling Params] \n";
    e << "mapping quality              : " << o.quality << "\ntumor purity value           : " << (o.estimate_purity ? "automatic estimation" : std::to_string(o.purity)) << "\n";
    e << "variant filtering            : " << (o.enable_filter ? "enabled" : "disabled") << "\n-------------------------------------------\n";
}

bool is_snp(const lpsh::SampleRecord &v) { return v.ref.size() == 1 && v.alt.size() == 1; }
bool is_ins(const lpsh::SampleRecord &v) { return v.ref.size() == 1 && v.alt.size() > 1; }
bool is_del(const lpsh::SampleRecord &v) { return v.ref.size() > 1 && v.alt.size() == 1; }

// parseVariantFiles + setChrVecAndChrLength + displaySnpCounts + setProcessingChromRegion (SomaticHaplotagProcess.cpp:105-235, HaplotagProcess.cpp:105-135)
int load_union(lpsh_som &job) {
    lpsh::SampleVcf nor, tum;
    std::time_t t0 = time(NULL);
    std::cerr << "parsing normal SNP VCF ... ";
    lpsh::load_sample_vcf(job.opt.snp_file, false, nor);
    std::cerr << difftime(time(NULL), t0) << "s\n";
    t0 = time(NULL);
    std::cerr << "parsing tumor SNP VCF ... ";
    lpsh::load_sample_vcf(job.opt.tumor_vcf, true, tum);
    std::cerr << difftime(time(NULL), t0) << "s\n";
    if (!job.opt.purity_only) for (const auto &c : tum.chr_length) {
        auto it = nor.chr_length.find(c.first);
        if (it == nor.chr_length.end()) { std::cerr << "[ERROR] (setChrVecAndChrLength) :tumor & normal VCFs chromosome count are not the same" << std::endl; return lpsh::fail("tumor & normal VCFs chromosome count are not the same"); }
        if (it->second != c.second) { std::cerr << "[ERROR] (setChrVecAndChrLength) :tumor & normal VCFs chromosome length are not the same => chr: " << c.first << std::endl; return lpsh::fail("tumor & normal VCFs chromosome length are not the same"); }
    }
    if (job.opt.purity_only) {            // PurityEstimProcess keeps HaplotagProcess::setChrVecAndChrLength: the NORMAL VCF's contigs
        job.chr_names = nor.chr_names;
        job.chr_length = nor.chr_length;
    } else if (tum.chr_names.empty()) {
        std::cerr << "[WARNING] tumor VCF chromosome count is empty" << std::endl;
        if (nor.chr_names.empty()) return lpsh::fail("tumor & normal VCFs chromosome count are empty");
        std::cerr << "[INFO] use normal VCF chromosome count" << std::endl;
        job.chr_names = nor.chr_names;
        job.chr_length = nor.chr_length;
    } else {
        job.chr_names = tum.chr_names;
        job.chr_length = tum.chr_length;
    }
    for (auto &c : nor.records) for (auto &kv : c.second) { UnionVar &u = job.variants[c.first][kv.first]; u.has_nor = true; u.nor = kv.second; }
    for (auto &c : tum.records) for (auto &kv : c.second) { UnionVar &u = job.variants[c.first][kv.first]; u.has_tum = true; u.tum = kv.second; }
    int n_nor = 0, n_snp = 0, n_both = 0, n_ins = 0, n_del = 0;
    for (const std::string &chr : job.chr_names)
        for (const auto &kv : job.variants[chr]) {
            const UnionVar &u = kv.second;
            if (u.has_tum) { n_snp += is_snp(u.tum); n_ins += is_ins(u.tum); n_del += is_del(u.tum); }
            n_nor += u.has_nor;
            n_both += u.has_nor && u.has_tum;
        }
    if (!job.opt.purity_only) std::cerr << "Normal SNP count: " << n_nor << "\nTumor SNP count: " << n_snp << "\nOverlap SNP count: " << n_both << "\nTumor Insert count: " << n_ins
              << "\nTumor Delete count: " << n_del << std::endl;
    if (!job.opt.region.empty()) {
        const size_t colon = job.opt.region.find(':');
        const std::string chr = colon != std::string::npos ? job.opt.region.substr(0, colon) : job.opt.region;
        if (std::find(job.chr_names.begin(), job.chr_names.end(), chr) == job.chr_names.end()) {
            std::cerr << "[ERROR] Incorrect chromosome for input region: " << chr << std::endl;
            exit(1);
        }
        job.chr_names.assign(1, chr);
    }
    for (auto it = job.variants.begin(); it != job.variants.end();) {
        if (std::find(job.chr_names.begin(), job.chr_names.end(), it->first) == job.chr_names.end()) it = job.variants.erase(it);
        else ++it;
    }
    return 0;
}

// getLastVarPos for both genome samples + FastaParser (HaplotagParsingBam.cpp:333-373, ParsingBam.cpp:17-59)
int load_som_reference(lpsh_som &job) {
    faidx_t *fai = fai_load(job.opt.fasta.c_str());
    if (!fai) return lpsh::fail("cannot load the FASTA index of " + job.opt.fasta);
    for (const std::string &chr : job.chr_names) {
        int last_nor = 0, last_any = 0;
        const auto &vars = job.variants[chr];
        if (!vars.empty()) last_any = vars.rbegin()->first;   // every entry holds a TUMOR record or a phased NORMAL record
        for (auto it = vars.rbegin(); it != vars.rend(); ++it) if (it->second.has_nor) { last_nor = it->first; break; }
        for (int pass = 0; pass < 2; pass++) {
            int len = 0;
            char *s = faidx_fetch_seq(fai, chr.c_str(), 0, (pass ? last_any : last_nor) + 5, &len);
            if (len == 0) std::cout << "nothing in reference file \n";
            (pass ? job.ref_tumor : job.ref_normal)[chr] = s ? s : "";
            free(s);
        }
    }
    fai_destroy(fai);
    job.have_reference = true;
    return 0;
}

// the union map of one contig as lps_variants (NORMAL side) + lps_tumor_variants
void pack_union(lpsh_som &job, const std::string &chr, lpsh::PackedContig &pc, TumorArrays &t) {
    pc.tagged_variants = true;
    t = TumorArrays();
    for (const auto &kv : job.variants[chr]) {
        const UnionVar &u = kv.second;
        const lpsh::SampleRecord &n = u.has_nor ? u.nor : u.tum;   // NORMAL columns of a tumor-only position are never read (nor_present = 0)
        pc.add_variant(kv.first, n.ref, n.alt);
        pc.v_hp1_is_alt.push_back(u.has_nor ? n.hp1_is_alt : 0);
        pc.v_ps.push_back(u.has_nor ? n.ps : 0);
        pc.v_gt_kind.push_back(1);
        const lpsh::SampleRecord &m = u.has_tum ? u.tum : u.nor;
        t.nor_present.push_back(u.has_nor); t.tum_present.push_back(u.has_tum);
        t.ref0.push_back(m.ref.empty() ? 0 : (uint8_t)m.ref[0]); t.alt0.push_back(m.alt.empty() ? 0 : (uint8_t)m.alt[0]);
        t.ref_len.push_back((uint16_t)std::min<size_t>(m.ref.size(), 65535)); t.alt_len.push_back((uint16_t)std::min<size_t>(m.alt.size(), 65535));
        t.gt_kind.push_back(u.has_tum ? (uint8_t)u.tum.gt_kind : 0);
        t.hp1_is_alt.push_back(u.has_tum ? u.tum.hp1_is_alt : 0);
        t.ps.push_back(u.has_tum ? u.tum.ps : -1);
        t.is_somatic.push_back(u.is_somatic); t.derive_hp.push_back(u.derive_hp);
    }
}

// VcfParser::writingResultVCF / writeProcess (HaplotagVcfParser.cpp:60-85, 548-614): the tumor VCF's header, then only the records whose
// TUMOR entry in the union map is a SNP, an insertion or a deletion, FILTER rewritten from the caller's verdict
void write_sc_line(lpsh_som &job, const std::string &line, bool &wrote_command, std::ostream &out) {
    if (line.size() >= 2 && line.compare(0, 2, "##") == 0) { out << line << std::endl; return; }
    if (line.size() >= 6 && (line.compare(0, 6, "#CHROM") == 0 || line.compare(0, 6, "#chrom") == 0)) {
        if (!wrote_command) {
            out << "##longphase_s_version=" << lpsh::REFERENCE_VERSION << std::endl << "##commandline=" << job.opt.command << std::endl;
            wrote_command = true;
        }
        out << line << std::endl;
        return;
    }
    std::istringstream split(line);
    std::vector<std::string> f((std::istream_iterator<std::string>(split)), std::istream_iterator<std::string>());
    if (f.empty()) return;
    if (f.size() < 7) { std::cerr << "[ERROR](VcfParser::writeProcess) => VCF file format error: " << line << std::endl; exit(EXIT_FAILURE); }
    auto chr = job.variants.find(f[0]);
    if (chr == job.variants.end()) return;
    auto hit = chr->second.find(std::stoi(f[1]) - 1);
    if (hit == chr->second.end() || !hit->second.has_tum) return;
    const UnionVar &u = hit->second;
    if (!(is_snp(u.tum) || is_ins(u.tum) || is_del(u.tum))) return;
    if (u.is_somatic) f[6] = "PASS";
    else if (f[6] == "PASS") f[6] = "LowQual";
    for (size_t i = 0; i < f.size(); i++) out << (i ? "\t" : "") << f[i];
    out << std::endl;
}

int write_sc_vcf(lpsh_som &job) {
    const std::string &path = job.opt.tumor_vcf, out_path = job.opt.prefix + "_sc.vcf";
    std::ofstream out(out_path.c_str());
    if (!out.is_open()) { std::cerr << "Fail to open output file: " << out_path << "\n"; exit(EXIT_FAILURE); }
    bool wrote_command = false;
    if (path.find("gz") != std::string::npos) {
        std::string text;
        if (!lpsh::read_gz(path, text)) { std::cerr << "Fail to open vcf: " << path << "\n"; return 0; }
        size_t at = 0;
        for (size_t nl; (nl = text.find('\n', at)) != std::string::npos; at = nl + 1) write_sc_line(job, text.substr(at, nl - at), wrote_command, out);
    } else if (path.find("vcf") != std::string::npos) {
        std::ifstream in(path.c_str());
        if (!in.is_open()) { std::cerr << "Fail to open vcf: " << path << "\n"; exit(1); }
        std::string line;
        while (!in.eof()) { std::getline(in, line); write_sc_line(job, line, wrote_command, out); }
    }
    return 0;
}

// SomaticVarCaller::writeSomaticVarCallingLog (SomaticVarCaller.cpp:1576-1926): one row per position called somatic, every column in
// the reference's own type (float / double / int / bool) so that operator<< prints the same characters
int write_somatic_var_out(lpsh_som &job) {
    const SomOptions &o = job.opt;
    const std::string path = o.prefix + "_somatic_var.out";
    std::ofstream f(path.c_str());
    if (!f.is_open()) { std::cerr << "Fail to open write file: " << path << "\n"; exit(1); }
    std::cerr << "writing somatic variants calling log ... ";
    std::time_t begin = time(NULL);
    int total = 0;
    for (size_t c = 0; c < job.chr_names.size(); c++)
        for (uint8_t v : job.contig[c].is_somatic) total += v ? 1 : 0;
    lps_somatic_filter_params fp;
    lps_somatic_filter_params_of(job.purity, &fp);
    f << "####################################\n#   Somatic Variants Calling Log   #\n####################################\n";
    f << "##normalSnpFile:" << o.snp_file << "\n##tumorSnvFile:" << o.tumor_vcf << "\n##bamFile:" << o.bam << "\n##tumorBamFile:" << o.tumor_bam << "\n"
      << "##resultPrefix:" << o.prefix << "\n##numThreads:" << o.threads << "\n##region:" << o.region << "\n##qualityThreshold:" << o.quality << "\n"
      << "##percentageThreshold:" << o.percentage << "\n##tagSupplementary:" << o.tag_supplementary << "\n##\n";
    f << "##======== Filter Parameters =========\n##Enable filter : " << o.enable_filter << "\n##Calling mapping quality :" << o.quality << "\n"
      << "##Tumor purity : " << fp.tumor_purity << "\n##Normal VAF maximum threshold : " << fp.nor_vaf_max << "\n"
      << "##Normal depth minimum threshold : " << fp.nor_depth_min << "\n##Messy read ratio threshold : " << fp.messy_read_ratio << "\n"
      << "##Somatic read count minimum threshold : " << fp.read_count_min << "\n"
      << "##Haplotag consistency filter VAF threshold : " << fp.hap_consistency_vaf_max << "\n"
      << "##Haplotag consistency filter read count threshold : " << fp.hap_consistency_read_count_max << "\n"
      << "##Haplotag consistency somatic read count minimum threshold : " << fp.hap_consistency_somatic_read_min << "\n"
      << "##Interval SNP count filter threshold : " << fp.interval_snp_count_vaf_max << "\n"
      << "##Interval SNP count filter read count threshold : " << fp.interval_snp_count_read_count_max << "\n"
      << "##Interval SNP count minimum threshold : " << fp.interval_snp_count_min << "\n##Z-score maximum threshold : " << fp.z_score_max << "\n"
      << "##DenseAlt filter condition1 threshold : " << fp.dense_alt_condition1 << "\n##DenseAlt filter condition2 threshold : " << fp.dense_alt_condition2 << "\n"
      << "##DenseAlt filter minimum same count threshold : " << fp.dense_alt_same_count_min << "\n##==================================== \n##\n"
      << "##Total Somatic SNPs: " << total << "\n##\n";
    static const char *COLS[] = {"#CHROM", "POS", "ID", "REF", "ALT", "AltCount", "ReadCount", "NorAltCount", "PureH1-1", "PureH2-1", "PureH3", "MixedHpRead", "UnTag",
        "PureH1-1ratio", "PureH2-1ratio", "PureH3ratio", "MixedHpReadRatio", "NorVAF", "TumVAF", "NorMpqVAF", "TumMpqVAF", "NorVAF_substract", "TumVAF_substract",
        "NorDepth", "TumDepth", "Subtract_Depth", "NorDeletionCount", "TumDeletionCount", "NorDeletionRatio", "TumDeletionRatio", "NorMpqReadRatio", "TumMpqReadRatio",
        "ShannonEntropy", "HomopolymerLength", "H1readCount", "H2readCount", "H1_1readCount", "H2_1readCount", "H3readCount", "GermlineReadHpCount",
        "GermlineReadHpImbalanceRatio", "SomaticReadHpImbalanceRatio", "BaseGermlineReadHpImbalanceRatio", "PercentageOfGermlineHp", "H1readCountInNorBam",
        "H2readCountInNorBam", "GermlineReadHpCountInNorBam", "GermlineReadHpImbalanceRatioInNorBam", "PercentageOfGermlineHpInNorBam",
        "GermlineReadHpImbalanceRatioDifference", "PercentageOfGermlineHpDifference", "SomaticRead_H1-1", "SomaticRead_H2-1", "SomaticRead_H3", "SomaticRead_unTag",
        "AltMeanCountPerVarRead", "zScore", "IntervalSnpCount", "IntervalMinDistance", "ExistNorSnp", "StatisticPurity", "isFilterOut", "NorNonDelAF", "TumNonDelAF"};
    for (const char *col : COLS) f << col << "\t";
    f << "GT\n";
    enum { HP_UNTAG_ = 0, HP_H1_ = 1, HP_H2_ = 2, HP_H3_ = 3, HP_H1_1_ = 5, HP_H2_1_ = 7 };   // ReadHP (HaplotagType.h:97-108)
    for (size_t c = 0; c < job.chr_names.size(); c++) {
        const std::string &chr = job.chr_names[c];
        const ContigState &S = job.contig[c];
        const ExtractCopy &N = S.normal, &T = S.tumor;
        const size_t nt = T.tum_var.size();
        if (S.is_somatic.size() != nt) continue;
        size_t k = 0;
        const auto &vars = job.variants[chr];
        for (auto it = vars.begin(); it != vars.end(); ++it) {
            if (!it->second.has_tum) continue;
            const size_t s = k++;
            if (s >= nt || !S.is_somatic[s]) continue;
            const UnionVar &u = it->second;
            if (u.tum.ref.empty() || u.tum.alt.empty()) {
                std::cerr << "[ERROR](write tag HP3 log file) => can't find RefBase or AltBase : chr:" << chr << " pos: " << it->first + 1 << " RefBase:" << u.tum.ref << " AltBase:" << u.tum.alt;
                exit(1);
            }
            const int32_t *tb = &T.pos_base[s * LPS_PB_FIELDS], *nb = &N.pos_base[s * LPS_PB_FIELDS];
            const int32_t *cc = &T.case_count[s * LPS_CASE_FIELDS];
            const float *tf = &T.ratios_f[s * LPS_RF_FIELDS], *nf = &N.ratios_f[s * LPS_RF_FIELDS];
            const double *td = &T.ratios_d[s * LPS_RD_FIELDS], *nd = &N.ratios_d[s * LPS_RD_FIELDS];
            const int32_t *thp = &T.read_hp_count[s * 9], *nhp = &N.read_hp_count[s * 9], *shp = &T.somatic_read_hp_count[s * 9];
            const int norDepth = nb[LPS_PB_DEPTH], tumDepth = tb[LPS_PB_DEPTH], subtractDepth = tumDepth - norDepth;
            const float tumVAF = tf[LPS_RF_VAF], tumMpqVAF = tf[LPS_RF_MPQ_VAF], norVAF = nf[LPS_RF_VAF], norMpqVAF = nf[LPS_RF_MPQ_VAF];
            const float norVAF_substract = (norMpqVAF - norVAF), tumVAF_substract = (tumMpqVAF - tumVAF);
            const int germlineReadHpCount = thp[HP_H1_] + thp[HP_H2_], germlineReadHpCountInNorBam = nhp[HP_H1_] + nhp[HP_H2_];
            const double germlineReadHpImbalanceRatioDifference = td[LPS_RD_GERMLINE_IMBALANCE] - nd[LPS_RD_GERMLINE_IMBALANCE];
            const double percentageOfGermlineHpDifference = td[LPS_RD_PCT_GERMLINE_HP] - nd[LPS_RD_PCT_GERMLINE_HP];
            double zScore = -1.0;
            if (S.in_dense[s]) {
                if (S.z_score[s] < 0.0) { std::cerr << "[ERROR](zScore) => chr: " << chr << " pos: " << it->first + 1 << " zScore: " << S.z_score[s] << std::endl; exit(1); }
                zScore = S.z_score[s];
            }
            const char *gt = u.tum.gt_kind == 3 ? "Homo" : u.tum.gt_kind == 1 ? "Hetero" : u.tum.gt_kind == 2 ? "UnphasedHetero" : "";
            const double shannonEntropy = 0.0;                       // SomaticData members the reference never assigns
            const int homopolymerLength = 0;
            f << chr << " \t" << it->first + 1 << "\t" << "." << "\t" << u.tum.ref << "\t" << u.tum.alt << "\t" << tb[LPS_PB_ALT] << "\t" << T.case_read_count[s] << "\t\t"
              << nb[LPS_PB_ALT] << "\t" << cc[LPS_CASE_PURE_H1_1] << "\t" << cc[LPS_CASE_PURE_H2_1] << "\t" << cc[LPS_CASE_PURE_H3] << "\t" << cc[LPS_CASE_MIXED] << "\t"
              << cc[LPS_CASE_UNTAG] << "\t\t" << tf[LPS_RF_PURE_H1_1_RATIO] << "\t" << tf[LPS_RF_PURE_H2_1_RATIO] << "\t" << tf[LPS_RF_PURE_H3_RATIO] << "\t"
              << tf[LPS_RF_MIXED_RATIO] << "\t\t" << norVAF << "\t" << tumVAF << "\t\t" << norMpqVAF << "\t" << tumMpqVAF << "\t\t" << norVAF_substract << "\t"
              << tumVAF_substract << "\t\t" << norDepth << "\t" << tumDepth << "\t" << subtractDepth << "\t" << nb[LPS_PB_DEL] << "\t" << tb[LPS_PB_DEL] << "\t"
              << nf[LPS_RF_DEL_RATIO] << "\t" << tf[LPS_RF_DEL_RATIO] << "\t" << nf[LPS_RF_LOW_MPQ_RATIO] << "\t" << tf[LPS_RF_LOW_MPQ_RATIO] << "\t"
              << shannonEntropy << "\t" << homopolymerLength << "\t\t" << thp[HP_H1_] << "\t" << thp[HP_H2_] << "\t" << thp[HP_H1_1_] << "\t" << thp[HP_H2_1_] << "\t"
              << thp[HP_H3_] << "\t" << germlineReadHpCount << "\t" << td[LPS_RD_GERMLINE_IMBALANCE] << "\t" << td[LPS_RD_SOMATIC_IMBALANCE] << "\t"
              << td[LPS_RD_ALLELIC_IMBALANCE] << "\t" << td[LPS_RD_PCT_GERMLINE_HP] << "\t" << nhp[HP_H1_] << "\t" << nhp[HP_H2_] << "\t" << germlineReadHpCountInNorBam << "\t"
              << nd[LPS_RD_GERMLINE_IMBALANCE] << "\t" << nd[LPS_RD_PCT_GERMLINE_HP] << "\t" << germlineReadHpImbalanceRatioDifference << "\t"
              << percentageOfGermlineHpDifference << "\t" << shp[HP_H1_1_] << "\t" << shp[HP_H2_1_] << "\t" << shp[HP_H3_] << "\t" << shp[HP_UNTAG_] << "\t"
              << S.mean_alt[s] << "\t" << zScore << "\t" << S.interval_snp_count[s] << "\t" << S.min_distance[s] << "\t" << u.has_nor << "\t"
              << (bool)(S.used_for_purity.size() == nt && S.used_for_purity[s]) << "\t" << (bool)S.is_filter_out[s] << "\t" << nf[LPS_RF_NONDEL_VAF] << "\t"
              << tf[LPS_RF_NONDEL_VAF] << "\t" << gt << "\n";
        }
    }
    f.close();
    std::cerr << difftime(time(NULL), begin) << "s\n";
    return 0;
}

std::string contig_region(const lpsh_som &job, const std::string &chr) {
    return !job.opt.region.empty() ? job.opt.region : chr + ":1-" + std::to_string(job.chr_length.at(chr));
}

void finish_contig(lpsh_som &job) {
    job.io.end_region();
    job.chunk_contig = -1;
    job.chunk.clear();
}

}  // namespace

extern "C" {

int lpsh_som_open(int argc, char **argv, lpsh_som **out) {
    if (!out) return -1;
    *out = nullptr;
    lpsh_som *job = new lpsh_som();
    const int rc = parse_som_options(argc, argv, job->opt);
    if (rc != 0) { delete job; return rc; }
    if (const char *e = getenv("LPS_TAG_CHUNK")) { const long v = atol(e); if (v > 0) job->chunk_reads = (size_t)v; }
    som_banner(job->opt);
    if (load_union(*job) != 0 || load_som_reference(*job) != 0) { delete job; return -1; }
    job->contig.resize(job->chr_names.size());
    memset(&job->purity_result, 0, sizeof(job->purity_result));
    *out = job;
    return 0;
}

int lpsh_som_n_contigs(const lpsh_som *h) { return h ? (int)h->chr_names.size() : 0; }
const char *lpsh_som_contig_name(const lpsh_som *h, int i) {
    return (h && i >= 0 && (size_t)i < h->chr_names.size()) ? h->chr_names[(size_t)i].c_str() : nullptr;
}
// pass 0 = the two extract passes (ParsingBamControl defaults: no mapping-quality filter), pass 1 = the tagging pass
int lpsh_som_params(const lpsh_som *h, int pass, lps_tag_params *out) {
    if (!h || !out) return -1;
    memset(out, 0, sizeof(*out));
    out->mapping_quality = h->opt.quality;
    out->mapq_filter = pass ? 1 : 0;
    out->tag_supplementary = h->opt.tag_supplementary;
    out->have_reference = 1;
    out->percentage_threshold = h->opt.percentage;
    return 0;
}

// every alignment of contig i of the NORMAL (which = 0) or TUMOR (which = 1) BAM with the contig's union map
int lpsh_som_pack(lpsh_som *h, int i, int which, lpsh_packed *out, lps_tumor_variants *tv) {
    if (!h || !out || !tv || i < 0 || (size_t)i >= h->chr_names.size()) return -1;
    const std::string &chr = h->chr_names[(size_t)i];
    const std::string &path = which ? h->opt.tumor_bam : h->opt.bam;
    h->pack = lpsh::PackedContig();
    pack_union(*h, chr, h->pack, h->tum);
    h->pack.ref = (which ? h->ref_tumor : h->ref_normal)[chr];
    samFile *in = hts_open(path.c_str(), "r");
    if (!in) return lpsh::fail("Cannot open bam file " + path);
    hts_set_fai_filename(in, h->opt.fasta.c_str());
    bam_hdr_t *hdr = sam_hdr_read(in);
    hts_idx_t *idx = hdr ? sam_index_load(in, path.c_str()) : NULL;
    if (!idx) { if (hdr) bam_hdr_destroy(hdr); sam_close(in); return lpsh::fail("Cannot open index for bam file " + path); }
    htsThreadPool pool = {NULL, 0};
    if (h->opt.threads > 1 && (pool.pool = hts_tpool_init(h->opt.threads))) hts_set_opt(in, HTS_OPT_THREAD_POOL, &pool);
    hts_itr_t *it = sam_itr_querys(idx, hdr, contig_region(*h, chr).c_str());
    bam1_t *aln = bam_init1();
    int inflate_rc = 0;
    if (it) {
        int done = 0;
        if (lpsh::gpu_inflate_requested()) {   // the region inflated in one batch on the device, records packed from memory
            done = lpsh::pack_region_inflated(path, it, h->pack);
            if (done < 0) inflate_rc = done;
            if (done == 0) h->pack.truncate_reads(0);
        }
        if (done == 0) {   // the contigs of an extract pass come one after the other: all -t threads read slices of this one (LPS_READ_SPLIT overrides)
            int readers = h->opt.threads >= 4 ? h->opt.threads : 1;
            if (const char *e = getenv("LPS_READ_SPLIT")) { const int v = atoi(e); if (v >= 1) readers = v; }
            if (readers > 1 && !it->multi) {
                done = lpsh::pack_region_split(path, h->opt.fasta, idx, it->tid, it->beg, it->end, readers, h->pack);
                if (done < 0) inflate_rc = done;
            }
        }
        if (done == 0) while (sam_itr_multi_next(in, it, aln) >= 0) h->pack.add_alignment(aln);
        hts_itr_destroy(it);
    }
    bam_destroy1(aln);
    hts_idx_destroy(idx);
    bam_hdr_destroy(hdr);
    sam_close(in);
    if (pool.pool) hts_tpool_destroy(pool.pool);
    if (inflate_rc != 0) return inflate_rc;
    h->pack.finish();
    h->pack.view(out);
    h->tum.view(tv);
    return 0;
}

int lpsh_som_set_extract(lpsh_som *h, int i, int which, const lps_extract_result *r) {
    if (!h || !r || i < 0 || (size_t)i >= h->contig.size()) return -1;
    (which ? h->contig[(size_t)i].tumor : h->contig[(size_t)i].normal).assign(*r);
    return 0;
}

// SomaticVarCaller::runTumorPurityEstimator (+ <prefix>_purity.out), or --tumor-purity (SomaticVarCaller.cpp:826-832, 937-949)
int lpsh_som_estimate(lpsh_som *h) {
    if (!h) return -1;
    const SomOptions &o = h->opt;
    const size_t nc = h->chr_names.size();
    for (size_t c = 0; c < nc; c++)
        if (!h->contig[c].normal.set || !h->contig[c].tumor.set) return lpsh::fail("extract results of contig " + h->chr_names[c] + " are missing");
    // positions with a SomaticData entry = tumor slots a tumor alignment reached (the keys of chrPosSomaticInfo)
    std::vector<std::vector<uint8_t>> touched(nc);
    for (size_t c = 0; c < nc; c++) {
        const ExtractCopy &T = h->contig[c].tumor;
        const size_t nt = T.tum_var.size();
        touched[c].assign(nt, 0);
        for (size_t k = 0; k < nt; k++) {
            int64_t hp_sum = 0;
            for (int j = 0; j < 9; j++) hp_sum += T.read_hp_count[k * 9 + (size_t)j] + (T.somatic_read_hp_count.empty() ? 0 : T.somatic_read_hp_count[k * 9 + (size_t)j]);
            touched[c][k] = T.pos_base[k * LPS_PB_FIELDS + LPS_PB_DEPTH] > 0 || hp_sum > 0;   // same test as lps_somatic_call's `touched`
        }
    }
    if (o.estimate_purity) {
        std::time_t t0 = time(NULL);
        std::cerr << "estimating tumor purity ... ";
        std::vector<double> t_imb, n_imb, n_pct;
        std::vector<int32_t> n_h1, n_h2;
        std::vector<std::pair<size_t, size_t>> where;     // (contig, tumor slot) of every entry handed to the estimator
        for (size_t c = 0; c < nc; c++) {
            const ExtractCopy &N = h->contig[c].normal, &T = h->contig[c].tumor;
            if (N.tum_var.size() != T.tum_var.size()) return lpsh::fail("normal and tumor extract results of " + h->chr_names[c] + " differ in size");
            for (size_t k = 0; k < T.tum_var.size(); k++) {
                if (!touched[c][k]) continue;
                t_imb.push_back(T.ratios_d[k * LPS_RD_FIELDS + LPS_RD_GERMLINE_IMBALANCE]);
                n_imb.push_back(N.ratios_d[k * LPS_RD_FIELDS + LPS_RD_GERMLINE_IMBALANCE]);
                n_pct.push_back(N.ratios_d[k * LPS_RD_FIELDS + LPS_RD_PCT_GERMLINE_HP]);
                n_h1.push_back(N.read_hp_count[k * 9 + 1]);
                n_h2.push_back(N.read_hp_count[k * 9 + 2]);
                where.push_back(std::make_pair(c, k));
            }
        }
        lps_purity_input in;
        memset(&in, 0, sizeof(in));
        in.n = (int32_t)t_imb.size();
        in.tumor_germline_imbalance = t_imb.data(); in.normal_germline_imbalance = n_imb.data(); in.normal_pct_germline_hp = n_pct.data();
        in.normal_h1 = n_h1.data(); in.normal_h2 = n_h2.data();
        std::vector<uint8_t> used(t_imb.size() + 1, 0);
        in.used = used.data();
        lps_purity_result &r = h->purity_result;
        if (lps_estimate_purity(&in, &r) != 0) return lpsh::fail("lps_estimate_purity failed");
        for (size_t c = 0; c < nc; c++) h->contig[c].used_for_purity.assign(h->contig[c].tumor.tum_var.size(), 0);
        for (size_t e = 0; e < where.size(); e++) h->contig[where[e].first].used_for_purity[where[e].second] = used[e];   // markStatisticFlag
        if (r.ok) {
            std::cerr << difftime(time(NULL), t0) << "s\n";
            std::ofstream f((o.prefix + "_purity.out").c_str());   // TumorPurityEstimator::writePurityResult (:375-424)
            if (!f.is_open()) std::cerr << "[ERROR] :Failed to open purity log file: " << o.prefix << "_purity.out\n[ERROR] : Failed to write purity log" << std::endl;
            else {
                f << "#==================================\n# TUMOR PURITY ESTIMATION REPORT\n#==================================\n";
                f << "#Initial data size: " << in.n << std::endl;
                f << "#==========filter parameters==========" << std::endl;
                f << "#GERMLINE_HP_IMBALANCE_RATIO_MIN_THR: " << 0.0f << std::endl << "#GERMLINE_HP_IMBALANCE_RATIO_IN_NOR_BAM_MIN_THR: " << 0.0f << std::endl;
                f << "#GERMLINE_HP_IMBALANCE_RATIO_IN_NOR_BAM_MAX_THR: " << 0.7f << std::endl << "#GERMLINE_HP_PERCENTAGE_IN_NOR_BAM_MAX_THR: " << 0.7f << std::endl;
                f << "#GERMLINE_HP_READ_COUNT_IN_NOR_BAM_MIN_THR: " << 5 << std::endl << "#GERMLINE_HP_READ_COUNT_IN_NOR_BAM_DYNAMIC_THR: " << r.read_count_threshold << std::endl;
                f << "#==========Initial filter out data count==========" << std::endl;
                f << "#imbalanceRatioInNorBam: " << r.filtered_normal_imbalance_zero << std::endl << "#imbalanceRatio: " << r.filtered_tumor_imbalance_zero << std::endl;
                f << "#imbalanceRatioInNorBam_over_thr: " << r.filtered_normal_imbalance_high << std::endl << "#readHpCountInNorBam: " << r.filtered_normal_read_count << std::endl;
                f << "#percentageOfGermlineHpInNorBam: " << r.filtered_pct_germline_hp << std::endl;
                f << "#==========Second filter out data count==========" << std::endl << "#peakValley count: " << r.filtered_valley << std::endl;
                f << "#==========Whisker filter out data count==========" << std::endl << "#iteration times: " << 1 << std::endl;
                f << "#remove outliers: " << r.filtered_outliers << std::endl << "#==========Statistical analysis===========" << std::endl;
                f << "Data size: " << r.n_used << std::endl << "Median: " << r.median << std::endl << "Q1: " << r.q1 << std::endl << "Q3: " << r.q3 << std::endl;
                f << "IQR: " << r.iqr << std::endl << "Whiskers: " << r.lower_whisker << " to " << r.upper_whisker << std::endl;
                f << "Outliers: " << r.n_outliers_left << std::endl << "#==========Estimation result===========" << std::endl;
                f << "Tumor purity: " << r.purity << std::endl;
            }
        } else {
            std::cerr << "[ERROR] Failed to estimate tumor purity, set purity to 0.0" << std::endl;
        }
        h->purity = r.purity;
    } else {
        h->purity = o.purity;
    }
    return 0;
}

// the purity stage, then the calling stage and getSomaticFlag for every contig (SomaticVarCaller.cpp:833-866, 2397-2412)
int lpsh_som_call(lpsh_som *h) {
    if (!h) return -1;
    if (lpsh_som_estimate(h) != 0) return -1;
    const SomOptions &o = h->opt;
    const size_t nc = h->chr_names.size();
    std::time_t t0 = time(NULL);
    std::cerr << "calling somatic variants ... ";
    int failed = 0;
    h->n_somatic = 0;
    for (const std::string &chr : h->chr_names) h->variants[chr];   // operator[] inserts: give every contig its slot before the threads start
#pragma omp parallel for schedule(dynamic) num_threads(o.threads)
    for (int c = 0; c < (int)nc; c++) {
        const std::string &chr = h->chr_names[(size_t)c];
        ExtractCopy &N = h->contig[(size_t)c].normal, &T = h->contig[(size_t)c].tumor;
        const size_t nt = T.tum_var.size();
        std::vector<int32_t> pos(nt);
        std::vector<uint8_t> callable(nt), is_somatic(nt, 0);
        std::vector<int8_t> derive(nt, 0);
        std::vector<std::map<int, UnionVar>::iterator> slot(nt);
        auto &vars = h->variants[chr];
        {
            size_t k = 0;
            for (auto it = vars.begin(); it != vars.end(); ++it)
                if (it->second.has_tum && k < nt) {
                    pos[k] = it->first;
                    callable[k] = is_snp(it->second.tum) || is_ins(it->second.tum) || is_del(it->second.tum);
                    slot[k++] = it;
                }
            if (k != nt) {
#pragma omp critical
                { lpsh::fail("tumor slots of " + chr + " do not match the union map"); failed = 1; }
                continue;
            }
        }
        const lps_extract_result rn = N.view(), rt = T.view();
        lps_somatic_call_input in;
        memset(&in, 0, sizeof(in));
        in.n_tum = (int32_t)nt; in.pos = pos.data(); in.callable = callable.data(); in.normal = &rn; in.tumor = &rt;
        in.purity = h->purity; in.enable_filter = o.enable_filter; in.percentage_threshold = o.percentage;
        lps_somatic_call_result out;
        memset(&out, 0, sizeof(out));
        out.is_somatic = is_somatic.data(); out.derive_hp = derive.data();
        ContigState &CS = h->contig[(size_t)c];
        if (o.write_calling_log) {
            CS.is_filter_out.assign(nt, 0); CS.in_dense.assign(nt, 0); CS.mean_alt.assign(nt, 0.f); CS.z_score.assign(nt, 0.f);
            CS.interval_snp_count.assign(nt, 0); CS.min_distance.assign(nt, 0);
            out.is_filter_out = CS.is_filter_out.data(); out.in_dense_interval = CS.in_dense.data(); out.mean_alt_per_var_read = CS.mean_alt.data();
            out.z_score = CS.z_score.data(); out.interval_snp_count = CS.interval_snp_count.data(); out.min_distance = CS.min_distance.data();
        }
        const int rc = nt ? lps_somatic_call(&in, &out) : 0;
        if (rc != 0) {
#pragma omp critical
            { lpsh::fail("somatic calling failed on " + chr + (rc == LPS_E_DATA ? ": a tumor position without any read record (the reference exits here)" : "")); failed = 1; }
            continue;
        }
        for (size_t k = 0; k < nt; k++) { slot[k]->second.is_somatic = is_somatic[k]; slot[k]->second.derive_hp = derive[k]; }   // getSomaticFlag
        CS.is_somatic = is_somatic;
#pragma omp critical
        h->n_somatic += nt ? out.n_somatic : 0;
    }
    std::cerr << difftime(time(NULL), t0) << "s\n";
    if (!failed && o.write_calling_log) write_somatic_var_out(*h);   // SomaticVarCaller.cpp:877-879
    if (!failed && o.write_sc_vcf) {      // SomaticHaplotagProcess.cpp:77-85
        std::time_t w0 = time(NULL);
        std::cerr << "writing somatic variants to vcf file ... ";
        write_sc_vcf(*h);
        std::cerr << difftime(time(NULL), w0) << "s\n";
    }
    return failed ? -1 : 0;
}

double lpsh_som_purity(const lpsh_som *h) { return h ? h->purity : 0.0; }
int64_t lpsh_som_n_somatic(const lpsh_som *h) { return h ? h->n_somatic : 0; }

int lpsh_som_tag_begin(lpsh_som *h) {
    if (!h) return -1;
    const SomOptions &o = h->opt;
    return h->io.open(o.tumor_bam, o.fasta, o.prefix + (o.cram ? ".cram" : ".bam"), o.cram ? "wc" : lpsh::bam_write_mode(), o.threads, o.command);
}

// next chunk of contig i of the tumor BAM into `ck` (reads + the NORMAL side of the union map): 1 = filled, 0 = exhausted, < 0 error
static int read_chunk(lpsh_som *h, int i, lpsh::Chunk &ck) {
    const std::string &chr = h->chr_names[(size_t)i];
    if (h->io.cur != i) {
        const int rc = h->io.start_region(i, contig_region(*h, chr));
        if (rc < 0) return rc;
    }
    ck.clear();
    lpsh::PackedContig &pc = ck.pack;
    TumorArrays unused;
    pack_union(*h, chr, pc, unused);
    pc.ref_shared = &h->ref_tumor[chr];
    const int got = h->io.fill(ck, h->chunk_reads);
    if (got <= 0) { ck.clear(); return got; }
    pc.finish();
    return 1;
}

// staged API: the chunk plus the TUMOR side of the union map, now carrying the caller's flags
int lpsh_som_tag_pack(lpsh_som *h, int i, lpsh_packed *out, lps_tumor_variants *tv) {
    if (!h || !out || !tv || i < 0 || (size_t)i >= h->chr_names.size() || !h->io.in) return -1;
    const int got = read_chunk(h, i, h->chunk);
    if (got != 1) return got;
    lpsh::PackedContig scratch;
    pack_union(*h, h->chr_names[(size_t)i], scratch, h->tum);
    h->chunk.pack.view(out);
    h->tum.view(tv);
    h->chunk_contig = i;
    return 1;
}

static int emit_chunk(lpsh_som *h, lpsh::Chunk &ck, const lps_somatic_tag_result *r);

int lpsh_som_tag_emit(lpsh_som *h, int i, const lps_somatic_tag_result *r) {
    if (!h || !r || h->chunk_contig != i || !h->io.has_output()) return -1;
    return emit_chunk(h, h->chunk, r);
}

// SomaticHaplotagChrProcessor: processRead's tag handling + addAuxiliaryTags (HaplotagProcess.cpp:318-355, SomaticHaplotagProcess.cpp:464-472)
static int emit_chunk(lpsh_som *h, lpsh::Chunk &ck, const lps_somatic_tag_result *r) {
    const size_t n = ck.records.size();
    if ((size_t)r->reads.n_reads != n) return lpsh::fail("verdict count does not match the chunk");
    for (size_t k = 0; k < n; k++) {
        bam1_t *b = ck.records[k];
        if (r->reads.category[k] == LPS_TAG_PROCESSED) {
            lpsh::drop_aux(b, "HP");
            lpsh::drop_aux(b, "PS");
            lpsh::drop_aux(b, "PQ");
            const int hp = r->reads.read_hp[k];
            if (hp != 0) {
                if (hp < 0 || hp >= LPS_READHP_FIELDS) return lpsh::fail("read haplotype out of range");
                const char *text = READ_HP_TEXT[hp];
                int ps = r->reads.ps[k], pq = r->reads.pq[k];
                bam_aux_append(b, "HP", 'Z', (int)strlen(text) + 1, (const uint8_t *)text);
                if (ps != -1) bam_aux_append(b, "PS", 'i', sizeof(int), (uint8_t *)&ps);
                bam_aux_append(b, "PQ", 'i', sizeof(int), (uint8_t *)&pq);
            }
        }
        if (h->io.write(b) < 0) { std::cerr << "[ERROR](BamFileRAII): write output bam file failed" << std::endl; return lpsh::fail("write output bam file failed"); }
    }
    // ReadStatistics arrive reduced over the chunk
    h->st_alignment += r->total_alignment; h->st_supplementary += r->total_supplementary; h->st_secondary += r->total_secondary;
    h->st_unmapped += r->total_unmapped; h->st_tag += r->total_tag; h->st_untag += r->total_untag; h->st_low += r->total_lower_quality;
    h->st_other += r->total_other_case; h->st_empty += r->total_empty_variant; h->st_similar += r->total_high_similarity;
    h->st_cross += r->total_cross_two_block; h->st_no_variant += r->total_without_variant; h->st_only_h3 += r->total_read_only_h3;
    for (int k = 0; k < LPS_READHP_FIELDS; k++) h->st_hp[k] += r->total_hp[k];
    ck.clear();
    return 0;
}

int lpsh_som_tag_end(lpsh_som *h) {
    if (!h) return -1;
    finish_contig(*h);
    const int rc = h->io.close();
    std::ostream &e = std::cerr;   // HaplotagProcess::printExecutionReport (HaplotagProcess.cpp:152-175)
    e << "-------------------------------------------\n";
    e << "total process time        : " << difftime(time(NULL), h->t_begin) << "s\n";
    e << "total alignment           : " << h->st_alignment << "\ntotal supplementary       : " << h->st_supplementary << "\n";
    e << "total secondary           : " << h->st_secondary << "\ntotal unmapped            : " << h->st_unmapped << "\n";
    e << "total tagged alignments   : " << h->st_tag << "\n    L----total HP1        : " << h->st_hp[1] << "\n    L----total HP2        : " << h->st_hp[2] << "\n";
    e << "    L----total HP1-1      : " << h->st_hp[5] << "\n    L----total HP2-1      : " << h->st_hp[7] << "\n    L----total HP3        : " << h->st_hp[3] << "\n";
    e << "         L----only H3 SNP : " << h->st_only_h3 << "\n";
    e << "total untagged            : " << h->st_untag << "\n    L----lower mapping quality        : " << h->st_low << "\n";
    e << "    L----no variant                   : " << h->st_empty << "\n    L----start pos > last variant pos : " << h->st_other << "\n";
    e << "    L----judge to untag               : " << h->st_hp[0] << "\n         L----high similarity         : " << h->st_similar << "\n";
    e << "         L----cross two block         : " << h->st_cross << "\n         L----no variant judge HP     : " << h->st_no_variant << "\n";
    e << "-------------------------------------------\n";
    return rc;
}

// The tagging pass with any judge: reader thread (htslib parsing + packing) | calling thread (judge, HP:Z / PS / PQ, sam_write1).
// Needs lpsh_som_call (the flags travel in the TUMOR side of the union map).  begin ... end included.
int lpsh_som_tag_run_with(lpsh_som *h, lpsh_som_judge_fn judge, void *user) {
    if (!h || !judge) return -1;
    if (lpsh_som_tag_begin(h) != 0) return -1;
    lps_tag_params tp;
    lpsh_som_params(h, 1, &tp);
    std::time_t t0 = time(NULL);
    std::cerr << "somatic tagging start ...\n";
    int tum_contig = -1;
    auto handle = [&](int i, lpsh::Chunk &ck) -> int {
        if (tum_contig != i) {   // TUMOR side of the union map: once per contig, on this thread
            lpsh::PackedContig scratch;
            pack_union(*h, h->chr_names[(size_t)i], scratch, h->tum);
            tum_contig = i;
        }
        lpsh_packed v;
        lps_tumor_variants tv;
        ck.pack.view(&v);
        h->tum.view(&tv);
        lps_somatic_tag_result r;
        memset(&r, 0, sizeof(r));
        std::vector<uint8_t> cat;
        std::vector<int8_t> hp;
        std::vector<int32_t> zero;
        if (v.variants.n == 0) {      // dispatch without variants: MAPQ and flags only (HaplotagParsingBam.cpp:457-476)
            const int n = v.batch.n_reads;
            cat.resize((size_t)n); hp.assign((size_t)n, 0); zero.assign((size_t)n, 0);
            for (int k = 0; k < n; k++) {
                const int flag = v.batch.flag[k];
                cat[(size_t)k] = lpsh::category_without_variants(v.batch.mapq[k], flag, tp);
                r.total_alignment++; r.total_untag++;
                if (cat[(size_t)k] == LPS_TAG_LOW_MAPQ) r.total_lower_quality++;
                else if (cat[(size_t)k] == LPS_TAG_UNMAPPED) r.total_unmapped++;
                else if (cat[(size_t)k] == LPS_TAG_SECONDARY) r.total_secondary++;
                else if (cat[(size_t)k] == LPS_TAG_SUPPLEMENTARY) r.total_supplementary++;
                else r.total_empty_variant++;
            }
            r.reads.n_reads = n; r.reads.category = cat.data(); r.reads.read_hp = hp.data(); r.reads.ps = zero.data(); r.reads.pq = zero.data();
        } else if (judge(user, i, &v, &tv, &r) != 0) {
            return lpsh::fail(std::string("contig ") + h->chr_names[(size_t)i] + ": the judge failed");
        }
        return emit_chunk(h, ck, &r);
    };
    const int rc = lpsh::run_chunk_pipeline((int)h->chr_names.size(), [&](int i, lpsh::Chunk &ck) { return read_chunk(h, i, ck); }, handle);
    std::cerr << "tag read " << difftime(time(NULL), t0) << "s\n";
    const int rc_end = lpsh_som_tag_end(h);
    return rc != 0 ? -1 : rc_end;
}

namespace {
struct SomDeviceJudge {
    lps_ctx *ctx = nullptr;
    lps_tag_params tp;
    int contig = -1;
    std::string error;
};
int som_device_judge(void *user, int contig, const lpsh_packed *v, const lps_tumor_variants *tv, lps_somatic_tag_result *out) {
    SomDeviceJudge *d = (SomDeviceJudge *)user;
    int rc = 0;
    if (d->contig != contig) {
        rc = lps_contig_set_reference(d->ctx, v->ref, v->ref_len);
        if (rc == 0) rc = lps_contig_set_variants(d->ctx, &v->variants, 0);
        if (rc == 0) rc = lps_contig_set_tumor_variants(d->ctx, tv);
        d->contig = contig;
    }
    if (rc == 0) rc = lps_batch_submit(d->ctx, &v->batch);
    if (rc == 0) rc = lps_somatic_tag_reads(d->ctx, &d->tp, 0, out);
    if (rc != 0) d->error = lps_last_error(d->ctx);
    return rc;
}
}  // namespace

int lpsh_som_run(lpsh_som *h) {
    if (!h) return -1;
    lps_ctx *ctx = nullptr;   // created after the first contig is packed: the driver starts (lpsh_som_main) while the BAM is decoded
    lps_tag_params xp, tp;
    lpsh_som_params(h, 0, &xp);
    lpsh_som_params(h, 1, &tp);
    int rc = 0;
    const int nc = (int)h->chr_names.size();
    // SomaticVarCaller::extractSomaticData: the NORMAL BAM, then the TUMOR BAM (SomaticVarCaller.cpp:907-935)
    for (int which = 0; which < 2 && rc == 0; which++) {
        std::time_t t0 = time(NULL);
        std::cerr << (which ? "extracting data from tumor BAM ... " : "extracting data from normal BAM ... ");
        for (int i = 0; i < nc && rc == 0; i++) {
            lpsh_packed v;
            lps_tumor_variants tv;
            if (lpsh_som_pack(h, i, which, &v, &tv) != 0) { rc = -1; break; }
            if (!ctx && lps_ctx_create(0, &ctx) != 0) { lpsh::fail("no usable CUDA device (there is no CPU fallback)"); rc = -1; break; }
            lps_extract_result r;
            if (v.variants.n == 0) {      // a contig without any variant: nothing reaches the parsers (processEmptyVariants)
                memset(&r, 0, sizeof(r));
            } else {
                rc = lps_contig_set_reference(ctx, v.ref, v.ref_len);
                if (rc == 0) rc = lps_contig_set_variants(ctx, &v.variants, 0);
                if (rc == 0) rc = lps_contig_set_tumor_variants(ctx, &tv);
                if (rc == 0) rc = lps_batch_submit(ctx, &v.batch);
                if (rc == 0) rc = which ? lps_extract_tumor(ctx, &xp, &r) : lps_extract_normal(ctx, &xp, &r);
                if (rc != 0) { lpsh::fail(std::string("contig ") + h->chr_names[(size_t)i] + ": " + lps_last_error(ctx)); break; }
            }
            lpsh_som_set_extract(h, i, which, &r);
        }
        std::cerr << difftime(time(NULL), t0) << "s\n";
    }
    h->pack = lpsh::PackedContig();
    if (h->opt.purity_only) {             // PurityEstimProcess::estimatePurity + printExecutionReport (PurityEstimationProcess.cpp:44-77)
        if (rc == 0) rc = lpsh_som_estimate(h);
        if (ctx) lps_ctx_destroy(ctx);
        std::cerr << "-------------------------------------------\ntotal process time:    " << difftime(time(NULL), h->t_begin) << "s\n"
                  << "estimated tumor purity: " << h->purity << "\n-------------------------------------------\n";
        return rc != 0 ? -1 : 0;
    }
    if (rc == 0) rc = lpsh_som_call(h);
    SomDeviceJudge d;
    d.ctx = ctx;
    d.tp = tp;
    if (rc == 0 && !ctx) { lpsh::fail("no contig to process"); rc = -1; }
    h->io.device_pass = true;   // the tagged BAM is deflated on the device as well (LPS_GPU_DEFLATE=0: htslib's writer)
    if (rc == 0) rc = lpsh_som_tag_run_with(h, som_device_judge, &d);
    if (rc != 0 && !d.error.empty()) lpsh::fail(d.error);
    if (ctx) lps_ctx_destroy(ctx);
    return rc != 0 ? -1 : 0;
}

void lpsh_som_close(lpsh_som *h) {
    if (!h) return;
    if (h->io.is_open()) lpsh_som_tag_end(h);
    delete h;
}

int lpsh_som_main(int argc, char **argv) {
    std::thread warm = lpsh::warm_up_device();   // the driver starts while the VCFs, the FASTA and the first BAM regions are read
    struct Join { std::thread &t; ~Join() { if (t.joinable()) t.join(); } } join_warm{warm};
    lpsh_som *job = nullptr;
    const int rc = lpsh_som_open(argc, argv, &job);
    if (rc == 2) return 0;
    const char *prog = argc > 0 ? argv[0] : "somatic_haplotag";
    if (rc != 0) { if (rc < 0) std::cerr << "[ERROR] " << prog << ": " << lpsh_last_error() << "\n"; return 1; }
    const int run = lpsh_som_run(job);
    if (run != 0) std::cerr << "[ERROR] " << prog << ": " << lpsh_last_error() << "\n";
    lpsh_som_close(job);
    return run != 0 ? 1 : 0;
}

}  // extern "C"
