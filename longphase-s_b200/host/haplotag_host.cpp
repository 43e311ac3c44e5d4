// haplotag_host.cpp — the `haplotag` sub-command above the C ABI: options, the phased-VCF loader, the ordered BAM
// reader / writer with its @PG line, HP / PS / PQ tagging from the device verdicts, the --log table and the report.
// See lps_host.h for the reference seams each stage replaces.
//
// The reference tags one record at a time inside its htslib loop (HaplotagParsingBam.cpp:453-492).  Here the loop appends
// records to a chunk (LPS_TAG_CHUNK alignments, default 65536), one lps_tag_reads call judges the chunk on the device, and
// the records are tagged and written in their original order, so the output file is the same byte stream.
// Scope notes: --sv-file and --mod-file are parsed but rejected (outside the rebuilt hot path).
#include "host_common.h"

#include <getopt.h>

#include <cmath>
#include <ctime>
#include <fstream>
#include <iterator>
#include <sstream>

namespace {

const char *TAG_USAGE =
    "Usage:  haplotag [OPTION] ... READSFILE\n"
    "      --help                          display this help and exit.\n\n"
    "required arguments:\n"
    "      -s, --snp-file=NAME             input SNP vcf file (phased).\n"
    "      -b, --bam-file=NAME             input bam file.\n"
    "      -r, --reference=NAME            reference fasta.\n"
    "optional arguments:\n"
    "      --tagSupplementary              tag supplementary alignment. default:false\n"
    "      -q, --qualityThreshold=Num      not tag alignment if the mapping quality less than threshold. default:1\n"
    "      -p, --percentageThreshold=Num   share of the alleles the winning haplotype needs. default:0.6\n"
    "      -t, --threads=Num               number of BAM (de)compression threads. default:1\n"
    "      -o, --out-prefix=NAME           prefix of the tagged BAM. default:result\n"
    "      --region=REGION                 chrom | chrom:start | chrom:start-end. default:\"\"(all regions)\n"
    "      --log                           an additional log file records the result of each read. default:false\n"
    "      --cram                          the output file will be in the cram format. default:bam\n"
    "not available in this build: --sv-file, --mod-file\n";

enum { T_HELP = 1, T_SUP, T_SV, T_MOD, T_REGION, T_CRAM, T_LOG };

const struct option TAG_LONG[] = {
    {"help", no_argument, NULL, T_HELP},
    {"snp-file", required_argument, NULL, 's'},
    {"bam-file", required_argument, NULL, 'b'},
    {"reference", required_argument, NULL, 'r'},
    {"sv-file", required_argument, NULL, T_SV},
    {"mod-file", required_argument, NULL, T_MOD},
    {"threads", required_argument, NULL, 't'},
    {"qualityThreshold", required_argument, NULL, 'q'},
    {"percentageThreshold", required_argument, NULL, 'p'},
    {"tagSupplementary", no_argument, NULL, T_SUP},
    {"out-prefix", required_argument, NULL, 'o'},
    {"region", required_argument, NULL, T_REGION},
    {"cram", no_argument, NULL, T_CRAM},
    {"log", no_argument, NULL, T_LOG},
    {NULL, 0, NULL, 0}};

struct TagOptions {
    int threads = 1, quality = 1;
    double percentage = 0.6;
    bool tag_supplementary = false, log = false, cram = false;
    std::string snp_file, sv_file, mod_file, bam, fasta, prefix = "result", region, command = "longphase-s ";
};

}  // namespace

struct lpsh_tag {
    TagOptions opt;
    std::vector<std::string> chr_names;               // VCF_Info::chrVec (##contig order), narrowed by --region
    std::map<std::string, int> chr_length;
    std::map<std::string, std::map<int, lpsh::SampleRecord>> variants;   // phased heterozygous records of the NORMAL sample
    std::map<std::string, std::string> reference;
    lpsh::TagBamIO io;                                // input / output BAM, index, thread pool, region reader
    std::ofstream log;
    lpsh::Chunk chunk;                                // the chunk of the staged API (lpsh_tag_pack / lpsh_tag_emit)
    int chunk_contig = -1;
    size_t chunk_reads = 8192;                        // records per device call; small enough for htslib's asynchronous
                                                      // BGZF decode / encode to stay busy around it (LPS_TAG_CHUNK overrides)
    // ReadStatistics (HaplotagProcess.h:21-45)
    int64_t st_alignment = 0, st_supplementary = 0, st_secondary = 0, st_unmapped = 0, st_tag = 0, st_untag = 0, st_low = 0,
            st_other = 0, st_empty = 0, st_similar = 0, st_no_variant = 0, st_hp[3] = {0, 0, 0};
    std::time_t t_begin = time(NULL);
};

namespace {

int parse_tag_options(int argc, char **argv, TagOptions &o) {   // ArgumentManager::parseOptions + Haplotag.cpp:60-150
    optind = 1;
    bool bad = false;
    for (int c; (c = getopt_long(argc, argv, "s:b:o:t:q:p:r:", TAG_LONG, NULL)) != -1;) {
        switch (c) {
            case 't': lpsh::take(optarg, o.threads); break;
            case 'o': lpsh::take(optarg, o.prefix); break;
            case 'q': lpsh::take(optarg, o.quality); break;
            case 'p': lpsh::take(optarg, o.percentage); break;
            case T_SUP: o.tag_supplementary = true; break;
            case T_REGION: lpsh::take(optarg, o.region); break;
            case T_CRAM: o.cram = true; break;
            case T_LOG: o.log = true; break;
            case 's': lpsh::take(optarg, o.snp_file); break;
            case 'b': lpsh::take(optarg, o.bam); break;
            case 'r': lpsh::take(optarg, o.fasta); break;
            case T_SV: lpsh::take(optarg, o.sv_file); break;
            case T_MOD: lpsh::take(optarg, o.mod_file); break;
            case T_HELP: std::cout << TAG_USAGE << std::endl; return 2;
            default: bad = true;
        }
    }
    for (int i = 0; i < argc; i++) { o.command += argv[i]; o.command += " "; }
    bad |= !lpsh::required_file("haplotag", o.snp_file, "SNP file");
    bad |= !lpsh::required_file("haplotag", o.bam, "BAM file");
    bad |= !lpsh::required_file("haplotag", o.fasta, "reference file");
    if (o.threads < 1) { std::cerr << "[ERROR] haplotag: invalid threads. value: " << o.threads << "\nplease check -t, --threads=Num\n"; bad = true; }
    if (o.percentage > 1 || o.percentage < 0) {
        std::cerr << "[ERROR] haplotag: invalid percentage threshold. value: " << o.percentage
                  << "\nthis value need: 0~1, please check -p, --percentageThreshold=Num\n";
        bad = true;
    }
    if (!o.sv_file.empty() || !o.mod_file.empty()) {
        std::cerr << "[ERROR] haplotag: --sv-file and --mod-file are not available in this build.\n";
        bad = true;
    }
    if (bad) { std::cerr << "\n"; std::cout << TAG_USAGE << std::endl; return 1; }
    return 0;
}

void tag_banner(const TagOptions &o) {   // HaplotagProcess::printParamsMessage
    std::ostream &e = std::cerr;
    e << "LongPhase-S v" << lpsh::REFERENCE_VERSION << " - Haplotag (" << lps_version() << ")\n\n";
    e << "phased SNP file:   " << o.snp_file << "\nphased SV file:    " << o.sv_file << "\nphased MOD file:   " << o.mod_file << "\n";
    e << "input bam file:    " << o.bam << "\ninput ref file:    " << o.fasta << "\noutput bam file:   " << o.prefix + (o.cram ? ".cram" : ".bam") << "\n";
    e << "number of threads: " << o.threads << "\nwrite log file:    " << (o.log ? "true" : "false") << "\n";
    e << "log file:          " << (o.log ? (o.prefix + ".out") : "") << "\n-------------------------------------------\n";
    e << "tag region:                    " << (!o.region.empty() ? o.region : "all") << "\n";
    e << "filter mapping quality below:  " << o.quality << "\npercentage threshold:          " << o.percentage << "\n";
    e << "tag supplementary:             " << (o.tag_supplementary ? "true" : "false") << "\n-------------------------------------------\n";
}

// HaplotagProcess::setProcessingChromRegion (HaplotagProcess.cpp:105-135)
void narrow_to_region(lpsh_tag &job) {
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
}

// getLastVarPos + FastaParser (HaplotagParsingBam.cpp:333-373, ParsingBam.cpp:17-59)
int load_tag_reference(lpsh_tag &job) {
    faidx_t *fai = fai_load(job.opt.fasta.c_str());
    if (!fai) return lpsh::fail("cannot load the FASTA index of " + job.opt.fasta);
    for (const std::string &chr : job.chr_names) {
        int last = 0;
        auto it = job.variants.find(chr);
        if (it != job.variants.end() && !it->second.empty()) last = it->second.rbegin()->first;   // every loaded record has a phase set
        int len = 0;
        char *s = faidx_fetch_seq(fai, chr.c_str(), 0, last + 5, &len);
        if (len == 0) std::cout << "nothing in reference file \n";
        job.reference[chr] = s ? s : "";
        free(s);
    }
    fai_destroy(fai);
    return 0;
}

void write_log_header(lpsh_tag &job) {   // GermlineTagLog::addParamsMessage / writeBasicColumns (HaplotagProcess.cpp:181-208)
    const TagOptions &o = job.opt;
    job.log << "##snpFile:" << o.snp_file << "\n##svFile:" << o.sv_file << "\n##bamFile:" << o.bam << "\n##resultPrefix:" << o.prefix << "\n"
            << "##numThreads:" << o.threads << "\n##region:" << o.region << "\n##qualityThreshold:" << o.quality << "\n"
            << "##percentageThreshold:" << o.percentage << "\n##tagSupplementary:" << o.tag_supplementary << "\n";
    job.log << "#ReadID\tCHROM\tReadStart\tConfidnet(%)\tHaplotype\tPhaseSet\tTotalAllele\tHP1Allele\tHP2Allele\tphasingQuality(PQ)\t(Variant,HP)\t(PhaseSet,Variantcount)\n";
}

void finish_contig(lpsh_tag &job) {
    job.io.end_region();
    job.chunk_contig = -1;
    job.chunk.clear();
}

}  // namespace

extern "C" {

int lpsh_tag_open(int argc, char **argv, lpsh_tag **out) {
    if (!out) return -1;
    *out = nullptr;
    lpsh_tag *job = new lpsh_tag();
    const int rc = parse_tag_options(argc, argv, job->opt);
    if (rc != 0) { delete job; return rc; }
    if (const char *e = getenv("LPS_TAG_CHUNK")) { const long v = atol(e); if (v > 0) job->chunk_reads = (size_t)v; }
    tag_banner(job->opt);
    std::time_t t0 = time(NULL);
    std::cerr << "parsing SNP VCF ... ";
    {
        lpsh::SampleVcf vcf;   // VcfParser::parserProcess for the NORMAL sample (HaplotagVcfParser.cpp:206-345)
        lpsh::load_sample_vcf(job->opt.snp_file, false, vcf);
        job->chr_names.swap(vcf.chr_names);
        job->chr_length.swap(vcf.chr_length);
        job->variants.swap(vcf.records);
    }
    std::cerr << difftime(time(NULL), t0) << "s\n";
    narrow_to_region(*job);
    *out = job;
    return 0;
}

int lpsh_tag_n_contigs(const lpsh_tag *h) { return h ? (int)h->chr_names.size() : 0; }
const char *lpsh_tag_contig_name(const lpsh_tag *h, int i) {
    return (h && i >= 0 && (size_t)i < h->chr_names.size()) ? h->chr_names[(size_t)i].c_str() : nullptr;
}
int lpsh_tag_params(const lpsh_tag *h, lps_tag_params *out) {
    if (!h || !out) return -1;
    memset(out, 0, sizeof(*out));
    out->mapping_quality = h->opt.quality;
    out->mapq_filter = 1;
    out->tag_supplementary = h->opt.tag_supplementary;
    out->have_reference = 1;
    out->percentage_threshold = h->opt.percentage;
    return 0;
}

int lpsh_tag_begin(lpsh_tag *h) {
    if (!h) return -1;
    const TagOptions &o = h->opt;
    if (h->chr_names.empty()) { std::cerr << "[ERROR](HaplotagBamParser): chrVec is empty" << std::endl; return lpsh::fail("the VCF header lists no contig"); }
    if (o.log) {
        h->log.open((o.prefix + ".out").c_str());
        if (!h->log.is_open()) { std::cerr << "Fail to open write file: " << o.prefix + ".out" << "\n"; return lpsh::fail("cannot open the log file"); }
        write_log_header(*h);
    }
    if (load_tag_reference(*h) != 0) return -1;
    // BamFileRAII (HaplotagParsingBam.cpp:20-82): "wb" or "wc"
    return h->io.open(o.bam, o.fasta, o.prefix + (o.cram ? ".cram" : ".bam"), o.cram ? "wc" : lpsh::bam_write_mode(), o.threads, o.command);
}

// next chunk of contig i into `ck`: 1 = filled, 0 = the contig is exhausted, < 0 error
static int read_chunk(lpsh_tag *h, int i, lpsh::Chunk &ck) {
    const std::string &chr = h->chr_names[(size_t)i];
    if (h->io.cur != i) {
        const std::string region = !h->opt.region.empty() ? h->opt.region : chr + ":1-" + std::to_string(h->chr_length[chr]);
        const int rc = h->io.start_region(i, region);
        if (rc < 0) return rc;
    }
    ck.clear();
    lpsh::PackedContig &pc = ck.pack;
    pc.tagged_variants = true;
    auto vars = h->variants.find(chr);
    if (vars != h->variants.end())
        for (const auto &kv : vars->second) {
            const lpsh::SampleRecord &v = kv.second;
            pc.add_variant(kv.first, v.ref, v.alt);
            pc.v_hp1_is_alt.push_back(v.hp1_is_alt);
            pc.v_ps.push_back(v.ps);
            pc.v_gt_kind.push_back(1);   // GenomeType::PHASED_HETERO
        }
    pc.ref_shared = &h->reference[chr];
    const int got = h->io.fill(ck, h->chunk_reads);
    if (got <= 0) { ck.clear(); return got; }
    pc.finish();
    return 1;
}

// ---- LPS_TAG_READERS=K: position slices of every contig's region, read by K threads with their own file handles -------------------
namespace {
struct TagSlice { int contig; int tid; hts_pos_t b, e; bool first; };
struct TagReader { samFile *in = nullptr; bam_hdr_t *hdr = nullptr; };

std::vector<TagSlice> tag_slices(lpsh_tag *h) {
    hts_pos_t step = 1000000;
    if (const char *e = getenv("LPS_TAG_SLICE_BP")) { const long v = atol(e); if (v >= 1000) step = v; }
    std::vector<TagSlice> out;
    for (size_t i = 0; i < h->chr_names.size(); i++) {
        const std::string &chr = h->chr_names[i];
        const std::string region = !h->opt.region.empty() ? h->opt.region : chr + ":1-" + std::to_string(h->chr_length[chr]);
        hts_itr_t *it = sam_itr_querys(h->io.idx, h->io.hdr, region.c_str());
        if (!it) continue;
        hts_pos_t end = it->end;
        const int64_t len = h->io.hdr->target_len && it->tid >= 0 ? (int64_t)h->io.hdr->target_len[it->tid] : 0;
        if (len > 0 && end > len) end = len;                 // "chr" or "chr:start" leave the end open
        for (hts_pos_t b = it->beg; b < end; b += step) out.push_back(TagSlice{(int)i, it->tid, b, std::min(b + step, end), b == it->beg});
        hts_itr_destroy(it);
    }
    return out;
}

int read_tag_slice(lpsh_tag *h, const TagSlice &sl, TagReader &rd, lpsh::Chunk &ck) {
    if (!rd.in) {
        rd.in = hts_open(h->opt.bam.c_str(), "r");
        if (!rd.in) return lpsh::fail("Cannot open bam file " + h->opt.bam);
        hts_set_fai_filename(rd.in, h->opt.fasta.c_str());
        rd.hdr = sam_hdr_read(rd.in);
        if (!rd.hdr) return lpsh::fail("Cannot read header from bam file " + h->opt.bam);
    }
    const std::string &chr = h->chr_names[(size_t)sl.contig];
    lpsh::PackedContig &pc = ck.pack;
    pc.tagged_variants = true;
    auto vars = h->variants.find(chr);
    if (vars != h->variants.end())
        for (const auto &kv : vars->second) {
            pc.add_variant(kv.first, kv.second.ref, kv.second.alt);
            pc.v_hp1_is_alt.push_back(kv.second.hp1_is_alt);
            pc.v_ps.push_back(kv.second.ps);
            pc.v_gt_kind.push_back(1);
        }
    pc.ref_shared = &h->reference[chr];
    hts_itr_t *it = sam_itr_queryi(h->io.idx, sl.tid, sl.b, sl.e);
    if (!it) return lpsh::fail("cannot query " + h->opt.bam);
    bam1_t *b = bam_init1();
    while (sam_itr_next(rd.in, it, b) >= 0) {
        if (!sl.first && b->core.pos < sl.b) continue;       // starts in an earlier slice: taken there
        pc.add_alignment(b);
        ck.records.push_back(b);
        b = bam_init1();
    }
    bam_destroy1(b);
    hts_itr_destroy(it);
    pc.finish();
    return 0;
}
}  // namespace

int lpsh_tag_pack(lpsh_tag *h, int i, lpsh_packed *out) {
    if (!h || !out || i < 0 || (size_t)i >= h->chr_names.size() || !h->io.in) return -1;
    const int got = read_chunk(h, i, h->chunk);
    if (got == 1) { h->chunk.pack.view(out); h->chunk_contig = i; }
    return got;
}

static int emit_chunk(lpsh_tag *h, int i, lpsh::Chunk &ck, const lps_tag_result *r);

int lpsh_tag_emit(lpsh_tag *h, int i, const lps_tag_result *r) {
    if (!h || !r || h->chunk_contig != i || !h->io.has_output()) return -1;
    return emit_chunk(h, i, h->chunk, r);
}

static int emit_chunk(lpsh_tag *h, int i, lpsh::Chunk &ck, const lps_tag_result *r) {
    const size_t n = ck.records.size();
    if ((size_t)r->n_reads != n) return lpsh::fail("verdict count does not match the chunk");
    const bool want_log = h->log.is_open();
    if (want_log && !r->call_off)
        for (size_t k = 0; k < n; k++)
            if (r->category[k] == LPS_TAG_PROCESSED) return lpsh::fail("--log needs lps_tag_reads with want_calls");
    const std::string &chr = h->chr_names[(size_t)i];
    for (size_t k = 0; k < n; k++) {
        bam1_t *b = ck.records[k];
        const int cat = r->category[k];
        h->st_alignment++;
        if (cat != LPS_TAG_PROCESSED) {
            h->st_untag++;
            if (cat == LPS_TAG_LOW_MAPQ) h->st_low++;
            else if (cat == LPS_TAG_UNMAPPED) h->st_unmapped++;
            else if (cat == LPS_TAG_SECONDARY) h->st_secondary++;
            else if (cat == LPS_TAG_SUPPLEMENTARY) h->st_supplementary++;
            else if (cat == LPS_TAG_EMPTY_VARIANTS) h->st_empty++;
            else h->st_other++;
        } else {
            // GermlineHaplotagChrProcessor::processRead (HaplotagProcess.cpp:318-355)
            if (b->core.flag & 0x800) h->st_supplementary++;
            int hp = r->hp[k], ps = r->ps[k], pq = r->pq[k];
            const double h1 = r->h1[k], h2 = r->h2[k];
            const double mx = h1 > h2 ? h1 : h2, mn = h1 > h2 ? h2 : h1;
            if (mx / (mx + mn) < h->opt.percentage) h->st_similar++;
            if (mx == 0) h->st_no_variant++;
            if (want_log) {   // GermlineTagLog::writeTagReadLog (HaplotagProcess.cpp:210-237)
                std::map<int, int> variants_hp, count_ps;
                for (uint64_t c = r->call_off[k]; c < r->call_off[k + 1]; c++) {
                    const lps_call &cl = r->calls[c];
                    if (cl.allele >= 0) variants_hp[ck.pack.v_pos[(size_t)cl.var]] = cl.allele;
                    count_ps[ck.pack.v_ps[(size_t)cl.var]]++;
                }
                h->log << bam_get_qname(b) << "\t" << chr << "\t" << b->core.pos << "\t" << (mx / (mx + mn)) << "\tH"
                       << (hp == 0 ? std::string(".") : std::to_string(hp)) << "\t"
                       << (hp == 0 ? std::string(".") : std::to_string(count_ps.begin()->first)) << "\t" << (int)(h1 + h2) << "\t" << (int)h1
                       << "\t" << (int)h2 << "\t" << pq << "\t";
                for (const auto &v : variants_hp) h->log << " " << v.first << "," << v.second;
                h->log << "\t";
                for (const auto &v : count_ps) h->log << " " << v.first << "," << v.second;
                h->log << "\n";
            }
            lpsh::drop_aux(b, "HP");
            lpsh::drop_aux(b, "PS");
            lpsh::drop_aux(b, "PQ");
            if (hp != 0) {
                h->st_hp[hp]++;
                h->st_tag++;
                bam_aux_append(b, "HP", 'i', sizeof(int), (uint8_t *)&hp);
                bam_aux_append(b, "PS", 'i', sizeof(int), (uint8_t *)&ps);
                bam_aux_append(b, "PQ", 'i', sizeof(int), (uint8_t *)&pq);
            } else {
                h->st_hp[0]++;
                h->st_untag++;
            }
        }
        if (h->io.write(b) < 0) { std::cerr << "[ERROR](BamFileRAII): write output bam file failed" << std::endl; return lpsh::fail("write output bam file failed"); }
    }
    ck.clear();
    return 0;
}

int lpsh_tag_end(lpsh_tag *h) {
    if (!h) return -1;
    finish_contig(*h);
    const int rc = h->io.close();
    if (h->log.is_open()) h->log.close();
    std::ostream &e = std::cerr;   // HaplotagProcess::printExecutionReport (HaplotagProcess.cpp:152-175)
    e << "-------------------------------------------\n";
    e << "total process time        : " << difftime(time(NULL), h->t_begin) << "s\n";
    e << "total alignment           : " << h->st_alignment << "\ntotal supplementary       : " << h->st_supplementary << "\n";
    e << "total secondary           : " << h->st_secondary << "\ntotal unmapped            : " << h->st_unmapped << "\n";
    e << "total tagged alignments   : " << h->st_tag << "\n    L----total HP1        : " << h->st_hp[1] << "\n    L----total HP2        : " << h->st_hp[2] << "\n";
    e << "    L----total HP1-1      : 0\n    L----total HP2-1      : 0\n    L----total HP3        : 0\n         L----only H3 SNP : 0\n";
    e << "total untagged            : " << h->st_untag << "\n    L----lower mapping quality        : " << h->st_low << "\n";
    e << "    L----no variant                   : " << h->st_empty << "\n    L----start pos > last variant pos : " << h->st_other << "\n";
    e << "    L----judge to untag               : " << h->st_hp[0] << "\n         L----high similarity         : " << h->st_similar << "\n";
    e << "         L----cross two block         : 0\n         L----no variant judge HP     : " << h->st_no_variant << "\n";
    e << "-------------------------------------------\n";
    return rc;
}

// The tagging pass with any judge: reader thread (htslib parsing + packing) | calling thread (judge, tags, sam_write1).
int lpsh_tag_run_with(lpsh_tag *h, lpsh_tag_judge_fn judge, void *user) {
    if (!h || !judge) return -1;
    if (lpsh_tag_begin(h) != 0) return -1;
    lps_tag_params tp;
    lpsh_tag_params(h, &tp);
    std::time_t t0 = time(NULL);
    std::cerr << "tag read start ...\n";
    double ms_read = 0, ms_judge = 0, ms_emit = 0;   // wall clock per stage (the reader runs on its own thread)
    long n_chunks = 0;
    auto handle = [&](int i, lpsh::Chunk &ck) -> int {
        lpsh_packed v;
        ck.pack.view(&v);
        lps_tag_result r;
        memset(&r, 0, sizeof(r));
        const double j0 = lpsh::now_ms();
        n_chunks++;
        std::vector<uint8_t> cat;
        std::vector<int8_t> hp;
        std::vector<int32_t> zero;
        if (v.variants.n == 0) {
            // a contig without variants never reaches the tagger: the dispatch of processSingleChrom (HaplotagParsingBam.cpp:457-476)
            // only looks at MAPQ and flags, so no device work exists for these records
            const int n = v.batch.n_reads;
            cat.resize((size_t)n); hp.assign((size_t)n, 0); zero.assign((size_t)n, 0);
            for (int k = 0; k < n; k++) {
                const int flag = v.batch.flag[k];
                cat[(size_t)k] = lpsh::category_without_variants(v.batch.mapq[k], flag, tp);
            }
            r.n_reads = n; r.category = cat.data(); r.hp = hp.data(); r.ps = zero.data(); r.pq = zero.data(); r.h1 = zero.data(); r.h2 = zero.data();
        } else if (judge(user, i, &v, h->opt.log ? 1 : 0, &r) != 0) {
            return lpsh::fail(std::string("contig ") + h->chr_names[(size_t)i] + ": the judge failed");
        }
        const double j1 = lpsh::now_ms();
        const int rc_emit = emit_chunk(h, i, ck, &r);
        ms_judge += j1 - j0;
        ms_emit += lpsh::now_ms() - j1;
        return rc_emit;
    };
    const double m0 = lpsh::now_ms();
    int n_readers = 1;
    if (const char *e = getenv("LPS_TAG_READERS")) { const int v = atoi(e); if (v >= 1 && v <= 256) n_readers = v; }
    int rc;
    if (n_readers > 1 && !lpsh::gpu_inflate_requested()) {
        // K readers on position slices, handled in order (opt-in: pays off once the BAM writer is no longer what the pass waits for)
        const std::vector<TagSlice> slices = tag_slices(h);
        std::vector<TagReader> rd((size_t)n_readers);
        rc = lpsh::run_ordered_slices(slices.size(), n_readers, (size_t)n_readers * 2,
                                      [&](size_t s, int r, lpsh::Chunk &ck) { return read_tag_slice(h, slices[s], rd[(size_t)r], ck); },
                                      [&](size_t s, lpsh::Chunk &ck) { return ck.records.empty() ? 0 : handle(slices[s].contig, ck); });
        for (TagReader &x : rd) { if (x.hdr) bam_hdr_destroy(x.hdr); if (x.in) sam_close(x.in); }
    } else {
        rc = lpsh::run_chunk_pipeline((int)h->chr_names.size(), [&](int i, lpsh::Chunk &ck) {
            const double r0 = lpsh::now_ms();
            const int got = read_chunk(h, i, ck);
            ms_read += lpsh::now_ms() - r0;
            return got;
        }, handle);
    }
    std::cerr << "tag read " << difftime(time(NULL), t0) << "s\n";
    std::cerr << "[timing] tagging pass " << (lpsh::now_ms() - m0) << " ms: " << n_chunks << " chunks; reader thread " << ms_read << " ms (parse + pack); "
              << "judge " << ms_judge << " ms; tag + write " << ms_emit << " ms\n";
    const int rc_end = lpsh_tag_end(h);
    return rc != 0 ? -1 : rc_end;
}

namespace {
// the device as judge: lps_tag_reads on the chunk; the contig's tables are set when the contig changes
struct DeviceJudge {
    double ms_create = 0, ms_tables = 0, ms_submit = 0, ms_tag = 0;
    lps_ctx *ctx = nullptr;
    lps_tag_params tp;
    int contig = -1;
    std::thread warm;
    std::string error;
};
int device_judge(void *user, int contig, const lpsh_packed *v, int want_calls, lps_tag_result *out) {
    DeviceJudge *d = (DeviceJudge *)user;
    const double t0 = lpsh::now_ms();
    if (!d->ctx) {
        if (d->warm.joinable()) d->warm.join();
        if (lps_ctx_create(0, &d->ctx) != 0) { d->error = "no usable CUDA device (there is no CPU fallback)"; return -1; }
    }
    const double t1 = lpsh::now_ms();
    int rc = 0;
    if (d->contig != contig) {
        rc = lps_contig_set_reference(d->ctx, v->ref, v->ref_len);
        if (rc == 0) rc = lps_contig_set_variants(d->ctx, &v->variants, 0);
        d->contig = contig;
    }
    const double t2 = lpsh::now_ms();
    if (rc == 0) rc = lps_batch_submit(d->ctx, &v->batch);
    const double t3 = lpsh::now_ms();
    if (rc == 0) rc = lps_tag_reads(d->ctx, &d->tp, want_calls, out);
    d->ms_create += t1 - t0; d->ms_tables += t2 - t1; d->ms_submit += t3 - t2; d->ms_tag += lpsh::now_ms() - t3;
    if (rc != 0) d->error = lps_last_error(d->ctx);
    return rc;
}
}  // namespace

int lpsh_tag_run(lpsh_tag *h) {
    if (!h) return -1;
    DeviceJudge d;   // no device -> the first chunk fails with "no usable CUDA device"; nothing is ever judged on the host
    lpsh_tag_params(h, &d.tp);
    h->io.device_pass = true;   // the tagged BAM is deflated on the device as well (LPS_GPU_DEFLATE=0: htslib's writer)
    const int rc = lpsh_tag_run_with(h, device_judge, &d);
    std::cerr << "[timing] judge on the device: context " << d.ms_create << " ms, contig tables " << d.ms_tables << " ms, lps_batch_submit " << d.ms_submit
              << " ms, lps_tag_reads " << d.ms_tag << " ms\n";
    if (rc != 0 && !d.error.empty()) lpsh::fail(d.error);
    if (d.ctx) lps_ctx_destroy(d.ctx);
    return rc;
}

void lpsh_tag_close(lpsh_tag *h) {
    if (!h) return;
    if (h->io.is_open()) lpsh_tag_end(h);
    delete h;
}

int lpsh_tag_main(int argc, char **argv) {
    std::thread warm = lpsh::warm_up_device();   // the driver starts while the VCF and the FASTA are read
    struct Join { std::thread &t; ~Join() { if (t.joinable()) t.join(); } } join_warm{warm};
    lpsh_tag *job = nullptr;
    const int rc = lpsh_tag_open(argc, argv, &job);
    if (rc == 2) return 0;
    if (rc != 0) { if (rc < 0) std::cerr << "haplotag: " << lpsh_last_error() << "\n"; return 1; }
    const int run = lpsh_tag_run(job);
    if (run != 0) std::cerr << "[ERROR] haplotag: " << lpsh_last_error() << "\n";
    lpsh_tag_close(job);
    return run != 0 ? 1 : 0;
}

}  // extern "C"
