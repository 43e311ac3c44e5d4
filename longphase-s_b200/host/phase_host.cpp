// phase_host.cpp — the `phase` sub-command above the C ABI: options, VCF / FASTA / BAM loading with htslib, SoA packing,
// the contig loop on the GPU(s), and the phased-VCF writer.  See lps_host.h for the reference seams each stage replaces.
//
// Scope notes: --sv-file, --mod-file and --dot are accepted by the option parser (same names as the reference) but rejected with
// a message: SV / MOD inputs and the graph dump are outside the hot path this repository rebuilds.
#include "host_common.h"

#include <getopt.h>
#include <omp.h>

#include <cmath>
#include <ctime>
#include <fstream>
#include <iterator>
#include <limits>
#include <set>
#include <sstream>

namespace {

const char *PHASE_USAGE =
    "Usage:  phase [OPTION] ... READSFILE\n"
    "   --help                      display this help and exit.\n\n"
    "required:\n"
    "   -s, --snp-file=NAME         input SNP vcf file.\n"
    "   -b, --bam-file=NAME         input bam file (may be repeated).\n"
    "   -r, --reference=NAME        reference fasta.\n"
    "   --ont | --pb                sequencing platform.\n\n"
    "optional:\n"
    "   -t, --threads=Num           contigs in flight / BAM decoding threads. default:1\n"
    "   -o, --out-prefix=NAME       prefix of phasing result. default: result\n"
    "   --indels                    phase small indels. default: False\n"
    "   --indelQuality=Num          drop indels with QUAL below the threshold (with --indels). default: 0\n"
    "   -q, --mappingQuality=Num    default:1          -x, --mismatchRate=Num      default:3\n"
    "   -p, --baseQuality=[0~90]    default:12         -e, --edgeWeight=[0~1]      default:0.1\n"
    "   -a, --connectAdjacent=Num   default:35         -d, --distance=Num          default:300000\n"
    "   -1, --edgeThreshold=[0~1]   default:0.7        -L, --overlapThreshold=[0~1] default:0.2\n"
    "   -m, --readConfidence=[0.5~1] default:0.65      -n, --snpConfidence=[0.5~1] default:0.75\n"
    "   --deepsomatic_output        keep only GERMLINE records of a DeepSomatic VCF and re-genotype them from AD / VAF first.\n"
    "not available in this build: --sv-file, --mod-file, --dot\n\n";

enum { O_HELP = 1, O_DOT, O_SV, O_MOD, O_ONT, O_PB, O_INDELS, O_INDELQ, O_DEEPSOMATIC, O_VERSION };

const struct option PHASE_LONG[] = {
    {"help", no_argument, NULL, O_HELP},
    {"dot", no_argument, NULL, O_DOT},
    {"ont", no_argument, NULL, O_ONT},
    {"pb", no_argument, NULL, O_PB},
    {"version", no_argument, NULL, O_VERSION},
    {"indels", no_argument, NULL, O_INDELS},
    {"indelQuality", required_argument, NULL, O_INDELQ},
    {"deepsomatic_output", no_argument, NULL, O_DEEPSOMATIC},
    {"sv-file", required_argument, NULL, O_SV},
    {"mod-file", required_argument, NULL, O_MOD},
    {"reference", required_argument, NULL, 'r'},
    {"snp-file", required_argument, NULL, 's'},
    {"bam-file", required_argument, NULL, 'b'},
    {"out-prefix", required_argument, NULL, 'o'},
    {"threads", required_argument, NULL, 't'},
    {"distance", required_argument, NULL, 'd'},
    {"edgeThreshold", required_argument, NULL, '1'},
    {"connectAdjacent", required_argument, NULL, 'a'},
    {"mappingQuality", required_argument, NULL, 'q'},
    {"mismatchRate", required_argument, NULL, 'x'},
    {"baseQuality", required_argument, NULL, 'p'},
    {"edgeWeight", required_argument, NULL, 'e'},
    {"snpConfidence", required_argument, NULL, 'n'},
    {"readConfidence", required_argument, NULL, 'm'},
    {"overlapThreshold", required_argument, NULL, 'L'},
    {"svWindow", required_argument, NULL, 'w'},
    {"svThreshold", required_argument, NULL, 'h'},
    {NULL, 0, NULL, 0}};

struct PhaseOptions {
    int threads = 1, distance = 300000, connect_adjacent = 35, mapping_quality = 1, base_quality = 12, indel_quality = 0;
    int sv_window = 20;
    double mismatch_rate = 3, edge_weight = 0.1, snp_confidence = 0.75, read_confidence = 0.65, edge_threshold = 0.7;
    double overlap_threshold = 0.2, sv_threshold = 0.1;
    bool ont = false, pb = false, indels = false, dot = false, deepsomatic = false;
    std::string snp_file, sv_file, mod_file, fasta, prefix = "result", command;
    std::vector<std::string> bams;
};

bool readable(const std::string &path) { return std::ifstream(path.c_str()).is_open(); }

struct VariantText {
    std::string ref, alt;
};

struct PhasedCall {
    int block;
    char hap_ref, hap_alt;
};

}  // namespace

struct lpsh_phase {
    PhaseOptions opt;
    std::vector<std::string> chr_names;
    std::map<std::string, std::map<int, VariantText>> variants;
    std::map<std::string, std::set<int>> low_qual_indels;
    std::map<std::string, std::string> reference;
    std::map<std::string, std::map<int, PhasedCall>> phased;   // PhasingResult, keyed (contig, position) instead of "chr_pos"
    std::vector<lpsh::PackedContig *> packed;
    std::ofstream removed_log;
    bool wrote_command_line = false;
};

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// options (Phasing.cpp:118-330)
int parse_phase_options(int argc, char **argv, PhaseOptions &o) {
    optind = 1;
    for (int c; (c = getopt_long(argc, argv, "s:b:o:t:r:d:1:a:q:x:p:e:n:m:L:w:h:", PHASE_LONG, NULL)) != -1;) {
        switch (c) {
            case 's': lpsh::take(optarg, o.snp_file); break;
            case 't': lpsh::take(optarg, o.threads); break;
            case 'o': lpsh::take(optarg, o.prefix); break;
            case 'r': lpsh::take(optarg, o.fasta); break;
            case 'd': lpsh::take(optarg, o.distance); break;
            case '1': lpsh::take(optarg, o.edge_threshold); break;
            case 'a': lpsh::take(optarg, o.connect_adjacent); break;
            case 'q': lpsh::take(optarg, o.mapping_quality); break;
            case 'x': lpsh::take(optarg, o.mismatch_rate); break;
            case 'p': lpsh::take(optarg, o.base_quality); break;
            case 'e': lpsh::take(optarg, o.edge_weight); break;
            case 'n': lpsh::take(optarg, o.snp_confidence); break;
            case 'm': lpsh::take(optarg, o.read_confidence); break;
            case 'w': lpsh::take(optarg, o.sv_window); break;
            case 'h': lpsh::take(optarg, o.sv_threshold); break;
            case 'L': lpsh::take(optarg, o.overlap_threshold); break;
            case 'b': { std::string f; lpsh::take(optarg, f); o.bams.push_back(f); break; }
            case O_SV: lpsh::take(optarg, o.sv_file); break;
            case O_MOD: lpsh::take(optarg, o.mod_file); break;
            case O_INDELS: o.indels = true; break;
            case O_INDELQ: lpsh::take(optarg, o.indel_quality); break;
            case O_DEEPSOMATIC: o.deepsomatic = true; break;
            case O_DOT: o.dot = true; break;
            case O_ONT: o.ont = true; break;
            case O_PB: o.pb = true; break;
            case O_HELP: std::cout << PHASE_USAGE; return 2;
        }
    }
    for (int i = 0; i < argc; i++) { o.command += argv[i]; o.command += " "; }
    bool bad = false;
    auto complain = [&](const std::string &what) { std::cerr << what; bad = true; };
    if (!o.ont && !o.pb) complain("phase: missing arguments. --ont or --pb\n");
    if (o.ont && o.pb) complain("phase: conflict arguments. --ont or --pb\n");
    if (o.snp_file.empty()) complain("phase: missing SNP file.\n");
    else if (!readable(o.snp_file)) complain("File " + o.snp_file + " not exist.\n\n");
    if (o.fasta.empty()) complain("phase: missing reference.\n");
    else if (!readable(o.fasta)) complain("File " + o.fasta + " not exist.\n\n");
    auto range = [&](bool ok, const char *name, double v, const char *hint) {
        if (!ok) { std::ostringstream m; m << "phase invalid " << name << ". value: " << v << "\n please check " << hint << "\n"; complain(m.str()); }
    };
    range(o.threads >= 1, "threads", o.threads, "-t, --threads=Num");
    range(o.distance >= 0, "distance", o.distance, "-d or --distance=Num");
    range(o.connect_adjacent >= 0, "connectAdjacent", o.connect_adjacent, "-a, --connectAdjacent=Num");
    range(o.mapping_quality >= 0, "mappingQuality", o.mapping_quality, "-m, --mappingQuality=Num");
    range(o.mismatch_rate >= 0, "mismatchRate", o.mismatch_rate, "-x, --mismatchRate=Num");
    range(o.base_quality >= 0, "baseQuality", o.base_quality, "-m, --mappingQuality=[0~90]");
    range(o.edge_weight >= 0, "edgeWeight", o.edge_weight, "-e, --edgeWeight=[0~1]");
    range(o.edge_threshold >= 0 && o.edge_threshold <= 1, "edgeThreshold", o.edge_threshold, "-1, --edgeThreshold=[0~1]");
    range(o.overlap_threshold >= 0 && o.overlap_threshold <= 1, "overlapThreshold", o.overlap_threshold, "-L, --overlapThreshold=[0~1]");
    range(o.read_confidence >= 0.5 && o.read_confidence <= 1, "readConfidence", o.read_confidence, "-m, --readConfidence=[0.5~1]");
    range(o.snp_confidence >= 0.5 && o.snp_confidence <= 1, "snpConfidence", o.snp_confidence, "-n, --snpConfidence=[0.5~1]");
    if (!o.sv_file.empty() || !o.mod_file.empty() || o.dot)
        complain("phase: --sv-file, --mod-file and --dot are not available in this build.\n");
    if (o.connect_adjacent > 127) complain("phase: --connectAdjacent above 127 is not supported by the device path.\n");
    if (bad) { std::cerr << "\n" << PHASE_USAGE; return 1; }
    return 0;
}

void phase_banner(const PhaseOptions &o) {   // PhasingProcess.cpp:7-43
    std::ostream &e = std::cerr;
    e << "LongPhase Ver " << lpsh::REFERENCE_VERSION << " (" << lps_version() << ")\n\n--- File Parameter --- \n";
    e << "SNP File           : " << o.snp_file << "\nSV  File           : " << o.sv_file << "\nMOD File           : " << o.mod_file << "\n";
    e << "REF File           : " << o.fasta << "\nOutput Prefix      : " << o.prefix << "\nNumber of Threads  : " << o.threads << "\n";
    e << "Generate Dot       : " << (o.dot ? "True" : "False") << "\nBAM File           : ";
    for (const std::string &b : o.bams) e << b << " ";
    e << "\n\n--- Phasing Parameter --- \nSeq Platform       : " << (o.ont ? "ONT" : "PB") << "\n";
    e << "Phase Indel        : " << (o.indels ? "True" : "False") << "\n";
    if (o.indels) e << "Indel Quality      : " << o.indel_quality << "\n";
    e << "Distance Threshold : " << o.distance << "\nConnect Adjacent   : " << o.connect_adjacent << "\nEdge Threshold     : " << o.edge_threshold << "\n";
    e << "Overlap Threshold  : " << o.overlap_threshold << "\nMapping Quality    : " << o.mapping_quality << "\nMismatch Rate      : " << o.mismatch_rate << "\n";
    e << "Variant Confidence : " << o.snp_confidence << "\nReadTag Confidence : " << o.read_confidence << "\n";
    e << "DeepSomatic Mode   : " << (o.deepsomatic ? "True" : "False") << "\n\n";
}

// ---------------------------------------------------------------------------------------------------------------------
// VCF -> het variants per contig (SnpParser::SnpParser, ParsingBam.cpp:219-359).  The same htslib calls in the same order,
// so that even the reference's peek behind the ALT string (its multi-allelic guard, :285 and :334) sees the same bytes.
bool unphased_or_phased_het(const int *gt, int n) {
    if (n < 2) return false;
    return (gt[0] == 2 && gt[1] == 4) || (gt[0] == 4 && gt[1] == 2) || (gt[0] == 2 && gt[1] == 5) || (gt[0] == 4 && gt[1] == 3);
}

// SnpParser::preprocessDeepsomaticVCF (ParsingBam.cpp:651-834): records whose FILTER mentions GERMLINE survive; their genotype becomes
// the allele pair (a <= b) whose expected fractions (1 or 0.5 / 0.5) are closest, in squared error, to the observed allele fractions
// taken from AD, or from VAF when AD is unusable.
std::vector<std::string> split_on(const std::string &text, char sep) {   // std::getline's splitting: no trailing empty piece
    std::vector<std::string> out;
    std::istringstream in(text);
    for (std::string piece; std::getline(in, piece, sep);) out.push_back(piece);
    return out;
}

void preprocess_deepsomatic(const std::string &in_path, const std::string &out_path) {
    std::ifstream in(in_path.c_str());
    std::ofstream out(out_path.c_str());
    if (!in.is_open()) { std::cerr << "Fail to open input VCF: " << in_path << "\n"; exit(1); }
    if (!out.is_open()) { std::cerr << "Fail to open output VCF: " << out_path << "\n"; exit(1); }
    for (std::string line; std::getline(in, line);) {
        if (line.compare(0, 1, "#") == 0) { out << line << "\n"; continue; }
        std::istringstream split(line);
        std::vector<std::string> f((std::istream_iterator<std::string>(split)), std::istream_iterator<std::string>());
        if (f.size() < 10 || f[6].find("GERMLINE") == std::string::npos) continue;
        const std::vector<std::string> keys = split_on(f[8], ':');
        std::vector<std::string> values = split_on(f[9], ':');
        int gt = -1, vaf = -1, ad = -1;
        for (size_t k = 0; k < keys.size(); k++) {
            if (keys[k] == "GT") gt = (int)k;
            if (keys[k] == "VAF") vaf = (int)k;
            if (keys[k] == "AD") ad = (int)k;
        }
        if (gt >= 0 && gt < (int)values.size()) {
            int n_alt = 0;
            if (!f[4].empty() && f[4] != ".") for (const std::string &a : split_on(f[4], ',')) n_alt += !a.empty();
            const int n_allele = n_alt + 1;
            std::vector<double> seen;
            if (ad >= 0 && ad < (int)values.size()) {
                std::vector<long long> counts;
                for (const std::string &tok : split_on(values[(size_t)ad], ',')) {
                    long long c = 0;
                    if (tok != "." && !tok.empty()) { try { c = std::stoll(tok); } catch (...) { c = 0; } }
                    counts.push_back(c);
                }
                long long total = 0;
                for (long long c : counts) total += c;
                if (total > 0 && (int)counts.size() == n_allele) for (long long c : counts) seen.push_back((double)c / (double)total);
            }
            if (seen.empty() && vaf >= 0 && vaf < (int)values.size()) {
                std::vector<double> alts;
                for (const std::string &tok : split_on(values[(size_t)vaf], ',')) {
                    if (tok == "." || tok.empty()) continue;
                    try { alts.push_back(std::stod(tok)); } catch (...) {}
                }
                if (n_alt == (int)alts.size() && n_alt >= 1) {
                    double sum = 0.0;
                    for (double v : alts) sum += v;
                    seen.push_back(std::max(0.0, 1.0 - sum));
                    seen.insert(seen.end(), alts.begin(), alts.end());
                }
            }
            if (!seen.empty()) {
                int best_a = 0, best_b = 0;
                double best = std::numeric_limits<double>::infinity();
                for (int a = 0; a < n_allele; a++)
                    for (int b = a; b < n_allele; b++) {
                        double cost = 0.0;
                        for (int i = 0; i < n_allele; i++) {
                            const double want = a == b ? (i == a ? 1.0 : 0.0) : ((i == a || i == b) ? 0.5 : 0.0);
                            const double diff = seen[(size_t)i] - want;
                            cost += diff * diff;
                        }
                        if (cost < best) { best = cost; best_a = a; best_b = b; }
                    }
                values[(size_t)gt] = std::to_string(best_a) + "/" + std::to_string(best_b);
                f[9].clear();
                for (size_t k = 0; k < values.size(); k++) f[9] += (k ? ":" : "") + values[k];
            }
        }
        for (size_t k = 0; k < f.size(); k++) out << (k ? "\t" : "") << f[k];
        out << "\n";
    }
}

int load_phase_vcf(lpsh_phase &job) {
    const PhaseOptions &o = job.opt;
    const bool quality_gate = o.indels && o.indel_quality > 0;
    if (quality_gate) {
        job.removed_log.open((o.prefix + "_removed_indels.log").c_str());
        if (job.removed_log.is_open()) job.removed_log << "#CHROM\tPOS\tREF\tALT\tQUAL\n";
    }
    htsFile *in = bcf_open(o.snp_file.c_str(), "r");
    if (!in) return lpsh::fail("cannot open " + o.snp_file);
    bcf_hdr_t *hdr = bcf_hdr_read(in);
    if (!hdr) { bcf_close(in); return lpsh::fail("cannot read the VCF header of " + o.snp_file); }
    int n_seq = 0;
    const char **seq = bcf_hdr_seqnames(hdr, &n_seq);
    for (int i = 0; i < n_seq; i++) job.chr_names.push_back(seq[i]);
    if (bcf_hdr_set_samples(hdr, "-", 0) != 0)
        std::cout << "error or a positive integer if the list contains samples not present in the VCF header\n";
    bcf1_t *rec = bcf_init();
    int *gt = NULL, gt_cap = 0;
    while (bcf_read(in, hdr, rec) == 0) {
        const bool snp = bcf_is_snp(rec);
        if (!snp && !o.indels) continue;
        const int n_gt = bcf_get_format_int32(hdr, rec, "GT", &gt, &gt_cap);
        if (n_gt < 0) { std::cerr << "pos " << rec->pos << " missing GT value\n"; exit(1); }
        if (!unphased_or_phased_het(gt, n_gt)) continue;
        const std::string chr = seq[rec->rid];
        const int pos = (int)rec->pos;
        VariantText v;
        v.ref = rec->d.allele[0];
        v.alt = rec->d.allele[1];
        if (snp) {
            if (rec->d.allele[1][2] != '\0') continue;           // a third allele follows the ALT base
        } else {
            const float q = std::isnan(rec->qual) ? 0.0f : rec->qual;
            if (o.indel_quality > 0 && q < o.indel_quality) {
                if (job.removed_log.is_open())
                    job.removed_log << chr << "\t" << (pos + 1) << "\t" << v.ref << "\t" << v.alt << "\t"
                                    << (std::isnan(rec->qual) ? std::string(".") : std::to_string(rec->qual)) << "\n";
                job.low_qual_indels[chr].insert(pos);
                continue;
            }
            if (rec->d.allele[1][v.alt.size() + 1] != '\0') continue;
        }
        job.variants[chr][pos] = v;
    }
    free(gt);
    free(seq);
    bcf_destroy(rec);
    bcf_hdr_destroy(hdr);
    bcf_close(in);
    return 0;
}

int last_variant(const lpsh_phase &job, const std::string &chr) {
    auto it = job.variants.find(chr);
    if (it == job.variants.end() || it->second.empty()) return -1;
    return it->second.rbegin()->first;
}

// FastaParser (ParsingBam.cpp:17-59): bases 0 .. lastSNP+5 of every contig that has a variant
int load_phase_reference(lpsh_phase &job) {
    faidx_t *fai = fai_load(job.opt.fasta.c_str());
    if (!fai) return lpsh::fail("cannot load the FASTA index of " + job.opt.fasta);
    for (const std::string &chr : job.chr_names) {
        job.reference[chr] = "";
        const int last = last_variant(job, chr);
        if (last == -1) continue;
        int len = 0;
        char *s = faidx_fetch_seq(fai, chr.c_str(), 0, last + 5, &len);
        if (len == 0) std::cout << "nothing in reference file \n";
        if (s) { job.reference[chr] = s; free(s); }
    }
    fai_destroy(fai);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// phased VCF (SnpParser::writeLine, ParsingBam.cpp:460-635).  The reference edits the text of every record in place; the
// helpers below name what its index arithmetic does.
int subfield_of(const std::string &format, size_t at) { return (int)std::count(format.begin(), format.begin() + at, ':'); }
size_t subfield_start(const std::string &sample, int k) {   // first character of the k-th ':' separated value (size() if there are fewer)
    size_t i = 0;
    for (int seen = 0; i < sample.size() && seen < k; i++) seen += sample[i] == ':';
    return i;
}
char peek(const std::string &s, size_t i) { return i < s.size() ? s[i] : '\0'; }

void write_phase_line(lpsh_phase &job, const std::string &line, bool &ps_defined, std::ostream &out) {
    const PhaseOptions &o = job.opt;
    const bool quality_gate = o.indels && o.indel_quality > 0;
    if (line.compare(0, 2, "##") == 0) {
        if (line.compare(0, 16, "##FORMAT=<ID=PS,") == 0) ps_defined = true;
        out << line << "\n";
        if (quality_gate && line.compare(0, 17, "##FILTER=<ID=PASS") == 0)
            out << "##FILTER=<ID=INDEL_QUAL_FILTERED,Description=\"Indel filtered due to QUAL below threshold (" << o.indel_quality << ")\">\n";
        return;
    }
    if (line.compare(0, 6, "#CHROM") == 0 || line.compare(0, 6, "#chrom") == 0) {
        if (!job.wrote_command_line) {
            if (!ps_defined) { out << "##FORMAT=<ID=PS,Number=1,Type=Integer,Description=\"Phase set identifier\">\n"; ps_defined = true; }
            out << "##longphaseVersion=" << lpsh::REFERENCE_VERSION << "\n##commandline=\"" << o.command << "\"\n";
            job.wrote_command_line = true;
        }
        out << line << "\n";
        return;
    }
    std::istringstream split(line);
    std::vector<std::string> f((std::istream_iterator<std::string>(split)), std::istream_iterator<std::string>());
    if (f.size() < 10) {                  // the reference indexes fields 8 and 9 unconditionally; short records are passed through
        if (f.empty()) return;
        for (size_t i = 0; i < f.size(); i++) out << (i ? "\t" : "") << f[i];
        out << "\n";
        return;
    }
    std::string &format = f[8], &sample = f[9];
    const int pos0 = std::stoi(f[1]) - 1;
    // an old PS key and its value go away
    const size_t ps_key = format.find("PS");
    if (ps_key != std::string::npos) {
        const int k = subfield_of(format, ps_key);
        if (format.find(':', ps_key + 1) != std::string::npos) format.erase(ps_key, 3);
        else if (ps_key > 0) format.erase(ps_key - 1, 3);
        const size_t at = subfield_start(sample, k);
        const size_t next = sample.find(':', at + 1);
        if (next != std::string::npos) sample.erase(at, next - at + 1);
        else if (at > 0) sample.erase(at - 1, sample.size() - at + 1);
    }
    // an old phased genotype becomes unphased, smaller allele first
    size_t gt_key = format.find("GT");
    if (gt_key != std::string::npos) {
        const size_t at = subfield_start(sample, subfield_of(format, gt_key));
        if (peek(sample, at + 1) == '|' && at + 2 < sample.size()) {
            if (sample[at] > sample[at + 2]) std::swap(sample[at], sample[at + 2]);
            sample[at + 1] = '/';
        }
    }
    const PhasedCall *call = nullptr;
    auto chr_phased = job.phased.find(f[0]);
    if (chr_phased != job.phased.end()) {
        auto hit = chr_phased->second.find(pos0);
        if (hit != chr_phased->second.end()) call = &hit->second;
    }
    auto chr_vars = job.variants.find(f[0]);
    const bool taken = chr_vars != job.variants.end() && chr_vars->second.count(pos0) != 0;
    format += ":PS";
    if (call && taken) {
        sample += ":" + std::to_string(call->block);
        gt_key = format.find("GT");
        const size_t at = subfield_start(sample, gt_key == std::string::npos ? 0 : subfield_of(format, gt_key));
        if (at + 2 < sample.size()) { sample[at] = call->hap_ref; sample[at + 1] = '|'; sample[at + 2] = call->hap_alt; }
    } else {
        sample += ":.";
    }
    if (quality_gate) {
        auto low = job.low_qual_indels.find(f[0]);
        if (low != job.low_qual_indels.end() && low->second.count(pos0)) f[6] = "INDEL_QUAL_FILTERED";
    }
    for (size_t i = 0; i < f.size(); i++) out << (i ? "\t" : "") << f[i];
    out << "\n";
}

int write_phase_vcf(lpsh_phase &job) {
    const PhaseOptions &o = job.opt;
    const std::string out_path = o.prefix + ".vcf";
    bool ps_defined = false;
    if (o.snp_file.find("gz") != std::string::npos) {
        // compressInput (ParsingBam.cpp:137-195): only newline-terminated lines are seen
        std::ofstream out(out_path.c_str());
        if (!out.is_open()) { std::cout << "Fail to open write file: " << out_path << "\n"; return 0; }
        std::string text;
        if (!lpsh::read_gz(o.snp_file, text)) { std::cout << "Fail to open vcf: " << o.snp_file << "\n"; return 0; }
        size_t at = 0;
        for (size_t nl; (nl = text.find('\n', at)) != std::string::npos; at = nl + 1)
            write_phase_line(job, text.substr(at, nl - at), ps_defined, out);
    } else if (o.snp_file.find("vcf") != std::string::npos) {
        // unCompressInput (:197-217): getline until eof, empty lines skipped
        std::ifstream in(o.snp_file.c_str());
        std::ofstream out(out_path.c_str());
        if (!out.is_open()) { std::cout << "Fail to open write file: " << out_path << "\n"; return 0; }
        if (!in.is_open()) { std::cout << "Fail to open vcf: " << o.snp_file << "\n"; return 0; }
        std::string line;
        while (!in.eof()) {
            std::getline(in, line);
            if (!line.empty()) write_phase_line(job, line, ps_defined, out);
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// the htslib loop of BamParser::direct_detect_alleles (ParsingBam.cpp:1243-1301): every record of chr:1-lastSNP of every
// -b file goes into the batch; the read filter (:1282-1291) and get_snp run on the device.
int pack_phase_contig(lpsh_phase &job, int i, htsThreadPool *pool, int readers = 1) {
    const std::string &chr = job.chr_names[(size_t)i];
    lpsh::PackedContig *pc = new lpsh::PackedContig();
    job.packed[(size_t)i] = pc;
    const auto &vars = job.variants[chr];
    for (const auto &kv : vars) pc->add_variant(kv.first, kv.second.ref, kv.second.alt);
    pc->ref = job.reference[chr];
    const int last = last_variant(job, chr);
    const std::string region = chr + ":1-" + std::to_string(last);
    for (const std::string &path : job.opt.bams) {
        samFile *in = hts_open(path.c_str(), "r");
        if (!in) return lpsh::fail("cannot open " + path);
        hts_set_fai_filename(in, job.opt.fasta.c_str());
        bam_hdr_t *hdr = sam_hdr_read(in);
        hts_idx_t *idx = hdr ? sam_index_load(in, path.c_str()) : NULL;
        if (!idx) {
            std::cout << "ERROR: Cannot open index for bam file\n";
            if (hdr) bam_hdr_destroy(hdr);
            sam_close(in);
            return lpsh::fail("cannot open the index of " + path);
        }
        hts_itr_t *it = sam_itr_querys(idx, hdr, region.c_str());
        if (pool && pool->pool) hts_set_opt(in, HTS_OPT_THREAD_POOL, pool);
        bam1_t *aln = bam_init1();
        if (it) {
            // LPS_GPU_INFLATE=1: the region's BGZF members are inflated in one batch on the device and parsed from memory
            // (SURVEY 8f rank 1); otherwise, or when that reader declines, htslib's reader as in the reference
            int done = 0;
            const char *gi = getenv("LPS_GPU_INFLATE");
            if (gi && gi[0] == '1') {
                const size_t before = pc->ref_start.size();
                done = lpsh::pack_region_inflated(path, it, *pc);
                if (done < 0) { hts_itr_destroy(it); bam_destroy1(aln); hts_idx_destroy(idx); bam_hdr_destroy(hdr); sam_close(in); return done; }
                if (done == 0) pc->truncate_reads(before);
            }
            // spare host threads (fewer contigs than -t): the region read by several readers on slices of it, same record sequence
            if (done == 0 && readers > 1) {
                done = lpsh::pack_region_split(path, job.opt.fasta, idx, it->tid, it->beg, it->end, readers, *pc);
                if (done < 0) { hts_itr_destroy(it); bam_destroy1(aln); hts_idx_destroy(idx); bam_hdr_destroy(hdr); sam_close(in); return done; }
            }
            if (done == 0) while (sam_itr_multi_next(in, it, aln) >= 0) pc->add_alignment(aln);
            hts_itr_destroy(it);
        }
        bam_destroy1(aln);
        hts_idx_destroy(idx);
        bam_hdr_destroy(hdr);
        sam_close(in);
    }
    pc->finish();
    return 0;
}

// threads left over when there are fewer contigs to phase than -t: they read slices of each contig's BAM region (LPS_READ_SPLIT overrides)
int readers_per_contig(const lpsh_phase &job) {
    if (const char *e = getenv("LPS_READ_SPLIT")) { const int v = atoi(e); if (v >= 1) return v; }
    int active = 0;
    for (const std::string &chr : job.chr_names) active += last_variant(job, chr) != -1;
    active = std::max(1, std::min(active, job.opt.threads));
    const int spare = job.opt.threads / active;
    // a split reader inflates inline, the single reader has the whole BGZF thread pool behind it: measured on 8 cores (507 MB, one contig)
    // 2.8 s unsplit, 5.1 s with 2 readers, 2.9 s with 4, 1.8 s with 8 - so only a wide split is taken
    return spare >= 4 ? spare : 1;
}

// Weight of every contig = mapped reads the index of the first BAM counts for it (hts_idx_get_stat: metadata of the .bai / .csi,
// nothing is decoded); without usable statistics, the position of its last variant.  order: contigs by descending weight.
void plan_contigs(const lpsh_phase &job, std::vector<int> &order, std::vector<double> &w) {
    const int n = (int)job.chr_names.size();
    w.assign((size_t)n, 0.0);
    for (int i = 0; i < n; i++) { const int lv = last_variant(job, job.chr_names[(size_t)i]); w[(size_t)i] = lv > 0 ? (double)lv : 0.0; }
    if (!job.opt.bams.empty()) {
        if (samFile *in = hts_open(job.opt.bams[0].c_str(), "r")) {
            hts_set_fai_filename(in, job.opt.fasta.c_str());
            bam_hdr_t *hdr = sam_hdr_read(in);
            hts_idx_t *idx = hdr ? sam_index_load(in, job.opt.bams[0].c_str()) : NULL;
            if (idx) {
                std::vector<double> reads((size_t)n, 0.0);
                bool all = true;
                for (int i = 0; i < n && all; i++) {
                    if (w[(size_t)i] == 0.0) continue;
                    const int tid = bam_name2id(hdr, job.chr_names[(size_t)i].c_str());
                    uint64_t mapped = 0, unmapped = 0;
                    if (tid < 0 || hts_idx_get_stat(idx, tid, &mapped, &unmapped) != 0) all = false;
                    else reads[(size_t)i] = (double)mapped + 1.0;
                }
                if (all) for (int i = 0; i < n; i++) if (w[(size_t)i] != 0.0) w[(size_t)i] = reads[(size_t)i];
                hts_idx_destroy(idx);
            }
            if (hdr) bam_hdr_destroy(hdr);
            sam_close(in);
        }
    }
    order.resize((size_t)n);
    for (int i = 0; i < n; i++) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w[(size_t)a] > w[(size_t)b]; });
}

// device_of: greedy longest-processing-time partition of the weights over the devices (called once the device count is known:
// the driver starts on a background thread while the first BAM regions are decoded).  LPS_PLACEMENT=roundrobin deals the
// contigs out in turn instead.
void assign_devices(const std::vector<int> &order, const std::vector<double> &w, int n_bins, std::vector<int> &device_of) {
    const int n = (int)order.size();
    if (n_bins < 1) n_bins = 1;
    device_of.assign((size_t)n, 0);
    const char *pl = getenv("LPS_PLACEMENT");
    if (pl && !strcmp(pl, "roundrobin")) { for (int k = 0; k < n; k++) device_of[(size_t)order[(size_t)k]] = k % n_bins; return; }
    std::vector<double> load((size_t)n_bins, 0.0);
    for (int k = 0; k < n; k++) {
        const int i = order[(size_t)k];
        int best = 0;
        for (int b = 1; b < n_bins; b++) if (load[(size_t)b] < load[(size_t)best]) best = b;
        device_of[(size_t)i] = best;
        load[(size_t)best] += w[(size_t)i];
    }
}

lps_phase_params device_params(const PhaseOptions &o, bool have_reference) {
    lps_phase_params p;
    memset(&p, 0, sizeof(p));
    p.mapping_quality = o.mapping_quality;
    p.is_ont = o.ont;
    p.have_reference = have_reference;
    p.connect_adjacent = o.connect_adjacent;
    p.base_quality = o.base_quality;
    p.distance = o.distance;
    p.edge_weight = o.edge_weight;
    p.edge_threshold = o.edge_threshold;
    p.overlap_threshold = o.overlap_threshold;
    p.read_confidence = o.read_confidence;
    p.snp_confidence = o.snp_confidence;
    return p;
}

}  // namespace

extern "C" {

int lpsh_phase_open(int argc, char **argv, lpsh_phase **out) {
    if (!out) return -1;
    *out = nullptr;
    lpsh_phase *job = new lpsh_phase();
    const int rc = parse_phase_options(argc, argv, job->opt);
    if (rc != 0) { delete job; return rc; }
    phase_banner(job->opt);
    if (job->opt.deepsomatic) {   // PhasingProcess.cpp:47-61: the rest of the run reads the preprocessed file
        std::time_t p0 = time(NULL);
        std::cerr << "preprocessing DeepSomatic VCF (filter GERMLINE, adjust GT by VAF) ... ";
        const std::string pre = job->opt.prefix + "_preprocessed.vcf";
        preprocess_deepsomatic(job->opt.snp_file, pre);
        std::cerr << difftime(time(NULL), p0) << "s\n";
        job->opt.snp_file = pre;
    }
    std::time_t t0 = time(NULL);
    std::cerr << "parsing VCF ... ";
    if (load_phase_vcf(*job) != 0) { delete job; return -1; }
    std::cerr << difftime(time(NULL), t0) << "s\n";
    t0 = time(NULL);
    std::cerr << "reading reference ... ";
    if (load_phase_reference(*job) != 0) { delete job; return -1; }
    std::cerr << difftime(time(NULL), t0) << "s\n";
    job->packed.assign(job->chr_names.size(), nullptr);
    *out = job;
    return 0;
}

int lpsh_phase_n_contigs(const lpsh_phase *h) { return h ? (int)h->chr_names.size() : 0; }
const char *lpsh_phase_contig_name(const lpsh_phase *h, int i) {
    return (h && i >= 0 && (size_t)i < h->chr_names.size()) ? h->chr_names[(size_t)i].c_str() : nullptr;
}
int lpsh_phase_last_variant(const lpsh_phase *h, int i) {
    return (h && i >= 0 && (size_t)i < h->chr_names.size()) ? last_variant(*h, h->chr_names[(size_t)i]) : -1;
}
int lpsh_phase_params(const lpsh_phase *h, lps_phase_params *out) {
    if (!h || !out) return -1;
    *out = device_params(h->opt, true);
    return 0;
}

int lpsh_phase_pack(lpsh_phase *h, int i, lpsh_packed *out) {
    if (!h || !out || i < 0 || (size_t)i >= h->chr_names.size()) return -1;
    if (last_variant(*h, h->chr_names[(size_t)i]) == -1) return lpsh::fail("contig has no variant");
    lpsh_phase_release(h, i);
    htsThreadPool pool = {NULL, 0};
    if (h->opt.threads > 1) pool.pool = hts_tpool_init(h->opt.threads);
    const int rc = pack_phase_contig(*h, i, &pool, readers_per_contig(*h));
    if (pool.pool) hts_tpool_destroy(pool.pool);
    if (rc != 0) return rc;
    h->packed[(size_t)i]->view(out);
    return 0;
}

void lpsh_phase_release(lpsh_phase *h, int i) {
    if (!h || i < 0 || (size_t)i >= h->packed.size()) return;
    delete h->packed[(size_t)i];
    h->packed[(size_t)i] = nullptr;
}

int lpsh_phase_set_result(lpsh_phase *h, int i, int32_t n_variants, const int32_t *ps, const int8_t *hap_ref) {
    if (!h || i < 0 || (size_t)i >= h->chr_names.size() || (n_variants && (!ps || !hap_ref))) return -1;
    const std::string &chr = h->chr_names[(size_t)i];
    const auto &vars = h->variants[chr];
    if ((size_t)n_variants != vars.size()) return lpsh::fail("result size does not match the contig's variant table");
    std::map<int, PhasedCall> &dst = h->phased[chr];
    dst.clear();
    int32_t k = 0;
    for (auto it = vars.begin(); it != vars.end(); ++it, ++k) {
        if (ps[k] == 0) continue;
        PhasedCall c;
        c.block = ps[k];
        c.hap_ref = (char)('0' + hap_ref[k]);
        c.hap_alt = (char)('0' + (1 - hap_ref[k]));
        dst[it->first] = c;
    }
    return 0;
}

int lpsh_phase_write_result(lpsh_phase *h) { return h ? write_phase_vcf(*h) : -1; }

int lpsh_phase_run(lpsh_phase *h) {
    if (!h) return -1;
    const PhaseOptions &o = h->opt;
    const int n = (int)h->chr_names.size();
    const int readers = readers_per_contig(*h);
    int n_dev = -1;   // counted when the first contig is packed: the driver starts (lpsh_phase_main) while the BAM is being decoded
    htsThreadPool pool = {NULL, 0};
    if (!(pool.pool = hts_tpool_init(o.threads))) fprintf(stderr, "Error creating thread pool\n");
    std::time_t t0 = time(NULL);
    int failed = 0;
    // every contig gets its slot in the shared maps BEFORE the threads start: operator[] on a std::map inserts, and concurrent
    // inserts from the contig threads would race on the tree (lpsh_phase_set_result, pack_phase_contig)
    for (const std::string &chr : h->chr_names) { h->phased[chr]; h->variants[chr]; h->reference[chr]; }
    // Placement: contigs are handed out heaviest first (so the largest one never starts last), and each is pinned to
    // the device a greedy LPT partition of the read counts gives it; the counts come from the BAM index, no record is read.
    std::vector<int> order, device_of((size_t)n, 0);
    std::vector<double> weight;
    plan_contigs(*h, order, weight);
#pragma omp parallel for schedule(dynamic, 1) num_threads(o.threads)
    for (int oi = 0; oi < n; oi++) {
        const int i = order[(size_t)oi];
        const std::string &chr = h->chr_names[(size_t)i];
        std::time_t c0 = time(NULL);
        if (last_variant(*h, chr) == -1) continue;
        if (pack_phase_contig(*h, i, &pool, readers) != 0) {
#pragma omp critical
            failed = 1;
            continue;
        }
        lpsh::PackedContig *pc = h->packed[(size_t)i];
#pragma omp critical(lpsh_device_count)
        if (n_dev < 0) { n_dev = lpsh::device_count(); assign_devices(order, weight, n_dev, device_of); }
        if (n_dev < 1) {
#pragma omp critical
            { lpsh::fail("no usable CUDA device (there is no CPU fallback)"); failed = 1; }
            lpsh_phase_release(h, i);
            continue;
        }
        if (pc->n_reads() > 0) {
            lps_ctx *ctx = nullptr;
            lpsh_packed v;
            pc->view(&v);
            const lps_phase_params p = device_params(o, !pc->ref.empty());
            lps_phase_result r;
            int rc = lps_ctx_create(device_of[(size_t)i], &ctx);
            if (rc == 0) rc = lps_contig_set_reference(ctx, v.ref, v.ref_len);
            if (rc == 0) rc = lps_contig_set_variants(ctx, &v.variants, o.ont);
            if (rc == 0) rc = lps_batch_submit(ctx, &v.batch);
            if (rc == 0) rc = lps_phase_contig(ctx, &p, &r);
            if (rc == LPS_E_CIGAR) { std::cerr << "alignment find unsupported CIGAR operation from read\n"; exit(1); }   // ParsingBam.cpp:1625-1628
            if (rc == 0) lpsh_phase_set_result(h, i, r.n_variants, r.ps, r.hap_ref);
            else {
#pragma omp critical
                { lpsh::fail(std::string("contig ") + chr + ": " + (ctx ? lps_last_error(ctx) : "lps_ctx_create failed")); failed = 1; }
            }
            if (ctx) lps_ctx_destroy(ctx);
        }
        lpsh_phase_release(h, i);
#pragma omp critical
        std::cerr << "(" << chr << "," << difftime(time(NULL), c0) << "s)";
    }
    hts_tpool_destroy(pool.pool);
    std::cerr << "\nparsing total:  " << difftime(time(NULL), t0) << "s\n";
    return failed ? -1 : 0;
}

void lpsh_phase_close(lpsh_phase *h) {
    if (!h) return;
    for (size_t i = 0; i < h->packed.size(); i++) delete h->packed[i];
    delete h;
}

int lpsh_phase_main(int argc, char **argv) {
    std::time_t t0 = time(NULL);
    std::thread warm = lpsh::warm_up_device();   // the driver starts while the VCF, the FASTA and the first BAM regions are read
    struct Join { std::thread &t; ~Join() { if (t.joinable()) t.join(); } } join_warm{warm};
    lpsh_phase *job = nullptr;
    const double m0 = lpsh::now_ms();
    const int rc = lpsh_phase_open(argc, argv, &job);
    if (rc == 2) return 0;
    if (rc != 0) { if (rc < 0) std::cerr << "phase: " << lpsh_last_error() << "\n"; return 1; }
    const double m1 = lpsh::now_ms();
    if (lpsh_phase_run(job) != 0) { std::cerr << "phase: " << lpsh_last_error() << "\n"; lpsh_phase_close(job); return 1; }
    std::time_t t1 = time(NULL);
    std::cerr << "writeResult SNP ... ";
    lpsh_phase_write_result(job);
    std::cerr << difftime(time(NULL), t1) << "s\n";
    std::cerr << "\ntotal process: " << difftime(time(NULL), t0) << "s\n";
    std::cerr << "[timing] load " << (m1 - m0) << " ms, contig loop " << (lpsh::now_ms() - m1) << " ms (incl. writing)\n";
    lpsh_phase_close(job);
    return 0;
}

}  // extern "C"
