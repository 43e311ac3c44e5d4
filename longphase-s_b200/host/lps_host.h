/*
 * lps_host.h — the C++ host that sits ABOVE the C ABI of include/lps.h: the reference's own command-line and file surface
 * (`longphase-s phase | haplotag | somatic_haplotag`) kept as it is, with the per-contig hot path handed to liblps_b200.so.
 *
 * The host keeps htslib for BAM / VCF / FASTA decoding, exactly as BASELINE.json:north_star asks: it packs the decoded
 * alignments into the SoA batch of include/lps.h, calls the kernels through the C ABI, and writes the reference's output
 * files (phased VCF, tagged BAM, --log table) byte for byte.  htslib is compiled from the sources vendored with the
 * reference (longphase-s_b200/host/Makefile), never copied into this repository.
 *
 * The entry points below are the stages of that host, exported with C linkage so that the CPU test-suite can drive every
 * stage that does not need a GPU (option parsing, VCF / FASTA / BAM loading and packing, result writers) and compare the
 * written files with those of the unmodified reference binary.  Only lpsh_*_run touches the GPU.
 *
 * Reference seams replaced (file:line relative to the reference tree):
 *   PhasingOptions / PhasingMain            src/phase/Phasing.cpp:118-372
 *   PhasingProcess::PhasingProcess          src/phase/PhasingProcess.cpp:5-208
 *   SnpParser::SnpParser / writeResult      src/phase/ParsingBam.cpp:219-359, 444-635
 *   FastaParser::FastaParser                src/phase/ParsingBam.cpp:17-59
 *   BamParser::direct_detect_alleles        src/phase/ParsingBam.cpp:1243-1301  (the htslib loop; get_snp is on the device)
 *   HaplotagOptions / HaplotagMain          src/haplotag/Haplotag.cpp:74-189
 *   HaplotagProcess / HaplotagBamParser     src/haplotag/HaplotagProcess.cpp:20-262, src/haplotag/HaplotagParsingBam.cpp:176-492
 *   VcfParser (germline)                    src/haplotag/HaplotagVcfParser.cpp:10-470
 */
#ifndef LPS_HOST_H
#define LPS_HOST_H

#include <stdint.h>
#include "lps.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- phase ------------------------------------------------------------------------------------------------------ */
typedef struct lpsh_phase lpsh_phase;

/* one contig as the device sees it; every pointer is owned by the job and valid until lpsh_phase_release / close    */
typedef struct {
    lps_variants variants;      /* het variants of the contig in ascending position (SnpParser::chrVariant)          */
    lps_read_batch batch;       /* all alignments of the region chr:1-lastSNP of every -b file, in file order        */
    const char *ref;            /* FastaParser::chrString[chr] (0 .. lastSNP+5)                                       */
    int64_t ref_len;
    const char *names;          /* read names, NUL terminated, concatenated                                          */
    const uint64_t *name_off;   /* [n_reads] offset of read r's name in names                                        */
} lpsh_packed;

/* argv[0] is the sub-command word ("phase"), as PhasingMain receives it (src/main.cpp:41).  Parses the options with the
 * reference's names / defaults / checks, prints its banner on stderr, loads the VCF and the FASTA.  Returns 0; 1 when the
 * options are invalid (usage printed, the reference exits with EXIT_FAILURE); 2 for --help.                          */
int lpsh_phase_open(int argc, char **argv, lpsh_phase **out);
int lpsh_phase_n_contigs(const lpsh_phase *h);                       /* SnpParser::getChrVec()                      */
const char *lpsh_phase_contig_name(const lpsh_phase *h, int i);
int lpsh_phase_last_variant(const lpsh_phase *h, int i);             /* SnpParser::getLastSNP, -1 = none            */
int lpsh_phase_params(const lpsh_phase *h, lps_phase_params *out);
/* decodes the contig's alignments with htslib and packs them; 0, or a negative value when a BAM / index cannot be opened */
int lpsh_phase_pack(lpsh_phase *h, int i, lpsh_packed *out);
void lpsh_phase_release(lpsh_phase *h, int i);
/* VairiantGraph::exportResult for contig i from the arrays of lps_phase_result (PhasingGraph.cpp:1049-1077)           */
int lpsh_phase_set_result(lpsh_phase *h, int i, int32_t n_variants, const int32_t *ps, const int8_t *hap_ref);
/* SnpParser::writeResult: <prefix>.vcf (ParsingBam.cpp:444-635)                                                      */
int lpsh_phase_write_result(lpsh_phase *h);
/* the contig loop of PhasingProcess (PhasingProcess.cpp:113-173) on the GPU(s): pack, lps_phase_contig, set_result for
 * every contig with -t host threads, one lps_ctx each, devices taken round-robin.  Prints the reference's progress on stderr. */
int lpsh_phase_run(lpsh_phase *h);
void lpsh_phase_close(lpsh_phase *h);
/* everything: what `longphase-s phase ...` does.  Returns the process exit code.                                       */
int lpsh_phase_main(int argc, char **argv);

/* ---- haplotag --------------------------------------------------------------------------------------------------- */
typedef struct lpsh_tag lpsh_tag;

int lpsh_tag_open(int argc, char **argv, lpsh_tag **out);            /* argv[0] = "haplotag"                          */
int lpsh_tag_n_contigs(const lpsh_tag *h);
const char *lpsh_tag_contig_name(const lpsh_tag *h, int i);
int lpsh_tag_params(const lpsh_tag *h, lps_tag_params *out);
/* opens input and output, writes the header (HaplotagParsingBam.cpp:20-66); must precede the first lpsh_tag_pack       */
int lpsh_tag_begin(lpsh_tag *h);
/* reads the NEXT chunk of contig i's alignments (region aware; LPS_TAG_CHUNK records, default 65536) and packs it with the
 * contig's phased variants; the records stay buffered for lpsh_tag_emit.  1 = a chunk is ready, 0 = contig exhausted, < 0 error */
int lpsh_tag_pack(lpsh_tag *h, int i, lpsh_packed *out);
/* applies HP / PS / PQ to the buffered records of the chunk and writes them in their original order
 * (HaplotagProcess.cpp:318-438), adds the chunk's share of the statistics and of the --log table                         */
int lpsh_tag_emit(lpsh_tag *h, int i, const lps_tag_result *r);
int lpsh_tag_end(lpsh_tag *h);                                         /* closes the files, prints the report            */
/* the verdicts of one packed chunk (the device in the real program; the test-suite passes its CPU checker): fills *out, 0 or != 0 */
typedef int (*lpsh_tag_judge_fn)(void *user, int contig, const lpsh_packed *chunk, int want_calls, lps_tag_result *out);
/* the whole tagging pass: a reader thread parses and packs chunk after chunk while the calling thread judges the previous one,
 * tags its records and writes them in their original order (begin ... end included)                                          */
int lpsh_tag_run_with(lpsh_tag *h, lpsh_tag_judge_fn judge, void *user);
int lpsh_tag_run(lpsh_tag *h);                                         /* lpsh_tag_run_with(lps_tag_reads on device 0)      */
void lpsh_tag_close(lpsh_tag *h);
int lpsh_tag_main(int argc, char **argv);

/* ---- somatic_haplotag ---------------------------------------------------------------------------------------------- *
 * SomaticHaplotagProcess::pipelineProcess (src/somatic_haplotag/SomaticHaplotagProcess.cpp:51-103): NORMAL + TUMOR VCFs into one
 * union map, extract pass over the normal BAM and over the tumor BAM, purity (<prefix>_purity.out), calling, tagging of the
 * tumor BAM (HP:Z, PS:i unless none, PQ:i).                                                                                */
typedef struct lpsh_som lpsh_som;
int lpsh_som_open(int argc, char **argv, lpsh_som **out);            /* argv[0] = "somatic_haplotag" or "estimate_purity" */
int lpsh_som_n_contigs(const lpsh_som *h);
const char *lpsh_som_contig_name(const lpsh_som *h, int i);
int lpsh_som_params(const lpsh_som *h, int pass, lps_tag_params *out); /* pass 0: extract passes, 1: tagging pass         */
/* every alignment of contig i of the NORMAL (which = 0) or TUMOR (1) BAM, packed with the contig's union map; the arrays stay
 * valid until the next lpsh_som_pack                                                                                      */
int lpsh_som_pack(lpsh_som *h, int i, int which, lpsh_packed *out, lps_tumor_variants *tv);
/* keeps a copy of the result of lps_extract_normal (which = 0) / lps_extract_tumor (1) for contig i                       */
int lpsh_som_set_extract(lpsh_som *h, int i, int which, const lps_extract_result *r);
/* runTumorPurityEstimator (or --tumor-purity), <prefix>_purity.out, the calling stage and getSomaticFlag for every contig
 * (SomaticVarCaller.cpp:816-866, 937-949, 2397-2412); needs both extract results of every contig                          */
int lpsh_som_call(lpsh_som *h);
/* the purity stage alone (what `estimate_purity` runs after the two extract passes; lpsh_som_call includes it)              */
int lpsh_som_estimate(lpsh_som *h);
double lpsh_som_purity(const lpsh_som *h);
int64_t lpsh_som_n_somatic(const lpsh_som *h);
int lpsh_som_tag_begin(lpsh_som *h);
/* next chunk of contig i of the tumor BAM, the union map now carrying isSomaticVariant / somaticReadDeriveByHP: 1, 0 or < 0 */
int lpsh_som_tag_pack(lpsh_som *h, int i, lpsh_packed *out, lps_tumor_variants *tv);
int lpsh_som_tag_emit(lpsh_som *h, int i, const lps_somatic_tag_result *r);
int lpsh_som_tag_end(lpsh_som *h);
typedef int (*lpsh_som_judge_fn)(void *user, int contig, const lpsh_packed *chunk, const lps_tumor_variants *tv, lps_somatic_tag_result *out);
/* the whole tagging pass, pipelined like lpsh_tag_run_with (begin ... end included); needs lpsh_som_call                     */
int lpsh_som_tag_run_with(lpsh_som *h, lpsh_som_judge_fn judge, void *user);
int lpsh_som_run(lpsh_som *h);
void lpsh_som_close(lpsh_som *h);
int lpsh_som_main(int argc, char **argv);

/* ---- tooling: the generator's SoA batches written as BAM + BAI with htslib (bench and tests; no samtools in the image) --- */
typedef struct lpsh_bamw lpsh_bamw;
lpsh_bamw *lpsh_bamw_open(const char *path, int n_contigs, const char **names, const int64_t *lens, int threads);
/* appends the batch (coordinate order) as records of contig `tid`; names = fixed-stride NUL-terminated read names      */
/* sizeof(lps_read_batch) this library was compiled with: a host built against an older include/lps.h would hand the kernels'
 * library a shorter struct (tests/test_abi.py compares it with the ctypes mirror and liblps_b200.so's own header hash)      */
int lpsh_sizeof_read_batch(void);
int lpsh_bamw_append(lpsh_bamw *w, int tid, const lps_read_batch *b, const char *names, int name_stride);
int64_t lpsh_decode_only(const char *in_path, int threads);                        /* sam_read1 over the whole file: the I/O floor   */
int lpsh_to_sam(const char *in_path, const char *fasta, const char *out_path);   /* BAM / CRAM -> SAM text                          */
int lpsh_bamw_close(lpsh_bamw *w);                                     /* closes and builds <path>.bai                   */

/* ---- BGZF inflation hook ---------------------------------------------------------------------------------------------- *
 * With LPS_GPU_INFLATE=1 `phase` reads a contig's BAM region as compressed bytes and inflates them in one batch; the inflater is
 * lps_bgzf_inflate on device 0 unless a hook is installed (the CPU test-suite installs zlib to check the reader around it).     */
typedef int (*lpsh_inflate_fn)(void *user, const uint8_t *data, uint64_t n_bytes, const lps_bgzf_block *blocks, uint64_t n_blocks,
                               uint8_t *out, uint64_t out_cap);
void lpsh_set_inflater(lpsh_inflate_fn fn, void *user);

/* With LPS_GPU_DEFLATE=1 the tagging passes write their BAM through a batched deflater: lps_bgzf_deflate on device 0, or this
 * hook (tests without a GPU hand in the host-compiled member encoder; the signature is lps_bgzf_deflate's minus the context). */
typedef int (*lpsh_deflate_fn)(void *user, const uint8_t *in, uint64_t in_len, uint32_t block_bytes, uint8_t *out, uint64_t out_cap,
                               uint64_t *out_len);
void lpsh_set_deflater(lpsh_deflate_fn fn, void *user);

const char *lpsh_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
