/*
 * lps.h — C ABI of the B200-native read-to-variant hot path of LongPhase-S.
 *
 * This is the drop-in boundary.  The reference has no FFI; its seams are
 * C++ member functions inside one binary.  Each entry point below names the reference
 * interface it replaces (file:line relative to the reference tree).  All functions return
 * LPS_OK (0) or a negative lps_status; none of them ever calls exit().  lps_last_error()
 * returns a human readable message for the last failure on that context.
 *
 * Threading: one context per (GPU, host thread).  A context is thread-compatible, not
 * thread-safe.  All pointers in the argument structs are HOST pointers (pageable or pinned)
 * unless the struct says otherwise; the library copies what it needs before returning unless
 * the function is documented as asynchronous.
 *
 * There is no CPU fallback: every compute entry point runs hand-written sm_100a kernels and
 * fails with LPS_E_CUDA when no device is usable.
 */
#ifndef LPS_H
#define LPS_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lps_ctx lps_ctx;

typedef enum {
    LPS_OK = 0,
    LPS_E_ARG = -1,       /* bad argument (null pointer, unsorted positions, ...)            */
    LPS_E_CUDA = -2,      /* CUDA runtime failure; message has the cudaError string          */
    LPS_E_STATE = -3,     /* call order violated (e.g. build_edges before call_alleles)      */
    LPS_E_CIGAR = -4,     /* unsupported CIGAR op (reference: exit(1), ParsingBam.cpp:1625)  */
    LPS_E_NOMEM = -5,
    LPS_E_DATA = -6       /* malformed input data (BGZF header, deflate stream, CRC)         */
} lps_status;

/* ---- variant table of one contig ------------------------------------------------------- *
 * Replaces the per-contig std::map<int,RefAlt> copy made in BamParser::BamParser
 * (src/phase/ParsingBam.cpp:1207-1235) and the std::map<int,MultiGenomeVar> of the tag family
 * (src/haplotag/HaplotagType.h:110-162).  SoA, strictly ascending by pos (0-based).        */
typedef struct {
    int32_t n;
    const int32_t *pos;       /* 0-based position (rec->pos)                                  */
    const uint8_t *ref0;      /* first character of REF (ASCII, case preserved)               */
    const uint8_t *alt0;      /* first character of ALT                                       */
    const uint16_t *ref_len;  /* strlen(REF), saturated at 65535                              */
    const uint16_t *alt_len;  /* strlen(ALT)                                                  */
    /* --- tag-family only; may be NULL for `phase` ------------------------------------------ */
    const uint8_t *hp1_is_alt; /* 1 when HP1 carries ALT (GT 1|0), 0 when HP1 carries REF (0|1) */
    const int32_t *ps;         /* phase-set id of the NORMAL record (PS), 0 = none              */
    const uint8_t *gt_kind;    /* GenomeType: 1 PHASED_HETERO, 2 UNPHASED_HETERO, 3 UNPHASED_HOMO */
} lps_variants;

/* ---- a batch of decoded alignments of ONE contig, in BAM (coordinate) order ------------- *
 * Replaces the bam1_t handed to BamParser::get_snp (src/phase/ParsingBam.cpp:1303) and to
 * ChromosomeProcessor::processRead (src/haplotag/HaplotagParsingBam.h:360-367).            */
typedef struct {
    int32_t n_reads;
    const int32_t *ref_start;   /* core.pos                                                   */
    const int32_t *l_qseq;      /* core.l_qseq                                                */
    const uint32_t *n_cigar;    /* core.n_cigar                                               */
    const uint64_t *cigar_off;  /* first op of read r in cigar[], in uint32 units             */
    const uint64_t *seq_off;    /* first byte of read r in seq4[] (BAM 4-bit packing, 2/byte) */
    const uint64_t *qual_off;   /* first byte of read r in qual[]                             */
    const uint16_t *flag;       /* core.flag                                                  */
    const uint8_t *mapq;        /* core.qual                                                  */
    const int32_t *name_rank;   /* rank of the read NAME in lexicographic (std::string <) order
                                   among the names of this batch; alignments sharing a name share
                                   a rank.  Needed because the reference folds float edge weights
                                   in std::map<std::string,...> order (PhasingGraph.cpp:697,848) */
    const uint32_t *cigar;      /* BAM encoding len<<4|op                                     */
    uint64_t cigar_len;         /* total uint32 in cigar[]                                    */
    const uint8_t *seq4;
    uint64_t seq_bytes;
    const uint8_t *qual;
    uint64_t qual_bytes;
    /* --- the CIGAR stream in 16 bits per op: what the kernels read -------------------------------------------------- *
     * When cigar16 != NULL, cigar[] is ignored (may be NULL): cigar16[i] holds op i of the same stream in 16 bits, len<<4|op
     * for len < 4095, and 0xFFF0|op for a longer op, whose true length is listed in cigar_long_len[] / cigar_long_at[] (index
     * into cigar16[], strictly ascending).  lps_pack_cigar16 produces all three from BAM's uint32 ops while the host appends a
     * record.  This IS the device-resident format (2 bytes per op in HBM and on the wire); a batch submitted with uint32 ops is
     * narrowed to it on the device on arrival (k_narrow_cigar32).  With lps_batch_submit_device the three arrays are device
     * pointers and are used where they lie: cigar16 must be 16-byte aligned (the walking kernel brings super-chunks of it into
     * shared memory with 16-byte granular bulk copies; it never reads past the 16-byte unit that holds the last op).          */
    const uint16_t *cigar16;        /* [cigar_len]                                              */
    const uint32_t *cigar_long_len; /* [n_cigar_long]                                           */
    const uint64_t *cigar_long_at;  /* [n_cigar_long]                                           */
    uint64_t n_cigar_long;
    /* --- optional wire format of the CIGAR stream in 8 bits per op (lps_batch_submit only) ------------------------------ *
     * The CIGAR stream is what a batch sends over PCIe once SEQ / QUAL stay on the host, and long-read CIGARs are almost
     * only short M / I / D ops.  When cigar8 != NULL, cigar[] and cigar16[] are ignored: cigar8[i] is
     *     0x00..0x7F  M of 1..128 bases      0x80..0xB7  I of 1..56 bases      0xB8..0xEF  D of 1..56 bases
     *     0xFF        any other op: the next unused entry of cigar_esc16[] (the op in the 16-bit encoding above; it may in
     *                 turn be a 0xFFF escape into cigar_long_len / cigar_long_at, whose indices are op indices as before)
     * cigar_esc_blk[k] = number of 0xFF bytes in cigar8[0 .. 256 k), k = 0 .. ceil(cigar_len / 256), so the device can expand
     * the stream in parallel (k_expand_cigar8 writes the 16-bit stream the kernels read; results cannot differ).
     * lps_pack_cigar8 produces all of it while the host appends records.                                              */
    const uint8_t *cigar8;          /* [cigar_len]                                              */
    const uint16_t *cigar_esc16;    /* [n_cigar_esc]                                            */
    uint64_t n_cigar_esc;
    const uint32_t *cigar_esc_blk;  /* [cigar_len / 256 + 2]                                    */
    /* --- optional wire format of SEQ + QUAL in ONE stream (phase calls only) ------------------------------------------ *
     * BamParser::get_snp looks at bam_seqi(qstring, i) and bam_get_qual(aln)[i] of the SAME query index i, for ~1 index per
     * 1000 bases (ParsingBam.cpp:1398-1425).  With seq4[] and qual[] in two arrays every allele call costs two scattered
     * sectors - two PCIe read requests when the buffers stay in pinned host memory.  When sq != NULL, seq4[] / qual[] /
     * qual_off[] are ignored (may be NULL) and seq_off[r] is the byte offset of read r in sq[] (a multiple of 16):
     * the read is a row of 16-byte units, unit u covering query indices 10 u .. 10 u + 9:
     *     bytes 0..9   bam_get_qual(aln)[10 u + k], k = 0..9            (0 past l_qseq)
     *     bytes 10..14 bam_get_seq(aln) nibbles of the same ten bases, BAM's packing (even k in the high nibble)
     *     byte  15     0
     * so one aligned 16-byte load serves a call.  lps_pack_sq writes a read's row while the host appends the record;
     * lps_sq_row_bytes gives its length.  Only lps_phase_* accept such a batch (the tag dialects and the window diff read
     * runs of bases, not single ones: LPS_E_STATE).  Pinned sq stays on the host like pinned seq4 / qual.               */
    const uint8_t *sq;              /* [sq_bytes]                                               */
    uint64_t sq_bytes;
} lps_read_batch;

/* one allele call: replaces struct Variant (src/shared/Util.h:63-75)                        */
typedef struct {
    int32_t var;      /* index into the contig's variant table                                */
    int16_t quality;  /* base quality, or -4 indel / -5 danger indel (ParsingBam.cpp:1485-1490) */
    int8_t allele;    /* 0 REF, 1 ALT                                                          */
    int8_t origin;    /* 0 = called inside an M/=/X op, 1 = called by the D-op rule            */
} lps_call;

typedef struct {
    int32_t mapping_quality;   /* -q, default 1 (Phasing.cpp:105)                              */
    int32_t is_ont;            /* --ont: enables SnpParser::filterSNP (ParsingBam.cpp:837-912) */
    int32_t have_reference;    /* ref_string != "" (ParsingBam.cpp:1544)                       */
    int32_t connect_adjacent;  /* -a, default 35                                               */
    int32_t base_quality;      /* -x? baseQuality default 12                                   */
    int32_t distance;          /* -d default 300000                                            */
    double edge_weight;        /* default 0.1                                                  */
    double edge_threshold;     /* default 0.7                                                  */
    double overlap_threshold;  /* default 0.2                                                  */
    double read_confidence;    /* default 0.65                                                 */
    double snp_confidence;     /* default 0.75                                                 */
} lps_phase_params;

/* read status codes written to lps_calls.read_status                                         */
enum { LPS_READ_OK = 0, LPS_READ_ABORTED = 1, LPS_READ_FILTERED = 2 };

/* result of lps_phase_call_alleles; arrays are owned by the context and stay valid until the
 * next lps_batch_submit / lps_ctx_destroy.                                                   */
typedef struct {
    int32_t n_reads;
    uint64_t n_calls;
    const uint64_t *call_off;     /* [n_reads+1] CSR offsets into calls[]                      */
    const lps_call *calls;        /* calls of read r = calls[call_off[r] .. call_off[r+1])     */
    const uint8_t *read_status;   /* [n_reads]                                                 */
    int32_t n_clips;              /* distinct clip positions                                   */
    const int32_t *clip_pos;      /* ascending                                                 */
    const int32_t *clip_front;    /* clipCount[pos][FRONT]                                     */
    const int32_t *clip_back;     /* clipCount[pos][BACK]                                      */
} lps_calls;

/* per-variant annotation computed on the device from the reference string                    */
typedef struct {
    int32_t n;
    const uint8_t *homopolymer;   /* homopolymerLength(pos) (Util.cpp:21-54), 1..10+            */
    const uint8_t *is_danger;     /* tandem-repeat indel flag (ParsingBam.cpp:378-417)         */
    const uint8_t *filtered;      /* erased by SnpParser::filterSNP (ONT only)                 */
} lps_variant_notes;

/* result of lps_phase_build_edges                                                             */
typedef struct {
    int32_t n_nodes;              /* variants with >=1 surviving call (totalVariantInfo keys)  */
    const int32_t *node_var;      /* [n_nodes] variant index of node k (ascending)             */
    const uint8_t *node_type;     /* [n_nodes] 0 SNP, 3 indel, 4 danger indel                  */
    int32_t window;               /* = connect_adjacent                                        */
    const float *weights;         /* [n_nodes][window][4] rr, ra, ar, aa between node k and k+1+d */
    uint64_t n_contrib;           /* pair contributions folded into the table                  */
    uint64_t n_contrib_far;       /* contributions to pairs farther than `window` nodes apart:
                                     written by the reference but never read by its sweep      */
} lps_edges;

/* result of the whole per-contig phase pipeline                                               */
typedef struct {
    int32_t n_variants;
    const int32_t *ps;            /* [n_variants] phase set (block start + 1) or 0 = unphased   */
    const int8_t *hap_ref;        /* [n_variants] haplotype of the REF allele (0/1), -1 unphased */
    int32_t n_reads;
    const int8_t *read_hp;        /* [n_reads] 0/1 haplotype, -1 untagged, -2 read not used     */
    const int32_t *hp_counts;     /* [n_variants][4] hp0_ref, hp0_alt, hp1_ref, hp1_alt         */
    const int32_t *ps_sweep;      /* [n_variants] phase set after edgeConnectResult, before readCorrection */
    const int8_t *hap_ref_sweep;  /* [n_variants] REF-allele haplotype after the sweep, -1 outside blocks  */
} lps_phase_result;

/* ---- BGZF inflation ------------------------------------------------------ *
 * Replaces bgzf_read_block -> inflate_block -> bgzf_uncompress (htslib/bgzf.c:988-1200, 792-812, 744-785; zlib inflate with
 * windowBits -15) for a batch of BGZF blocks: 77 % of the reference's `phase` wall time.  One block per table entry.       */
typedef struct {
    uint64_t comp_off;   /* first byte of the raw deflate stream (18 bytes into the member)     */
    uint32_t comp_len;   /* its length (member length - 18 - 8)                                  */
    uint32_t out_len;    /* ISIZE                                                                */
    uint64_t out_off;    /* where the block's bytes go in the output (sum of the earlier ISIZEs)  */
    uint32_t crc32;      /* CRC32 field of the trailer                                           */
    uint32_t reserved_;
} lps_bgzf_block;
/* Walks the members of a BGZF byte range like bgzf_read_block does, with htslib's header check (check_header, bgzf.c:876-883).
 * blocks may be NULL to count only.  Returns LPS_E_DATA for a malformed or truncated member, LPS_E_ARG when cap is too small
 * (n_blocks / out_bytes are still set).  Pure host code.                                                                     */
int lps_bgzf_scan(const uint8_t *data, uint64_t n_bytes, lps_bgzf_block *blocks, uint64_t cap, uint64_t *n_blocks, uint64_t *out_bytes);
/* Inflates the blocks on the device: host buffers in, host buffer out (copies inside).  check_crc != 0 also verifies every
 * block's CRC32 on the host, as htslib does (bgzf.c:777-781).  LPS_E_DATA names the first bad block in lps_last_error.      */
int lps_bgzf_inflate(lps_ctx *ctx, const uint8_t *data, uint64_t n_bytes, const lps_bgzf_block *blocks, uint64_t n_blocks, uint8_t *out,
                     uint64_t out_cap, int check_crc);
/* Same with DEVICE pointers throughout (data, block table, output); no CRC check.  lps_stats.ms_kernel_bgzf has the kernel time. */
int lps_bgzf_inflate_device(lps_ctx *ctx, const uint8_t *d_data, const lps_bgzf_block *d_blocks, uint64_t n_blocks, uint8_t *d_out);

/* ---- BGZF deflation (the writer side) --------------------------------------- *
 * Replaces bgzf_write -> bgzf_flush -> deflate_block -> bgzf_compress (htslib/bgzf.c:553-612, 697, 1440-1500; zlib deflate with
 * windowBits -15) for a batch of blocks.  The input is cut into pieces of block_bytes (1 .. 65280 = htslib's BGZF_BLOCK_SIZE; the
 * last piece may be shorter) and every piece becomes one complete BGZF member (header with the BC field, deflate stream, CRC32,
 * ISIZE); out receives the members back to back, *out_len their total size.  No EOF marker is appended.  Each member is ONE
 * dynamic-Huffman block of literals (no string matching: bases and qualities, ~95 % of a long-read BAM, hold no repeats; measured
 * 0.631 of the input against 0.617 for zlib's default level), or a stored block when that is not smaller; any inflater reads it and
 * the BAM records are those htslib would have written - the compressed bytes are not.  out_cap >= lps_bgzf_deflate_bound().    */
uint64_t lps_bgzf_deflate_bound(uint64_t in_len, uint32_t block_bytes);
int lps_bgzf_deflate(lps_ctx *ctx, const uint8_t *in, uint64_t in_len, uint32_t block_bytes, uint8_t *out, uint64_t out_cap, uint64_t *out_len);
/* The member encoder the kernel runs (one __host__ __device__ function), compiled for the host: tests check it against zlib's
 * inflate here and compare the kernel's bytes with it on the GPU box.  n <= 65280, out_cap >= 65312.  Not a fallback: nothing in
 * this library or the host calls it.                                                                                          */
int lps_bgzf_deflate_block_host(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t out_cap, uint32_t *out_len);

/* ---- process-wide ------------------------------------------------------------------------ */
/* How host threads wait for the device on `device`: 0 = spin (lowest latency, one core per waiting thread), 1 = block on an
 * interrupt (cudaDeviceScheduleBlockingSync), 2 = spin but yield the core between polls (cudaDeviceScheduleYield): the mode
 * for more host threads than cores.  Must be called before the first context on the device is created.                   */
int lps_set_blocking_sync(int device, int on);

/* ---- lifetime ---------------------------------------------------------------------------- */
int lps_ctx_create(int device, lps_ctx **out);
void lps_ctx_destroy(lps_ctx *ctx);
const char *lps_last_error(const lps_ctx *ctx);
const char *lps_version(void);

/* ---- per-contig static data -------------------------------------------------------------- */
/* FastaParser::chrString (ParsingBam.cpp:17-59) as used by homopolymerLength (Util.cpp:21-54),
 * getVariants_markindel (ParsingBam.cpp:378-417) and getWindowsDiffRef (SomaticVarCaller.cpp:627).
 * The string is copied before the call returns (the caller may free it at once).                   */
int lps_contig_set_reference(lps_ctx *ctx, const char *ref_ascii, int64_t len);
/* BamParser::BamParser variant copy + mark-indel (ParsingBam.cpp:1207-1235, 378-417) and, when
 * is_ont, the variant side of SnpParser::filterSNP (ParsingBam.cpp:866-888).                  */
int lps_contig_set_variants(lps_ctx *ctx, const lps_variants *v, int is_ont);
int lps_contig_get_notes(lps_ctx *ctx, lps_variant_notes *out);

/* ---- reads -------------------------------------------------------------------------------- */
/* Copies the batch to the device and waits for the copies.  The offsets are validated first (cigar_off + n_cigar within
 * cigar_len, seq_off / qual_off + the read's bytes within seq_bytes / qual_bytes, l_qseq >= 0): LPS_E_ARG otherwise.
 * Lifetime: the per-read arrays and the CIGAR stream may be freed as soon as the call returns.  seq4 / qual are different when
 * they are PINNED host memory (cudaHostAlloc / cudaHostRegister): then they are NOT copied - the kernels gather the few sectors
 * they need straight from the host buffers (zero-copy) - and must stay valid and unchanged until the next lps_batch_submit /
 * lps_batch_submit_device on this context, or lps_ctx_destroy.  Pageable seq4 / qual are copied like everything else.        */
int lps_batch_submit(lps_ctx *ctx, const lps_read_batch *b);
/* Appends n BAM CIGAR ops (bam_get_cigar(aln), core.n_cigar) to a compact stream: out16[0..n) receives the 16-bit ops;
 * an op of length >= 4095 also appends (length, base_index + i) to long_len[] / long_at[] starting at slot *n_long,
 * which is advanced.  base_index = number of ops already in the stream.  Returns 0, or LPS_E_ARG when long_cap is too
 * small (nothing is written past long_cap).  Pure host code; thread-safe on disjoint outputs.                      */
int lps_pack_cigar16(const uint32_t *cigar, uint64_t n, uint64_t base_index, uint16_t *out16, uint32_t *long_len,
                     uint64_t *long_at, uint64_t long_cap, uint64_t *n_long);
/* The 8-bit wire format (lps_read_batch.cigar8): appends n BAM CIGAR ops.  out8[0..n) receives the bytes; ops that need it append
 * their 16-bit form to esc16[] at slot *n_esc (advanced) and, when 4095 bases or longer, (length, base_index + i) to long_len[] /
 * long_at[] at slot *n_long (advanced).  esc_blk[] is indexed by ABSOLUTE op index / 256 (the caller allocates
 * total_ops / 256 + 2 entries for the whole stream and passes the same array to every call); entries up to the block that holds the
 * last appended op + 1 are kept up to date.  Returns 0 or LPS_E_ARG when a table is too small.  Pure host code.            */
int lps_pack_cigar8(const uint32_t *cigar, uint64_t n, uint64_t base_index, uint8_t *out8, uint16_t *esc16, uint64_t esc_cap, uint64_t *n_esc,
                    uint32_t *esc_blk, uint32_t *long_len, uint64_t *long_at, uint64_t long_cap, uint64_t *n_long);
/* The interleaved SEQ + QUAL row of one read (lps_read_batch.sq): 16 * ceil(l_qseq / 10) bytes.                            */
uint64_t lps_sq_row_bytes(int32_t l_qseq);
/* Writes that row from bam_get_seq(aln) (4-bit packing) and bam_get_qual(aln); out must hold lps_sq_row_bytes(l_qseq) bytes.
 * Returns 0 or LPS_E_ARG.  Pure host code; thread-safe on disjoint outputs.                                               */
int lps_pack_sq(const uint8_t *seq4, const uint8_t *qual, int32_t l_qseq, uint8_t *out);
/* lps_pack_sq for reads 0 .. n_reads of a batch in the two-array layout: row r is written at sq + sq_off[r] (16-byte aligned
 * offsets the caller laid out with lps_sq_row_bytes).  Callers may split the reads over threads by passing shifted per-read
 * arrays (the offsets stay absolute).                                                                                     */
int lps_pack_sq_batch(int32_t n_reads, const int32_t *l_qseq, const uint64_t *seq_off, const uint64_t *qual_off, const uint8_t *seq4,
                      const uint8_t *qual, const uint64_t *sq_off, uint8_t *sq);
/* What the kernel reads from a row for query index qi (the same index arithmetic, compiled for the host): bam_seqi code and
 * base quality.  For integrators' and this repo's tests.  Returns 0, or LPS_E_ARG for qi outside [0, l_qseq).              */
int lps_sq_peek(const uint8_t *row, int32_t l_qseq, int32_t qi, uint8_t *seq_code, uint8_t *quality);
/* Same, for buffers that already live in device memory (a batch that stays resident across calls): nothing is copied except the
 * name ranks and flags the host needs (to group the alignments of one read name).  The arrays must stay valid until the next
 * submit on this context.  Either cigar (uint32 ops, narrowed into a buffer of the context) or cigar16 (+ its side table);
 * either seq4 + qual or sq (16-byte aligned).                                                                              */
int lps_batch_submit_device(lps_ctx *ctx, const lps_read_batch *b_dev);

/* ---- phase -------------------------------------------------------------------------------- */
/* BamParser::direct_detect_alleles read filter + BamParser::get_snp + getClip
 * (ParsingBam.cpp:1243-1301, 1303-1634, 1636-1645) and the call-erasing half of
 * SnpParser::filterSNP (ParsingBam.cpp:891-911).  want_host!=0 copies the result to the host. */
int lps_phase_call_alleles(lps_ctx *ctx, const lps_phase_params *p, int want_host, lps_calls *out);
/* VairiantGraph::addEdge (PhasingGraph.cpp:694-889): overlap filter on the device, CNV filter from the clip map
 * (host; it only acts when clip pile-ups exist), merge by read name, fan-out and the ordered float fold of
 * SubEdge::addSubEdge (:25-70).                                                                  */
int lps_phase_build_edges(lps_ctx *ctx, const lps_phase_params *p, int want_host, lps_edges *out);
/* VairiantGraph::phasingProcess + exportResult (PhasingGraph.cpp:286-474, 891-1029, 1049-1077): edgeConnectResult as a
 * segmented sweep on the device (k_sweep.cu; windows beyond 63 successors and LPS_HOST_SWEEP=1 use the host sweep below),
 * then readCorrection.                                                                          */
int lps_phase_solve(lps_ctx *ctx, const lps_phase_params *p, lps_phase_result *out);
/* all of the above for one contig (the body of the loop at PhasingProcess.cpp:113-173) as one asynchronous pipeline: the host
 * waits for the device twice (the allele-calling kernel's counters, the end).  Result arrays live in pinned memory of the
 * context until the next call.  LPS_STAGED=1 runs the three calls above instead.                */
int lps_phase_contig(lps_ctx *ctx, const lps_phase_params *p, lps_phase_result *out);
/* edgeConnectResult on the host, no device: VairiantGraph::edgeConnectResult (PhasingGraph.cpp:286-474) over
 * the one-byte summaries of VariantEdge::findBestEdgePair (:166-228) that the device derives from every edge cell.
 * votes[k*window + d] describes the edge from graph node k to node k+1+d: bits 0-1 link (1 same haplotype, 2 opposite,
 * 0 none), bit 2 weight-20 rule, bit 3 (para+cross) <= 1, bit 4 edgeSimilarRatio < 0.2.  node_type as lps_edges.node_type.
 * Writes the phase-set id (0 = none) and the REF-allele haplotype (-1 = none) of every node; returns the SIMD path taken
 * (0 scalar, 1 AVX2, 2 AVX-512; the environment variable LPS_SWEEP=scalar|avx2|avx512 caps it) or a negative error.   */
int lps_sweep_votes(const lps_phase_params *p, int32_t n_nodes, int32_t window, const int32_t *node_pos, const uint8_t *node_type,
                    const uint8_t *votes, int32_t *node_ps, int8_t *node_hap_ref);

/* ---- haplotag (germline tag dialect) -------------------------------------------------------- */
/* dispatch category of every alignment, in the order of ChromosomeProcessor::processSingleChrom
 * (src/haplotag/HaplotagParsingBam.cpp:457-486); only PROCESSED alignments reach processRead          */
enum { LPS_TAG_PROCESSED = 0, LPS_TAG_LOW_MAPQ = 1, LPS_TAG_UNMAPPED = 2, LPS_TAG_SECONDARY = 3,
       LPS_TAG_SUPPLEMENTARY = 4, LPS_TAG_EMPTY_VARIANTS = 5, LPS_TAG_OTHER = 6 };

typedef struct {
    int32_t mapping_quality;      /* -q qualityThreshold, default 1 (src/haplotag/Haplotag.cpp:60-72)      */
    int32_t mapq_filter;          /* ParsingBamControl::mappingQualityFilter (true for `haplotag`)         */
    int32_t tag_supplementary;    /* --tagSupplementary                                                    */
    int32_t have_reference;       /* ref_string != "" (HaplotagStrategy.cpp:159)                           */
    double percentage_threshold;  /* -p, default 0.6                                                       */
} lps_tag_params;

/* result of lps_tag_reads: one entry per alignment of the batch, arrays owned by the context              */
typedef struct {
    int32_t n_reads;
    const uint8_t *category;      /* LPS_TAG_*                                                             */
    const int8_t *hp;             /* ReadHP: 0 unTag, 1 H1, 2 H2 -> HP:i (HaplotagProcess.cpp:357-361)     */
    const int32_t *ps;            /* PS:i (smallest phase set seen, HaplotagStrategy.cpp:295-298), 0 if untagged */
    const int32_t *pq;            /* PQ:i (HaplotagStrategy.cpp:279-288)                                   */
    const int32_t *h1, *h2;       /* hpCount[GERMLINE_H1], hpCount[GERMLINE_H2]                            */
    /* optional (want_calls): the variants that touched countPS, CSR per alignment; lps_call.allele is the
     * variantsHP value (0 = HP1, 1 = HP2, -1 = only counted towards countPS), lps_call.origin 0 SNP in an M op,
     * 1 SNP by the D-op rule, 2 indel                                                                      */
    uint64_t n_calls;
    const uint64_t *call_off;
    const lps_call *calls;
    /* ReadStatistics (src/haplotag/HaplotagProcess.h:21-45) reduced over the batch                         */
    int64_t total_alignment, total_supplementary, total_secondary, total_unmapped, total_tag, total_untag,
            total_lower_quality, total_other_case, total_empty_variant, total_high_similarity,
            total_without_variant, total_hp1, total_hp2, total_hp0;
} lps_tag_result;

/* GermlineHaplotagChrProcessor::processRead for every alignment of the batch: CigarParser::parsingCigar
 * (HaplotagParsingBam.cpp:541-647) + GermlineHaplotagStrategy::judgeSnpHap / judgeDeletionHap / judgeReadHap
 * (HaplotagStrategy.cpp:20-300).  The variant table must carry hp1_is_alt and ps (lps_contig_set_variants). */
int lps_tag_reads(lps_ctx *ctx, const lps_tag_params *p, int want_calls, lps_tag_result *out);

/* ---- somatic family: union variant map, the two extract passes, somatic tagging ------------------------------ */
/* TUMOR side of the union map std::map<int, MultiGenomeVar> (src/haplotag/HaplotagType.h:146-162).  All arrays are
 * parallel to the table given to lps_contig_set_variants (same n, same ascending positions), which carries the NORMAL
 * records; a position may hold a NORMAL record, a TUMOR record, or both.  Tumor-present positions are numbered
 * 0..n_tum-1 in ascending order ("tumor slots"); every per-position product of this family is indexed by slot.     */
typedef struct {
    int32_t n;
    const uint8_t *nor_present;   /* MultiGenomeVar::isExists(NORMAL); NULL = present everywhere                 */
    const uint8_t *tum_present;   /* MultiGenomeVar::isExists(TUMOR)                                             */
    const uint8_t *ref0, *alt0;   /* first characters of the tumor record's REF / ALT                            */
    const uint16_t *ref_len, *alt_len;
    const uint8_t *gt_kind;       /* GenomeType of the tumor record (1 phased het, 2 unphased het, 3 unphased hom) */
    const uint8_t *hp1_is_alt;    /* phased tumor records only (unused by the passes below, kept for the logs)   */
    const int32_t *ps;            /* PhasedSet of a phased tumor record (must not be -1, HaplotagStrategy.cpp:337) */
    const uint8_t *is_somatic;    /* MultiGenomeVar::isSomaticVariant (set by SomaticVarCaller::getSomaticFlag)  */
    const int8_t *derive_hp;      /* MultiGenomeVar::somaticReadDeriveByHP: 0 none, 1 GERMLINE_H1, 2 GERMLINE_H2 */
} lps_tumor_variants;
int lps_contig_set_tumor_variants(lps_ctx *ctx, const lps_tumor_variants *t);

/* PosBase counters (src/haplotag/HaplotagType.h:165-224), in this order                                          */
enum { LPS_PB_ALT = 0, LPS_PB_A, LPS_PB_C, LPS_PB_G, LPS_PB_T, LPS_PB_UNKNOWN, LPS_PB_DEPTH, LPS_PB_DEL,
       LPS_PB_MPQ_ALT, LPS_PB_MPQ_A, LPS_PB_MPQ_C, LPS_PB_MPQ_G, LPS_PB_MPQ_T, LPS_PB_MPQ_UNKNOWN, LPS_PB_MPQ_DEPTH,
       LPS_PB_FIELDS };
/* read-case counters of SomaticData (HaplotagType.h:226-233), filled by classifyReadsByCase                      */
enum { LPS_CASE_CLEAN_HP3 = 0, LPS_CASE_PURE_H1_1, LPS_CASE_PURE_H2_1, LPS_CASE_PURE_H3, LPS_CASE_MIXED, LPS_CASE_UNTAG,
       LPS_CASE_FIELDS };
/* derived per-position ratios: PosBase::{VAF, nonDelVAF, filteredMpqVAF, lowMpqReadRatio, delRatio} and, tumor pass only,
 * SomaticData::{Mixed_HP, pure_H1_1, pure_H2_1, pure_H3}_readRatio (HaplotagType.h:184-190, 239-242)                  */
enum { LPS_RF_VAF = 0, LPS_RF_NONDEL_VAF, LPS_RF_MPQ_VAF, LPS_RF_LOW_MPQ_RATIO, LPS_RF_DEL_RATIO, LPS_RF_MIXED_RATIO,
       LPS_RF_PURE_H1_1_RATIO, LPS_RF_PURE_H2_1_RATIO, LPS_RF_PURE_H3_RATIO, LPS_RF_FIELDS };
/* PosBase::{germlineHaplotypeImbalanceRatio, percentageOfGermlineHp}, SomaticData::{allelicImbalanceRatio,
 * somaticHaplotypeImbalanceRatio} (tumor pass)                                                                       */
enum { LPS_RD_GERMLINE_IMBALANCE = 0, LPS_RD_PCT_GERMLINE_HP, LPS_RD_ALLELIC_IMBALANCE, LPS_RD_SOMATIC_IMBALANCE, LPS_RD_FIELDS };
enum { LPS_READHP_FIELDS = 9 };   /* ReadHP: unTag 0, H1, H2, H3, H4, H1_1, H1_2, H2_1, H2_2 (HaplotagType.h:97-108)  */
enum { LPS_WINDOW = 100, LPS_WINDOW_BINS = 2 * LPS_WINDOW + 1 };   /* getWindowsDiffRef windowSize, offsets -100..100 */

/* per-alignment products shared by the three passes; arrays owned by the context                                 */
typedef struct {
    int32_t n_reads;
    const uint8_t *category;      /* LPS_TAG_* dispatch category                                                 */
    const int8_t *read_hp;        /* ReadHP of processed alignments, 0 otherwise                                 */
    const int32_t *ps;            /* somatic tagging: PS:i value, -1 = VarData::NONE_PHASED_SET (no PS tag), 0 untagged */
    const int32_t *pq;
    const int32_t *h1, *h2, *h3;  /* hpCount[1..3]; hpCount[4] is never incremented by the reference             */
    const uint8_t *n_ps;          /* min(norCountPS.size(), 2)                                                   */
    const int32_t *end_pos;       /* ref_pos after parsingCigar (ReadVarHpCount::endPos)                          */
    const int32_t *read_len;      /* query_pos after parsingCigar (ReadVarHpCount::readLength)                    */
} lps_read_tags;

typedef struct {
    int32_t n_tum;
    const int32_t *tum_var;        /* [n_tum] variant index of each tumor slot                                    */
    const int32_t *pos_base;       /* [n_tum][LPS_PB_FIELDS]                                                      */
    const int32_t *read_hp_count;  /* [n_tum][9] PosBase::ReadHpCount[ReadHP]                                     */
    lps_read_tags reads;
    /* ---- tumor pass only (NULL / 0 after lps_extract_normal) ---- */
    const int32_t *somatic_read_hp_count; /* [n_tum][9] SomaticData::somaticReadHpCount                          */
    const int32_t *case_count;     /* [n_tum][LPS_CASE_FIELDS]                                                    */
    const int32_t *allele_count;   /* [n_tum][2] SomaticData::alleleCount                                         */
    const int32_t *window_hist;    /* [n_tum][2][LPS_WINDOW_BINS] entries of PosSomaticOffsetBase[allele] per offset
                                      (bin = offset + 100): all the DenseAlt filter reads (SomaticVarCaller.cpp:1160) */
    uint64_t n_window_items;       /* (alignment, tumor position) pairs whose window was scanned                  */
    /* ---- postProcess of both passes (SomaticVarCaller.cpp:176-210, 520-603; calculateBaseCommonInfo :13-40), host arithmetic
     * in the reference's own float / double types; zero for slots whose tumor record is not a SNP / insertion / deletion ---- */
    const float *ratios_f;         /* [n_tum][LPS_RF_FIELDS]                                                      */
    const double *ratios_d;        /* [n_tum][LPS_RD_FIELDS]                                                      */
    const int32_t *case_read_count;/* [n_tum] SomaticData::CaseReadCount (tumor pass)                             */
    /* per alignment, CSR: every variant that entered variantsHP or tumorSnpPosVec.  lps_call.allele = variantsHP
     * (SnpHP 1 H1, 2 H2, 3 H3, 0 none), lps_call.quality bit0 = in tumorSnpPosVec, bit1 = in tumorAllelePosVec
     * (ReadVarHpCount::posHpPairs, tumorPosReadCorrBaseHP; SomaticVarCaller.cpp:407-459)                         */
    uint64_t n_calls;
    const uint64_t *call_off;
    const lps_call *calls;
} lps_extract_result;

/* pass A of SomaticVarCaller::extractSomaticData over the NORMAL BAM: ExtractNorDataChrProcessor::processRead with the
 * ExtractNorDataCigarParser hooks and countBaseNucleotide (src/somatic_haplotag/SomaticVarCaller.cpp:123-293,
 * src/haplotag/HaplotagParsingBam.cpp:682-730).  postProcess' ratios (:176-210) are left to the host.            */
int lps_extract_normal(lps_ctx *ctx, const lps_tag_params *p, lps_extract_result *out);
/* pass B over the TUMOR BAM: ExtractTumDataChrProcessor::processRead / classifyReadsByCase, the ExtractTumDataCigarParser
 * hooks, judgeSomaticSnpHap / judgeSomaticReadHap and getWindowsDiffRef (SomaticVarCaller.cpp:334-759,
 * HaplotagStrategy.cpp:315-602, 617-638).  postProcess' ratios (:520-603) are left to the host.                   */
int lps_extract_tumor(lps_ctx *ctx, const lps_tag_params *p, lps_extract_result *out);

typedef struct {
    lps_read_tags reads;           /* read_hp = ReadHP after inheritHaplotype -> HP:Z, ps -> PS:i, pq -> PQ:i      */
    const int8_t *hp_before;       /* [n_reads] ReadHP before inheritHaplotype                                    */
    const float *derive_similarity;/* [n_reads] deriveByHpSimilarity (SomaticHaplotagProcess.cpp:493)              */
    int32_t n_tum;
    const int32_t *tum_var;
    const int32_t *hp_before_count;   /* [n_tum][9] chrReadHpResult readHpCounter before inheritance (somatic positions) */
    const int32_t *hp_after_count;    /* [n_tum][9] ... after inheritance                                          */
    const int32_t *h3_before_count;   /* [n_tum][9] somaticBaseReadHpCounter before inheritance                    */
    const int32_t *h3_after_count;    /* [n_tum][9] ... after                                                      */
    const int32_t *cover_start, *cover_end; /* [n_tum] recordAlignCoverRegion (INT_MAX / INT_MIN when untouched)  */
    uint64_t n_calls;              /* per alignment CSR of variantsHP (lps_call.allele = SnpHP, quality bit2 = somatic variant) */
    const uint64_t *call_off;
    const lps_call *calls;
    /* ReadStatistics (HaplotagProcess.h:21-45)                                                                   */
    int64_t total_alignment, total_supplementary, total_secondary, total_unmapped, total_tag, total_untag,
            total_lower_quality, total_other_case, total_empty_variant, total_high_similarity, total_cross_two_block,
            total_without_variant, total_read_only_h3, total_hp[LPS_READHP_FIELDS];
} lps_somatic_tag_result;

/* SomaticHaplotagChrProcessor::judgeHaplotype for every alignment of the TUMOR batch: SomaticHaplotagCigarParser hooks,
 * SomaticHaplotagStrategy::judgeTumorOnlySnpHap, judgeSomaticReadHap, inheritHaplotype and the PS rule
 * (src/somatic_haplotag/SomaticHaplotagProcess.cpp:310-579, src/haplotag/HaplotagStrategy.cpp:452-668).          */
int lps_somatic_tag_reads(lps_ctx *ctx, const lps_tag_params *p, int want_calls, lps_somatic_tag_result *out);

/* ---- tumor purity: the host math that consumes the two extract passes ---------------------------------------------- *
 * TumorPurityEstimator::estimateTumorPurity (src/somatic_haplotag/TumorPurityEstimator.cpp:31-84).  One entry per tumor position
 * of the job, contigs concatenated in chrVec order: the `ratios_d` / `read_hp_count` columns of lps_extract_result of the TUMOR
 * pass (germline imbalance) and of the NORMAL pass (germline imbalance, percentage of germline HP reads, ReadHpCount[H1], [H2]).
 * Positions no read touched may be included: their zero ratios are rejected by the first filters, as in the reference.
 * Pure host arithmetic (double); no context and no device work.                                                           */
typedef struct {
    int32_t n;
    const double *tumor_germline_imbalance;    /* SomaticData::base.germlineHaplotypeImbalanceRatio                        */
    const double *normal_germline_imbalance;   /* chrPosNorBase[chr][pos].germlineHaplotypeImbalanceRatio                   */
    const double *normal_pct_germline_hp;      /* ...percentageOfGermlineHp                                                 */
    const int32_t *normal_h1, *normal_h2;      /* ...ReadHpCount[H1], ReadHpCount[H2]                                       */
    uint8_t *used;                             /* optional out [n]: SomaticData::statisticPurity (markStatisticFlag)        */
} lps_purity_input;
typedef struct {
    double purity;                             /* 0.0 when the estimate failed (reference prints an error and uses 0.0)     */
    int32_t ok;
    int32_t read_count_threshold;              /* bimodal-valley threshold on the normal germline-HP read count             */
    double median, q1, q3, iqr, lower_whisker, upper_whisker;   /* BoxPlotValue after the outlier round                    */
    int32_t n_after_lcvf, n_used;
    int32_t filtered_normal_imbalance_zero, filtered_tumor_imbalance_zero, filtered_normal_imbalance_high, filtered_normal_read_count,
            filtered_pct_germline_hp, filtered_valley, filtered_outliers;              /* FilterCounts of the _purity.out log */
    int32_t n_outliers_left;                   /* BoxPlotValue::outliers of the final statistic (values outside the new whiskers) */
} lps_purity_result;
int lps_estimate_purity(const lps_purity_input *in, lps_purity_result *out);

/* ---- somatic calling: the host stage between the extract passes and the tagging pass ------------------ *
 * SomaticVarCaller::variantCalling without extraction and logs (src/somatic_haplotag/SomaticVarCaller.cpp:816-866):
 * setFilterParamsWithPurity (:951-1060), getDenseTumorSnpInterval (:1243-1351), somaticFeatureFilter (:1062-1230),
 * calibrateReadHP (:1366-1403), calculateReadSetHP (:1418-1439), statisticSomaticPosReadHP (:1441-1518) and getSomaticFlag
 * (:2397-2412), for one contig.  `normal` / `tumor` are the results of lps_extract_normal / lps_extract_tumor of that contig
 * (same tumor slots).  Pure host code.                                                                                     */
typedef struct {
    int32_t n_tum;
    const int32_t *pos;            /* [n_tum] 0-based position of every tumor slot, ascending                              */
    const uint8_t *callable;       /* [n_tum] 1 when the TUMOR record is a SNP, an insertion or a deletion                 */
    const lps_extract_result *normal, *tumor;
    double purity;                 /* lps_estimate_purity's value or --tumorPurity                                          */
    int32_t enable_filter;         /* CallerConfig::enableFilter (default 1)                                                */
    double percentage_threshold;   /* -p, as in lps_tag_params                                                              */
} lps_somatic_call_input;
enum { LPS_FILTER_TINC = 0, LPS_FILTER_MESSY_READ, LPS_FILTER_READ_COUNT, LPS_FILTER_HAP_CONSISTENCY, LPS_FILTER_VARIANT_CLUSTER,
       LPS_FILTER_DENSE_ALT, LPS_FILTER_FIELDS };
typedef struct {
    /* caller-allocated, one entry per tumor slot; any pointer may be NULL */
    uint8_t *touched;              /* the position has a SomaticData entry (a tumor alignment reached it)                   */
    uint8_t *is_somatic;           /* SomaticData::isHighConSomaticSNP -> MultiGenomeVar::isSomaticVariant                  */
    int8_t *derive_hp;             /* SomaticData::somaticReadDeriveByHP (SnpHP: 0 none, 1 H1, 2 H2)                        */
    uint8_t *is_filter_out;        /* SomaticData::isFilterOut                                                              */
    uint8_t *filtered_by;          /* [n_tum][LPS_FILTER_FIELDS] the per-filter flags                                       */
    uint8_t *in_dense_interval;
    float *mean_alt_per_var_read, *z_score;
    int32_t *interval_snp_count, *min_distance, *dense_alt_same_count;
    /* caller-allocated, one entry per alignment of the tumor batch; may be NULL */
    int8_t *read_hp;               /* ReadVarHpCount::hpResult after calibration (ReadHP), -1 without a tumor position      */
    int32_t *read_h3;              /* ReadVarHpCount::HP3 after calibrateReadHP, -1 without a tumor position                */
    /* written by the call */
    int32_t tier;                  /* FilterTier chosen from the purity: 1 (0.9-1.0) .. 5 (below 0.3)                       */
    int32_t n_somatic;             /* positions flagged somatic                                                             */
} lps_somatic_call_result;
/* Returns LPS_E_DATA where the reference prints "[ERROR](calibrate read HP) ..." / "(statistic all read HP) ..." and exits. */
int lps_somatic_call(const lps_somatic_call_input *in, lps_somatic_call_result *out);
/* SomaticVarFilterParams as setFilterParamsWithPurity fills it for a purity (SomaticVarCaller.h:59-104, SomaticVarCaller.cpp:951-1045):
 * what lps_somatic_call applies, for the header of the calling log (<prefix>_somatic_var.out).                               */
typedef struct {
    int32_t tier;                                  /* 1 (purity 0.9-1.0) .. 5 (below 0.3 or out of range)                     */
    float tumor_purity, nor_vaf_max;
    int32_t nor_depth_min;
    float messy_read_ratio;
    int32_t read_count_min;
    float hap_consistency_vaf_max;
    int32_t hap_consistency_read_count_max, hap_consistency_somatic_read_min;
    float interval_snp_count_vaf_max;
    int32_t interval_snp_count_read_count_max, interval_snp_count_min;
    float z_score_max, dense_alt_condition1, dense_alt_condition2;
    int32_t dense_alt_same_count_min;
} lps_somatic_filter_params;
int lps_somatic_filter_params_of(double purity, lps_somatic_filter_params *out);


/* ---- timing / accounting ------------------------------------------------------------------ */
typedef struct {
    float ms_call_alleles;   /* device time of the allele-calling kernels of the last call      */
    float ms_tag_reads;      /* device time of the last lps_tag_reads / lps_extract_* / lps_somatic_tag_reads */
    float ms_build_edges;
    float ms_read_correction;
    float ms_h2d;
    float ms_d2h;
    float ms_kernel_call_alleles; /* the k_call_alleles launch alone (CUDA events on the launching stream) */
    float ms_kernel_fold_edges;   /* the k_fold_edges launch alone                                          */
    float ms_kernel_window_diff;  /* the k_window_diff launch alone (lps_extract_tumor)                     */
    float ms_wall_call_alleles;   /* host wall clock of the last lps_phase_call_alleles              */
    float ms_wall_build_edges;    /* ... lps_phase_build_edges                                        */
    float ms_wall_solve;          /* ... lps_phase_solve                                              */
    float ms_host_filters;        /* overlap filter + CNV intervals/filter on the host                */
    float ms_host_sweep;          /* edgeConnectResult chain on the host, over the one-byte votes     */
    uint64_t kernel_launches; /* kernels launched by this context since creation                 */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    int32_t sweep_simd;           /* code path of the last host sweep: 0 scalar, 1 AVX2, 2 AVX-512   */
    float ms_kernel_bgzf;         /* k_bgzf_inflate of the last lps_bgzf_inflate[_device] call       */
    float ms_sweep;               /* device time of the edgeConnectResult kernels (lps_phase_contig)  */
    uint32_t sweep_fallbacks;     /* contigs whose segmented sweep did not verify and was redone sequentially (on the device) */
    uint32_t slow_path_contigs;   /* contigs redone stage by stage because host code was needed between the kernels          */
} lps_stats;
int lps_get_stats(lps_ctx *ctx, lps_stats *out);
/* CUDA events on the context's stream (the stream every kernel of this library is launched on), so a
 * caller can time a region on the device: slots 0..3.                                                */
int lps_event_record(lps_ctx *ctx, int slot);
int lps_event_elapsed_ms(lps_ctx *ctx, int slot_a, int slot_b, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* LPS_H */
