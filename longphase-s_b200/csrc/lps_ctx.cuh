// lps_ctx.cuh — context, device buffers and error plumbing of liblps_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/lps.h"

#define LPS_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->fail(LPS_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));            \
            return LPS_E_CUDA;                                                                       \
        }                                                                                            \
    } while (0)

// growable device buffer; never shrinks, so steady-state batches do no cudaMalloc
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }          // lps_ctx_destroy makes the context's device current before it deletes the context
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// pinned host staging buffer
template <typename T> struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    PinBuf() = default;
    PinBuf(const PinBuf &) = delete;
    PinBuf &operator=(const PinBuf &) = delete;
    ~PinBuf() { release(); }
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMallocHost((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// device view of a read batch (pointers are DEVICE pointers)
struct DevBatch {
    int32_t n_reads = 0;
    const int32_t *ref_start = nullptr, *l_qseq = nullptr;
    const uint32_t *n_cigar = nullptr;
    const uint64_t *cigar_off = nullptr, *seq_off = nullptr, *qual_off = nullptr;
    const uint16_t *flag = nullptr;
    const uint8_t *mapq = nullptr;
    const int32_t *name_rank = nullptr;
    // CIGAR stream as it lives in HBM: 16 bits per op, len << 4 | op (BAM's op codes); a length >= 4095 is stored as 0xFFF and its
    // true value sits in the side table (long_at = op index in the stream, ascending; long_len)
    const uint16_t *cigar16 = nullptr;
    uint64_t cigar_len = 0;
    const uint64_t *long_at = nullptr;
    const uint32_t *long_len = nullptr;
    uint32_t n_long = 0;
    const uint8_t *seq4 = nullptr;
    uint64_t seq_bytes = 0;
    const uint8_t *qual = nullptr;
    uint64_t qual_bytes = 0;
    // SEQ + QUAL interleaved (lps_read_batch.sq): seq4 points at the rows, seq_off[r] is the row of read r, qual / qual_off are unused
    int sq = 0;
};

// ---- interleaved SEQ + QUAL rows (include/lps.h, lps_read_batch.sq) -------------------------------------------------------
// A row is made of 16-byte units of ten bases: bytes 0..9 base qualities, bytes 10..14 the ten 4-bit codes in BAM's packing.
// The same arithmetic serves the kernel's gather, the host packer and lps_sq_peek.
constexpr uint32_t SQ_BASES = 10, SQ_UNIT = 16;
__host__ __device__ __forceinline__ uint32_t sq_unit_of(uint32_t qi) { return qi / SQ_BASES; }
// w = the unit as four little-endian words, k = qi - 10 * unit
__host__ __device__ __forceinline__ void sq_extract(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t k, unsigned &code, unsigned &quality) {
    const uint32_t wq = k < 4u ? w0 : k < 8u ? w1 : w2;
    quality = (wq >> ((k & 3u) * 8u)) & 0xFFu;
    const uint32_t sb = 10u + (k >> 1);                       // byte 10..14 holds bases 2 (sb - 10) and 2 (sb - 10) + 1
    const uint32_t ws = sb < 12u ? w2 : w3;
    const unsigned byte = (ws >> ((sb & 3u) * 8u)) & 0xFFu;
    code = (byte >> ((~k & 1u) << 2)) & 0xFu;                 // bam_seqi: even index in the high nibble (unit starts are even)
}

// device view of the variant table
struct DevVariants {
    int32_t n = 0;
    const int32_t *pos = nullptr;
    const uint8_t *ref0 = nullptr, *alt0 = nullptr;
    const uint16_t *ref_len = nullptr, *alt_len = nullptr;
    const uint8_t *hom = nullptr, *danger = nullptr, *filtered = nullptr;
    // one 8-byte record per variant for the walking kernel (k_pack_vrec): x = position, y = ref0 | alt0 << 8 | flags << 16 |
    // min(homopolymer, 255) << 24; flags: 1 strlen(REF) == 1, 2 strlen(ALT) == 1, 4 danger, 8 erased by filterSNP, 16 HP1 carries ALT
    const uint2 *vrec = nullptr;
};

// TUMOR side of the union variant map + per-slot counter block of the somatic family (see lps.h); DEVICE pointers
struct DevSomatic {
    const uint8_t *nor_present = nullptr, *nor_gt = nullptr, *tum_present = nullptr;
    const uint8_t *t_ref0 = nullptr, *t_alt0 = nullptr, *t_gt = nullptr, *is_somatic = nullptr;
    const uint16_t *t_ref_len = nullptr, *t_alt_len = nullptr;
    const int8_t *derive_hp = nullptr;
    const int32_t *slot_of_var = nullptr;   // tumor slot of a variant, -1 when it has no TUMOR record
    const int32_t *tum_var = nullptr;       // [n_tum]
    const int32_t *prev_nor = nullptr;      // position of the last earlier variant with a phased-het NORMAL record, INT_MIN if none
    int32_t n_tum = 0;
    // per-slot counters, every array [n_tum][k]
    int32_t *pos_base = nullptr, *read_hp_count = nullptr, *somatic_read_hp_count = nullptr, *case_count = nullptr, *allele_count = nullptr;
    int32_t *hp_before_count = nullptr, *hp_after_count = nullptr, *h3_before_count = nullptr, *h3_after_count = nullptr;
    int32_t *cover_start = nullptr, *cover_end = nullptr, *window_hist = nullptr;
};

// one (alignment, tumor position) pair whose +-100 window is compared with the reference by k_window_diff
struct WdItem { uint32_t read, slot2, opi, qidx, off; };

enum { LPS_MODE_PHASE = 0, LPS_MODE_GERMLINE = 1, LPS_MODE_EXTRACT_NORMAL = 2, LPS_MODE_EXTRACT_TUMOR = 3, LPS_MODE_SOMATIC_TAG = 4 };

// counters written by the allele-calling kernel (copied back once per call).  Every hot counter sits on its own 128-byte
// line: same-line atomics serialise in one L2 slice, and with ~100 k reads per launch the work counter, the pool counter and
// the clip counter together saturated it.
struct alignas(128) CounterLine { unsigned long long v; unsigned long long pad[15]; };
struct CallCounters {
    CounterLine next_read;      // dynamic read fetch of the persistent k_call_alleles launch (low 32 bits used)
    CounterLine tmp_calls;      // slots handed out from the scratch call pool (in per-warp blocks)
    CounterLine n_calls;        // calls actually written (added once per warp, at its exit)
    CounterLine clips;          // clip events appended
    CounterLine wd_items;       // window-diff work items appended (tumor extract pass)
    CounterLine gathers;        // SNP candidates whose base + quality were gathered (zero-copy accounting)
    CounterLine overflow_cands; // candidate slots needed by reads that overflowed the smem buffer
    CounterLine overflow_reads;
    CounterLine aborted_reads;  // reads dropped by get_snp's bounds check (their later clip events are cancelled)
    CounterLine bad_cigar;      // reads with an unsupported CIGAR op
};

struct lps_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_k[4] = {nullptr, nullptr, nullptr, nullptr};   // ... and the streams its kernels alternate over
    cudaStream_t stream_up = nullptr, stream_down = nullptr;   // lps_bgzf_inflate: upload / download streams of the chunk pipeline
    cudaEvent_t ev[8] = {};
    cudaEvent_t user_ev[4] = {};
    cudaEvent_t kev[6] = {};   // around the hot kernels
    cudaEvent_t ev_clips = nullptr;   // the clip map has reached the pinned staging buffers
    cudaEvent_t ev_fork = nullptr;    // main stream -> side stream
    cudaStream_t stream_side = nullptr;   // work that is off the contig's critical path (the clip map: sort, run-length encoding, copy)
    DevBuf<uint8_t> d_cub_tmp_side;
    std::string err;
    int err_code = 0;
    lps_stats stats = {};

    // ---- contig ----
    DevBuf<char> d_ref;
    int64_t ref_len = 0;
    DevBuf<int32_t> d_vpos;
    DevBuf<uint8_t> d_vref0, d_valt0, d_vhom, d_vdanger, d_vfiltered;
    DevBuf<uint16_t> d_vref_len, d_valt_len;
    std::vector<int32_t> h_vpos;
    std::vector<uint8_t> h_vhom, h_vdanger, h_vfiltered;
    DevVariants var;
    DevBuf<uint8_t> d_vhp1_is_alt;
    DevBuf<int32_t> d_vps;
    DevBuf<uint2> d_vrec;
    bool have_tag_variants = false;
    DevBuf<int8_t> d_pq_lut;
    bool have_variants = false;
    int is_ont = 0;

    // ---- batch ----
    DevBuf<int32_t> d_ref_start, d_l_qseq, d_name_rank;
    DevBuf<uint32_t> d_n_cigar, d_cigar;
    DevBuf<uint8_t> d_bgzf_in, d_bgzf_out, d_bgzf_status;   // lps_bgzf_inflate
    DevBuf<lps_bgzf_block> d_bgzf_blocks;
    DevBuf<uint8_t> d_defl_slots, d_defl_out;                // lps_bgzf_deflate: one fixed slot per member, the contiguous stream
    DevBuf<uint64_t> d_defl_sizes, d_defl_off;
    std::vector<uint8_t> h_bgzf_status;
    DevBuf<uint16_t> d_cigar16;                     // the CIGAR stream in 16 bits per op (what the kernels read)
    DevBuf<uint8_t> d_cigar8;                       // 8-bit wire format of a submitted batch, expanded into d_cigar16 on arrival
    DevBuf<uint16_t> d_cigar_esc16;
    DevBuf<uint32_t> d_cigar_esc_blk;
    DevBuf<unsigned int> d_n_long;                  // escaped ops found while narrowing a uint32 stream on the device
    DevBuf<uint64_t> d_long_keys;                   // ... (op index << 28 | length), sorted into the side table
    DevBuf<uint32_t> d_cigar_long_len;
    DevBuf<uint64_t> d_cigar_long_at;
    DevBuf<uint64_t> d_cigar_off, d_seq_off, d_qual_off;
    DevBuf<uint16_t> d_flag;
    DevBuf<uint8_t> d_mapq, d_seq4, d_qual;
    DevBatch batch;
    std::vector<int32_t> h_name_rank;
    std::vector<int32_t> h_multi_members, h_multi_group_off;   // alignments of names that occur more than once, grouped by name
    DevBuf<int32_t> d_multi_members, d_multi_group_off, d_multi_kept, d_dead_list, d_pos_of_read;
    DevBuf<uint32_t> d_multi_ncalls;
    uint64_t sum_l_qseq = 0;
    bool have_batch = false;
    bool zero_copy = false;                         // SEQ/QUAL stayed in pinned host memory (gathered over PCIe)

    // ---- allele calls ----
    DevBuf<lps_call> d_calls_tmp, d_calls;          // scratch pool (allocation order) and final CSR
    DevBuf<uint64_t> d_tmp_start, d_call_off;       // per read
    DevBuf<uint32_t> d_ncalls;
    DevBuf<uint8_t> d_status;
    DevBuf<uint32_t> d_clip_keys, d_clip_keys_sorted, d_clip_unique, d_clip_counts;
    DevBuf<uint2> d_clip_meta;
    DevBuf<int32_t> d_abort_of_read;
    DevBuf<uint4> d_work;                           // read descriptors of k_prep_reads (3 x uint4 each), four segments
    DevBuf<uint32_t> d_seg_count;
    DevBuf<unsigned long long> d_dbg_times;
    int sm_count = 148;
    DevBuf<int32_t> d_num_runs;
    DevBuf<CallCounters> d_counters, d_counters2;
    DevBuf<uint8_t> d_ovf_cand;                     // candidate lists of the overflow pass (reads with more candidates than the smem buffer)
    DevBuf<uint32_t> d_overflow_reads;
    DevBuf<uint64_t> d_overflow_cand, d_overflow_off;
    DevBuf<uint8_t> d_cub_tmp;
    uint64_t n_calls = 0;
    bool have_calls = false;
    std::vector<uint64_t> h_call_off;
    std::vector<lps_call> h_calls;
    std::vector<uint8_t> h_status;
    std::vector<int32_t> h_clip_pos, h_clip_front, h_clip_back;
    bool host_calls_valid = false;

    // ---- haplotag ----
    DevBuf<int8_t> d_tag_hp;
    DevBuf<int32_t> d_tag_ps, d_tag_pq, d_tag_h1, d_tag_h2;
    DevBuf<uint8_t> d_tag_cat;
    std::vector<int8_t> h_tag_hp;
    std::vector<int32_t> h_tag_ps, h_tag_pq, h_tag_h1, h_tag_h2;
    std::vector<uint8_t> h_tag_cat;
    std::vector<uint16_t> h_flag;

    // ---- somatic family ----
    DevBuf<uint8_t> d_nor_present, d_nor_gt, d_tum_present, d_t_ref0, d_t_alt0, d_t_gt, d_is_somatic;
    DevBuf<uint16_t> d_t_ref_len, d_t_alt_len;
    DevBuf<int8_t> d_derive_hp;
    DevBuf<int32_t> d_slot_of_var, d_tum_var, d_prev_nor, d_som_counters;
    DevBuf<WdItem> d_wd_items;
    DevSomatic som;
    size_t som_counter_words = 0;
    bool have_tumor_variants = false;
    std::vector<int32_t> h_tum_var, h_som_counters, h_case_reads;
    std::vector<uint8_t> h_t_alt0;
    std::vector<uint16_t> h_t_ref_len, h_t_alt_len;
    std::vector<float> h_ratios_f;
    std::vector<double> h_ratios_d;
    DevBuf<int32_t> d_tag_h3, d_tag_end, d_tag_len;
    DevBuf<uint8_t> d_tag_nps;
    DevBuf<int8_t> d_tag_hpb;
    DevBuf<float> d_tag_sim;
    std::vector<int32_t> h_tag_h3, h_tag_end, h_tag_len;
    std::vector<uint8_t> h_tag_nps;
    std::vector<int8_t> h_tag_hpb;
    std::vector<float> h_tag_sim;
    uint64_t n_wd_items = 0;

    // ---- graph ----
    std::vector<int32_t> h_aln_read;                // stage-C alignments (batch index), BAM order
    std::vector<uint8_t> h_read_dead;               // per read: removed by the overlap filter
    std::vector<int32_t> h_cnv_start, h_cnv_end;
    DevBuf<uint8_t> d_read_dead, d_call_erased;
    bool have_erased = false;
    DevBuf<uint64_t> d_var_lastw;                   // per variant: (read_idx+1) << 3 | type of the last writer
    DevBuf<int32_t> d_node_of_var, d_node_var;
    DevBuf<uint8_t> d_node_type;
    DevBuf<uint64_t> d_aln_keys, d_aln_keys_sorted; // (rank << 32 | read) of alive alignments
    DevBuf<uint32_t> d_alive_cnt;                   // alive calls per read
    DevBuf<uint64_t> d_grp_off;                     // offset of each sorted alignment inside M
    DevBuf<uint32_t> d_M;                           // merged calls: node << 2 | allele << 1 | q_hi
    DevBuf<uint32_t> d_M_node, d_M_node_sorted, d_M_idx, d_M_idx_sorted;
    DevBuf<uint32_t> d_M_gend;                      // end (exclusive) of the merged group that owns entry m
    DevBuf<uint32_t> d_node_cnt;
    DevBuf<uint64_t> d_node_off;
    DevBuf<float> d_weights;
    DevBuf<uint8_t> d_vote_info;                    // [n_nodes][lps_vote_row_stride(window)] vote bytes of voter k, shifted to 16-node blocks (host_phase.cpp)
    DevBuf<int8_t> d_last_link;                     // [n_nodes] largest successor offset a node links to, -1 if none
    PinBuf<int8_t> p_last_link;
    DevBuf<int32_t> d_node_pos, d_n_nodes;            // d_n_nodes[0]: node count of the graph, as the device knows it
    DevBuf<uint16_t> d_sweep_meta;                  // per node: type | gap << 3 | (last_link + 1) << 8 (k_fold_edges -> k_sweep)
    DevBuf<uint8_t> d_sweep_flags, d_sweep_halo, d_sweep_flip, d_sweep_multi;
    DevBuf<int32_t> d_sweep_first_nb, d_sweep_ok, d_sweep_start;
    PinBuf<uint8_t> p_vote_info;                    // pinned staging of the vote bytes for the host sweep
    DevBuf<unsigned long long> d_edge_counters;     // [0] contrib, [1] far
    DevBuf<uint2> d_tie_groups;
    DevBuf<uint32_t> d_M_unsorted, d_tie_off, d_tie_stage;
    DevBuf<unsigned int> d_n_tie;
    DevBuf<int32_t> d_first_pos, d_last_pos, d_ps_final;
    DevBuf<int8_t> d_hap_final;
    int32_t n_nodes = 0;
    int32_t window = 0;
    uint64_t n_merged = 0;
    bool have_graph = false;
    std::vector<int32_t> h_node_var;
    std::vector<uint8_t> h_node_type;
    std::vector<float> h_weights;
    uint64_t n_contrib = 0, n_contrib_far = 0;

    // ---- solution ----
    DevBuf<int32_t> d_ps, d_hp_counts;
    DevBuf<int8_t> d_hap_ref, d_read_hp;
    std::vector<int32_t> h_ps, h_hp_counts;
    std::vector<int8_t> h_hap_ref, h_read_hp, h_hap_sweep;
    std::vector<int32_t> h_ps_sweep;

    // ---- asynchronous per-contig path (lps_phase_contig): results and the clip map land in pinned memory ----
    PinBuf<uint32_t> p_clip_unique, p_clip_counts;
    PinBuf<int32_t> p_num_runs, p_ps, p_ps_sweep, p_hp_counts, p_status;
    PinBuf<int8_t> p_hap, p_hap_sweep, p_read_hp;
    int32_t n_clip_events = 0;
    bool clips_pending = false;

    int fail(int code, const std::string &msg) {
        err_code = code;
        err = msg;
        return code;
    }
};

// kernels (k_*.cu)
int lps_launch_annotate(lps_ctx *ctx);
int lps_prepare_call_alleles(lps_ctx *ctx);   // per-device function attributes of k_call_alleles (called by lps_ctx_create)
int lps_launch_call_alleles(lps_ctx *ctx, const lps_phase_params *p, const lps_tag_params *t = nullptr, int want_calls = 0,
                            int mode = -1 /* LPS_MODE_*; -1: PHASE when t is null, GERMLINE otherwise */, bool defer_clips = false);
int lps_finish_clips(lps_ctx *ctx);   // waits for the clip map of a deferred call and builds h_clip_pos / front / back
int lps_launch_window_diff(lps_ctx *ctx, int have_reference);
int lps_launch_overlap_filter(lps_ctx *ctx, const lps_phase_params *p);
int lps_launch_build_edges(lps_ctx *ctx, const lps_phase_params *p, bool sync_ties);
int lps_fetch_graph_counts(lps_ctx *ctx);
int lps_launch_sweep(lps_ctx *ctx, const lps_phase_params *p, int n_upper);
int lps_launch_read_correction(lps_ctx *ctx, const lps_phase_params *p);
// host restatements that sit between the kernels (host_phase.cpp)
void lps_host_index_names(lps_ctx *ctx);
void lps_host_cnv_intervals(const std::vector<int32_t> &pos, const std::vector<int32_t> &front,
                            const std::vector<int32_t> &back, std::vector<int32_t> &cs, std::vector<int32_t> &ce);
int lps_host_cnv_filter(lps_ctx *ctx, std::vector<uint8_t> &erased);
void lps_host_post_process(int n_tum, const int32_t *tum_var, const uint8_t *t_alt0, const uint16_t *t_ref_len, const uint16_t *t_alt_len,
                           const int32_t *pos_base, const int32_t *read_hp_count, const int32_t *case_count, bool tumor, float *rf,
                           double *rd, int32_t *case_reads);
int lps_vote_row_stride(int window);
int lps_host_sweep(const lps_phase_params *p, int32_t n_nodes, int32_t window, const int32_t *node_pos, const uint8_t *node_type,
                   const uint8_t *votes, const int8_t *last_link, int32_t *node_ps, int8_t *node_hap_ref);
