// k_call_alleles.cu — KERNEL 1 (phase dialect): per-read CIGAR walk + allele call at every overlapped
// SNP / indel.  Replaces BamParser::direct_detect_alleles' read filter, BamParser::get_snp and getClip
// (reference src/phase/ParsingBam.cpp:1243-1301, 1303-1634, 1636-1645) and the call-erasing half of
// SnpParser::filterSNP (:891-911).
//
// Mapping: ONE WARP PER READ.
//   * the CIGAR is streamed in chunks of 32*K ops (K ops per lane, 128-bit loads, the next chunk is in
//     flight while the current one is scanned); a warp exclusive scan of (ref advance, query advance)
//     replaces the sequential walk;
//   * the pending variant position is compared with the chunk's reference span once per chunk; only a
//     hit does more work (owner lane found by ballot), so the common "no variant in this chunk" path is
//     a handful of instructions;
//   * SNP hits are only RECORDED (variant index, query index) in a per-warp shared-memory candidate
//     list; the scattered seq-nibble / base-quality gathers happen afterwards, 32 at a time, so their
//     DRAM latency overlaps instead of serialising inside the scan;
//   * calls are compacted with warp ballots and written contiguously per read into a scratch pool
//     (one atomicAdd per read); a CSR in read order is rebuilt by k_gather_calls after a prefix sum.
//
// Sequential quirks of the reference that are reproduced: cursor == lower_bound of the
// op start; only the FIRST pending variant of a D op is examined (homopolymer >= 3 rule); a variant
// whose query index is beyond l_qseq drops the whole read but keeps the clips seen before it; indel
// alleles look at the op that follows the M op; clips are FRONT iff the CIGAR index is 0.
#include <cub/cub.cuh>
#include "lps_ctx.cuh"

namespace {

constexpr int WARPS_PER_CTA = 8;
constexpr int CAND_CAP = 192;        // candidates buffered per warp in shared memory
constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t PAD_OP = 1u;     // zero-length insertion: advances nothing

// candidate encoding: x = kind << 30 | payload
//   kind 0: SNP seen inside an M/=/X op, payload = query index
//   kind 1: SNP seen by the D-op rule,   payload = query index
//   kind 2: resolved indel call,         payload = origin << 2 | allele << 1 | danger
struct Cand { int32_t var; uint32_t x; };

struct K1Args {
    DevBatch b;
    DevVariants v;
    int mapping_quality;
    int have_reference;
    int apply_filter;
    int last_var_pos;
    lps_call *calls_tmp;
    unsigned long long calls_cap;
    uint64_t *tmp_start;
    uint32_t *ncalls;
    uint8_t *status;
    uint32_t *clip_keys;
    unsigned long long clip_cap;
    CallCounters *counters;
    // overflow pass
    const uint32_t *overflow_reads;   // null in the main pass
    const uint64_t *overflow_off;
    Cand *overflow_buf;
    uint32_t *overflow_list_out;      // main pass: list of reads that overflowed
    uint64_t *overflow_need_out;      // main pass: candidates they need
    uint32_t overflow_list_cap;
};

__device__ __forceinline__ int warp_lower_bound(const int32_t *__restrict__ pos, int n, int key, int lane) {
    int lo = 0, hi = n;   // answer in [lo, hi]
    while (hi - lo > 32) {
        int step = (hi - lo + 31) >> 5;
        int q = lo + (lane + 1) * step - 1;
        if (q > hi - 1) q = hi - 1;
        bool pred = pos[q] < key;
        unsigned m = __ballot_sync(FULL, pred);
        int cnt = __popc(m);
        int nlo = cnt ? (min(lo + cnt * step - 1, hi - 1) + 1) : lo;
        int nhi = cnt < 32 ? min(lo + (cnt + 1) * step - 1, hi - 1) : hi;
        lo = nlo; hi = nhi;
    }
    int q = lo + lane;
    bool pred = (q < hi) && (pos[q] < key);
    return lo + __popc(__ballot_sync(FULL, pred));
}

template <int K>
__device__ __forceinline__ void load_ops(const uint32_t *__restrict__ cig, uint64_t total, int64_t g0, int64_t lo, int64_t hi,
                                         uint32_t (&ops)[K]) {
    // g0: global index of this lane's first op (multiple of 4).  Ops outside [lo, hi) become "I, len 0"
    // (no effect on either cursor, not a clip).
    if (g0 + K <= lo || g0 >= hi) {
#pragma unroll
        for (int j = 0; j < K; j++) ops[j] = PAD_OP;
        return;
    }
    if ((uint64_t)(g0 + K) <= total && g0 >= 0) {
#pragma unroll
        for (int j = 0; j < K; j += 4) {
            uint4 t = __ldg(reinterpret_cast<const uint4 *>(cig + g0 + j));
            ops[j] = t.x; ops[j + 1] = t.y; ops[j + 2] = t.z; ops[j + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < K; j++) ops[j] = (g0 + j >= 0 && (uint64_t)(g0 + j) < total) ? cig[g0 + j] : PAD_OP;
    }
#pragma unroll
    for (int j = 0; j < K; j++) if (g0 + j < lo || g0 + j >= hi) ops[j] = PAD_OP;
}

template <int K>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_call_alleles(K1Args a) {
    __shared__ Cand s_cand[WARPS_PER_CTA][CAND_CAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wid = (long long)blockIdx.x * WARPS_PER_CTA + wib;
    const bool overflow_pass = a.overflow_reads != nullptr;
    int r;
    Cand *cand;
    int cand_cap;
    if (!overflow_pass) {
        if (wid >= a.b.n_reads) return;
        r = (int)wid;
        cand = s_cand[wib];
        cand_cap = CAND_CAP;
    } else {
        if (wid >= a.overflow_list_cap) return;
        r = (int)a.overflow_reads[wid];
        cand = a.overflow_buf + a.overflow_off[wid];
        cand_cap = (int)(a.overflow_off[wid + 1] - a.overflow_off[wid]);
    }
    const int nv = a.v.n;
    const int ref_start = a.b.ref_start[r];
    const int lq = a.b.l_qseq[r];
    const int ncig = (int)a.b.n_cigar[r];
    const int flag = a.b.flag[r];
    // iterator region "chr:1-lastSNP" (ParsingBam.cpp:1273) + read filter (:1282-1291)
    if (ref_start >= a.last_var_pos || (int)a.b.mapq[r] < a.mapping_quality || (flag & (0x4 | 0x100 | 0x400))) {
        if (lane == 0) { a.ncalls[r] = 0; a.tmp_start[r] = 0; a.status[r] = LPS_READ_FILTERED; }
        return;
    }
    const int32_t *__restrict__ vpos = a.v.pos;
    int cur = warp_lower_bound(vpos, nv, ref_start, lane);
    int prev_vp = cur > 0 ? vpos[cur - 1] : INT_MIN;
    int win_base = cur;
    int vwin = (win_base + lane < nv) ? vpos[win_base + lane] : INT_MAX;

    const int64_t lo = (int64_t)a.b.cigar_off[r], hi = lo + ncig;
    const int64_t abase = lo & ~(int64_t)3;
    constexpr int CH = 32 * K;
    const uint32_t *__restrict__ cig = a.b.cigar;

    int ref_pos = ref_start, qpos = 0;
    int ncand = 0;
    bool aborted = false, bad = false;

    uint32_t nxt_ops[K];
    load_ops<K>(cig, a.b.cigar_len, abase + (int64_t)lane * K, lo, hi, nxt_ops);
    for (int64_t cb = abase; cb < hi; cb += CH) {
        uint32_t ops[K];
#pragma unroll
        for (int j = 0; j < K; j++) ops[j] = nxt_ops[j];
        if (cb + CH < hi) load_ops<K>(cig, a.b.cigar_len, cb + CH + (int64_t)lane * K, lo, hi, nxt_ops);
        // per-lane advances
        int rs = 0, qs = 0;
        unsigned opmask = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            unsigned op = ops[j] & 15u;
            int len = (int)(ops[j] >> 4);
            rs += ((0x18Du >> op) & 1u) ? len : 0;   // M D N = X consume the reference
            qs += ((0x193u >> op) & 1u) ? len : 0;   // M I S = X consume the query
            opmask |= 1u << op;
        }
        int ri = rs, qi = qs;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(FULL, ri, d);
            int u = __shfl_up_sync(FULL, qi, d);
            if (lane >= d) { ri += t; qi += u; }
        }
        const int rtot = __shfl_sync(FULL, ri, 31), qtot = __shfl_sync(FULL, qi, 31);
        const int r0 = ref_pos + ri - rs, q0 = qpos + qi - qs;   // this lane's first op starts here
        const int chunk_end = ref_pos + rtot;
        int abort_op = INT_MAX;

        // ---- variants whose position falls inside this chunk ----
        while (true) {
            if (cur - win_base >= 32) {
                win_base = cur;
                vwin = (win_base + lane < nv) ? vpos[win_base + lane] : INT_MAX;
            }
            const int vp = __shfl_sync(FULL, vwin, cur - win_base);
            if (vp >= chunk_end) break;     // also covers cur == nv (INT_MAX)
            // owner lane: the one whose ops span vp
            const unsigned own = __ballot_sync(FULL, vp >= r0 && vp < r0 + rs);
            if (own == 0) { prev_vp = vp; cur++; continue; }   // cannot happen: the lanes tile [ref_pos, chunk_end)
            const int ol = __ffs(own) - 1;
            int o_op = 0, o_len = 0, o_r = 0, o_q = 0, o_idx = 0;
            if (lane == ol) {
                int rr = r0, qq = q0;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    unsigned op = ops[j] & 15u;
                    int len = (int)(ops[j] >> 4);
                    int radv = ((0x18Du >> op) & 1u) ? len : 0;
                    if (vp >= rr && vp < rr + radv) { o_op = (int)op; o_len = len; o_r = rr; o_q = qq; o_idx = j; }
                    rr += radv;
                    qq += ((0x193u >> op) & 1u) ? len : 0;
                }
            }
            o_op = __shfl_sync(FULL, o_op, ol); o_len = __shfl_sync(FULL, o_len, ol);
            o_r = __shfl_sync(FULL, o_r, ol);   o_q = __shfl_sync(FULL, o_q, ol);
            o_idx = __shfl_sync(FULL, o_idx, ol);
            const int64_t gidx = cb + (int64_t)ol * K + o_idx;   // global index of the covering op
            const int opi = (int)(gidx - lo);                     // CIGAR index inside the read

            if (o_op == 0 || o_op == 7 || o_op == 8) {
                const int off = vp - o_r;
                if (o_q + off + 1 > lq) { aborted = true; abort_op = opi; break; }            // :1453-1455
                const int rl = a.v.ref_len[cur], al = a.v.alt_len[cur];
                if (rl == 1 && al == 1) {
                    if (lane == 0 && ncand < cand_cap) { cand[ncand].var = cur; cand[ncand].x = (uint32_t)(o_q + off); }
                    ncand++;
                } else if ((rl == 1) != (al == 1)) {
                    if (opi + 1 < ncig) {                                                       // :1470, :1495
                        const unsigned nop = cig[gidx + 1] & 15u;
                        const unsigned want = (rl == 1) ? 1u : 2u;                             // I for an insertion, D for a deletion
                        const int allele = (o_r + o_len - 1 == vp && nop == want) ? 1 : 0;
                        if (lane == 0 && ncand < cand_cap) {
                            cand[ncand].var = cur;
                            cand[ncand].x = (2u << 30) | ((unsigned)allele << 1) | (unsigned)a.v.danger[cur];
                        }
                        ncand++;
                    }
                }
            } else if (o_op == 2) {
                // D-op rule (:1539-1607): only the first pending variant of the op, homopolymer >= 3
                if (a.have_reference && prev_vp < o_r && a.v.hom[cur] >= 3) {
                    if (o_q + 1 > lq) { aborted = true; abort_op = opi; break; }               // :1559-1561
                    const int rl = a.v.ref_len[cur], al = a.v.alt_len[cur];
                    if (rl == 1 && al == 1) {
                        if (lane == 0 && ncand < cand_cap) { cand[ncand].var = cur; cand[ncand].x = (1u << 30) | (uint32_t)o_q; }
                        ncand++;
                    } else if (rl != 1 && al == 1) {
                        if (lane == 0 && ncand < cand_cap) { cand[ncand].var = cur; cand[ncand].x = (2u << 30) | (1u << 2) | (1u << 1); }
                        ncand++;
                    }
                }
            }
            // N ops and variants deeper inside a D op are skipped by the reference's catch-up loop
            prev_vp = vp;
            cur++;
        }

        // ---- clips (S/H longer than 5) and unsupported ops ----
        if (__any_sync(FULL, (opmask & ~0x18Fu) != 0)) {
            int rr = r0;
#pragma unroll
            for (int j = 0; j < K; j++) {
                unsigned op = ops[j] & 15u;
                int len = (int)(ops[j] >> 4);
                int64_t g = cb + (int64_t)lane * K + j;
                if ((op == 4 || op == 5) && len > 5 && (int)(g - lo) < abort_op) {
                    unsigned long long slot = atomicAdd(&a.counters->clips, 1ull);
                    if (slot < a.clip_cap) a.clip_keys[slot] = ((uint32_t)rr << 1) | (g == lo ? 0u : 1u);
                }
                if (op > 8 && (int)(g - lo) < abort_op) bad = true;
                rr += ((0x18Du >> op) & 1u) ? len : 0;
            }
        }
        if (aborted) break;
        ref_pos += rtot; qpos += qtot;
    }
    bad = __any_sync(FULL, bad);
    if (bad && lane == 0) atomicAdd(&a.counters->bad_cigar, 1u);

    if (aborted || bad) {
        if (lane == 0) { a.ncalls[r] = 0; a.tmp_start[r] = 0; a.status[r] = LPS_READ_ABORTED; }
        return;
    }
    if (ncand > cand_cap) {
        // rare: more candidates than the shared buffer holds — redo this read in the overflow pass
        if (lane == 0) {
            unsigned k = atomicAdd(&a.counters->overflow_reads, 1u);
            atomicAdd(&a.counters->overflow_cands, (unsigned long long)ncand);
            if (k < a.overflow_list_cap) { a.overflow_list_out[k] = (uint32_t)r; a.overflow_need_out[k] = (uint64_t)ncand; }
            a.ncalls[r] = 0; a.tmp_start[r] = 0; a.status[r] = LPS_READ_OK;
        }
        return;
    }
    __syncwarp();

    // ---- resolve candidates: gather base + quality, decide the allele, drop filterSNP variants ----
    const uint8_t *__restrict__ seq = a.b.seq4 + a.b.seq_off[r];
    const uint8_t *__restrict__ qual = a.b.qual + a.b.qual_off[r];
    int nvalid = 0;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
        const int c = c0 + lane;
        bool valid = false;
        lps_call out;
        out.var = 0; out.quality = 0; out.allele = 0; out.origin = 0;
        if (c < ncand) {
            const Cand cd = cand[c];
            const unsigned kind = cd.x >> 30;
            out.var = cd.var;
            if (kind == 2) {
                out.allele = (int8_t)((cd.x >> 1) & 1u);
                out.origin = (int8_t)((cd.x >> 2) & 1u);
                out.quality = (cd.x & 1u) ? -5 : -4;
                valid = true;
            } else {
                const int qi = (int)(cd.x & 0x3fffffffu);
                const unsigned byte = seq[qi >> 1];
                const unsigned code = (byte >> ((~qi & 1) << 2)) & 0xfu;            // bam_seqi
                const char base = "=ACMGRSVTWYHKDBN"[code];                          // seq_nt16_str
                const char rb = (char)a.v.ref0[cd.var], ab = (char)a.v.alt0[cd.var];
                out.quality = (int16_t)qual[qi];
                out.origin = (int8_t)kind;
                if (base == rb) { out.allele = 0; valid = true; }
                else if (base == ab) { out.allele = 1; valid = true; }
            }
            if (valid && a.apply_filter && a.v.filtered[cd.var]) valid = false;
        }
        // compact inside the shared buffer (reuse it as the staging area for the final write)
        const unsigned m = __ballot_sync(FULL, valid);
        __syncwarp();
        if (valid) {
            const int dst = nvalid + __popc(m & ((1u << lane) - 1u));
            Cand packed;
            packed.var = out.var;
            packed.x = ((uint32_t)(uint16_t)out.quality) | ((uint32_t)(uint8_t)out.allele << 16) | ((uint32_t)(uint8_t)out.origin << 24);
            cand[dst] = packed;   // dst <= c, and all reads of this round happened before the __syncwarp
        }
        nvalid += __popc(m);
        __syncwarp();
    }
    unsigned long long start = 0;
    if (lane == 0 && nvalid) start = atomicAdd(&a.counters->tmp_calls, (unsigned long long)nvalid);
    start = __shfl_sync(FULL, start, 0);
    if (lane == 0) { a.ncalls[r] = (uint32_t)nvalid; a.tmp_start[r] = start; a.status[r] = LPS_READ_OK; }
    if (start + nvalid <= a.calls_cap) {
        for (int c = lane; c < nvalid; c += 32) {
            const Cand cd = cand[c];
            lps_call out;
            out.var = cd.var;
            out.quality = (int16_t)(cd.x & 0xffffu);
            out.allele = (int8_t)((cd.x >> 16) & 0xffu);
            out.origin = (int8_t)((cd.x >> 24) & 0xffu);
            a.calls_tmp[start + c] = out;
        }
    }
}

// scratch pool (allocation order) -> CSR in read order; one warp per read
__global__ void k_gather_calls(int n_reads, const uint64_t *__restrict__ tmp_start, const uint32_t *__restrict__ ncalls,
                               const uint64_t *__restrict__ call_off, const lps_call *__restrict__ tmp,
                               lps_call *__restrict__ out) {
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (wid >= n_reads) return;
    uint32_t n = ncalls[wid];
    uint64_t s = tmp_start[wid], d = call_off[wid];
    const uint64_t *src = reinterpret_cast<const uint64_t *>(tmp);
    uint64_t *dst = reinterpret_cast<uint64_t *>(out);
    for (uint32_t c = lane; c < n; c += 32) dst[d + c] = src[s + c];
}

__global__ void k_widen_u32(int n, const uint32_t *__restrict__ in, uint64_t *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

}  // namespace

static_assert(sizeof(lps_call) == 8, "lps_call must be 8 bytes");

int lps_launch_call_alleles(lps_ctx *ctx, const lps_phase_params *p) {
    const int n = ctx->batch.n_reads;
    const int nv = ctx->var.n;
    cudaStream_t st = ctx->stream;
    LPS_CUDA(ctx, ctx->d_tmp_start.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_call_off.reserve((size_t)n + 2));
    LPS_CUDA(ctx, ctx->d_ncalls.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_status.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_counters.reserve(1));
    // scratch pool capacity from the variant density of the contig; re-run on overflow
    double density = 0.0;
    if (nv > 1) density = (double)nv / ((double)ctx->h_vpos[nv - 1] - (double)ctx->h_vpos[0] + 1.0);
    size_t cap = (size_t)((double)ctx->sum_l_qseq * density * 1.5) + (size_t)n * 4 + 4096;
    if (cap < ctx->d_calls_tmp.cap) cap = ctx->d_calls_tmp.cap;
    size_t clip_cap = (size_t)n * 2 + 1024;
    if (clip_cap < ctx->d_clip_keys.cap) clip_cap = ctx->d_clip_keys.cap;
    const uint32_t ovf_cap = 1u << 16;
    LPS_CUDA(ctx, ctx->d_overflow_reads.reserve(ovf_cap));
    LPS_CUDA(ctx, ctx->d_overflow_cand.reserve(ovf_cap));

    CallCounters hc;
    for (int attempt = 0; attempt < 3; attempt++) {
        LPS_CUDA(ctx, ctx->d_calls_tmp.reserve(cap));
        LPS_CUDA(ctx, ctx->d_clip_keys.reserve(clip_cap));
        LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(CallCounters), st));
        K1Args a;
        memset(&a, 0, sizeof(a));
        a.b = ctx->batch; a.v = ctx->var;
        a.mapping_quality = p->mapping_quality; a.have_reference = p->have_reference && ctx->ref_len > 0;
        a.apply_filter = p->is_ont;
        a.last_var_pos = nv ? ctx->h_vpos[nv - 1] : -1;
        a.calls_tmp = ctx->d_calls_tmp.p; a.calls_cap = ctx->d_calls_tmp.cap;
        a.tmp_start = ctx->d_tmp_start.p; a.ncalls = ctx->d_ncalls.p; a.status = ctx->d_status.p;
        a.clip_keys = ctx->d_clip_keys.p; a.clip_cap = ctx->d_clip_keys.cap;
        a.counters = ctx->d_counters.p;
        a.overflow_list_out = ctx->d_overflow_reads.p; a.overflow_need_out = ctx->d_overflow_cand.p;
        a.overflow_list_cap = ovf_cap;
        const int grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        if (grid > 0) {
            k_call_alleles<8><<<grid, WARPS_PER_CTA * 32, 0, st>>>(a);
            ctx->stats.kernel_launches++;
        }
        LPS_CUDA(ctx, cudaGetLastError());
        LPS_CUDA(ctx, cudaMemcpyAsync(&hc, ctx->d_counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        LPS_CUDA(ctx, cudaStreamSynchronize(st));
        if (hc.bad_cigar) return ctx->fail(LPS_E_CIGAR, "alignment find unsupported CIGAR operation");
        if (hc.overflow_reads > ovf_cap) return ctx->fail(LPS_E_NOMEM, "too many reads overflow the candidate buffer");
        if (hc.overflow_reads) {
            // second pass for the few reads with more candidates than the shared buffer: same kernel,
            // candidate lists in global memory sized from the first pass
            std::vector<uint64_t> need(hc.overflow_reads), off(hc.overflow_reads + 1, 0);
            LPS_CUDA(ctx, cudaMemcpy(need.data(), ctx->d_overflow_cand.p, 8 * (size_t)hc.overflow_reads, cudaMemcpyDeviceToHost));
            for (uint32_t i = 0; i < hc.overflow_reads; i++) off[i + 1] = off[i] + need[i];
            LPS_CUDA(ctx, ctx->d_overflow_off.reserve(off.size()));
            LPS_CUDA(ctx, cudaMemcpy(ctx->d_overflow_off.p, off.data(), 8 * off.size(), cudaMemcpyHostToDevice));
            DevBuf<Cand> ovf;
            LPS_CUDA(ctx, ovf.reserve((size_t)off.back() + 1));
            K1Args b2 = a;
            b2.overflow_reads = ctx->d_overflow_reads.p; b2.overflow_off = ctx->d_overflow_off.p; b2.overflow_buf = ovf.p;
            b2.overflow_list_cap = hc.overflow_reads;
            // the overflow pass must not append the clips / counters of these reads a second time
            b2.clip_cap = 0;
            DevBuf<CallCounters> scratch;
            LPS_CUDA(ctx, scratch.reserve(1));
            LPS_CUDA(ctx, cudaMemcpyAsync(scratch.p, ctx->d_counters.p, sizeof(CallCounters), cudaMemcpyDeviceToDevice, st));
            b2.counters = scratch.p;
            const int g2 = ((int)hc.overflow_reads + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
            k_call_alleles<8><<<g2, WARPS_PER_CTA * 32, 0, st>>>(b2);
            ctx->stats.kernel_launches++;
            LPS_CUDA(ctx, cudaGetLastError());
            CallCounters h2;
            LPS_CUDA(ctx, cudaMemcpyAsync(&h2, scratch.p, sizeof(h2), cudaMemcpyDeviceToHost, st));
            LPS_CUDA(ctx, cudaStreamSynchronize(st));
            hc.tmp_calls = h2.tmp_calls;
            ovf.release(); scratch.release();
        }
        if (hc.tmp_calls <= ctx->d_calls_tmp.cap && hc.clips <= ctx->d_clip_keys.cap) break;
        cap = (size_t)hc.tmp_calls + 1024;
        clip_cap = (size_t)hc.clips + 1024;
        if (attempt == 2) return ctx->fail(LPS_E_NOMEM, "call pool sizing did not converge");
    }

    // CSR offsets in read order
    const int tb = 256;
    k_widen_u32<<<(n + tb) / tb, tb, 0, st>>>(n, ctx->d_ncalls.p, ctx->d_call_off.p + 0);
    ctx->stats.kernel_launches++;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->d_call_off.p, ctx->d_call_off.p, n + 1, st);
    LPS_CUDA(ctx, ctx->d_cub_tmp.reserve(tmp_bytes));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_call_off.p + n, 0, 8, st));
    cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, tmp_bytes, ctx->d_call_off.p, ctx->d_call_off.p, n + 1, st);
    ctx->stats.kernel_launches++;
    ctx->n_calls = hc.tmp_calls;
    LPS_CUDA(ctx, ctx->d_calls.reserve((size_t)ctx->n_calls + 1));
    if (n > 0) {
        const long long threads = (long long)n * 32;
        k_gather_calls<<<(unsigned)((threads + tb - 1) / tb), tb, 0, st>>>(n, ctx->d_tmp_start.p, ctx->d_ncalls.p, ctx->d_call_off.p,
                                                                         ctx->d_calls_tmp.p, ctx->d_calls.p);
        ctx->stats.kernel_launches++;
    }
    LPS_CUDA(ctx, cudaGetLastError());

    // clipCount map: sort the (pos << 1 | side) keys, run-length encode
    const int nclip = (int)hc.clips;
    ctx->h_clip_pos.clear(); ctx->h_clip_front.clear(); ctx->h_clip_back.clear();
    if (nclip > 0) {
        LPS_CUDA(ctx, ctx->d_clip_keys_sorted.reserve((size_t)nclip));
        LPS_CUDA(ctx, ctx->d_clip_unique.reserve((size_t)nclip));
        LPS_CUDA(ctx, ctx->d_clip_counts.reserve((size_t)nclip));
        LPS_CUDA(ctx, ctx->d_num_runs.reserve(1));
        size_t b1 = 0, b2 = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, b1, ctx->d_clip_keys.p, ctx->d_clip_keys_sorted.p, nclip, 0, 32, st);
        cub::DeviceRunLengthEncode::Encode(nullptr, b2, ctx->d_clip_keys_sorted.p, ctx->d_clip_unique.p, ctx->d_clip_counts.p,
                                           ctx->d_num_runs.p, nclip, st);
        LPS_CUDA(ctx, ctx->d_cub_tmp.reserve(b1 > b2 ? b1 : b2));
        cub::DeviceRadixSort::SortKeys(ctx->d_cub_tmp.p, b1, ctx->d_clip_keys.p, ctx->d_clip_keys_sorted.p, nclip, 0, 32, st);
        cub::DeviceRunLengthEncode::Encode(ctx->d_cub_tmp.p, b2, ctx->d_clip_keys_sorted.p, ctx->d_clip_unique.p, ctx->d_clip_counts.p,
                                           ctx->d_num_runs.p, nclip, st);
        ctx->stats.kernel_launches += 2;
        int32_t runs = 0;
        LPS_CUDA(ctx, cudaMemcpyAsync(&runs, ctx->d_num_runs.p, 4, cudaMemcpyDeviceToHost, st));
        LPS_CUDA(ctx, cudaStreamSynchronize(st));
        std::vector<uint32_t> keys((size_t)runs), cnts((size_t)runs);
        LPS_CUDA(ctx, cudaMemcpy(keys.data(), ctx->d_clip_unique.p, 4 * (size_t)runs, cudaMemcpyDeviceToHost));
        LPS_CUDA(ctx, cudaMemcpy(cnts.data(), ctx->d_clip_counts.p, 4 * (size_t)runs, cudaMemcpyDeviceToHost));
        ctx->stats.d2h_bytes += 8ull * (uint64_t)runs;
        for (int i = 0; i < runs; i++) {
            int32_t pos = (int32_t)(keys[i] >> 1);
            if (ctx->h_clip_pos.empty() || ctx->h_clip_pos.back() != pos) {
                ctx->h_clip_pos.push_back(pos); ctx->h_clip_front.push_back(0); ctx->h_clip_back.push_back(0);
            }
            if (keys[i] & 1u) ctx->h_clip_back.back() += (int32_t)cnts[i]; else ctx->h_clip_front.back() += (int32_t)cnts[i];
        }
    }
    ctx->have_calls = true;
    ctx->host_calls_valid = false;
    return LPS_OK;
}
