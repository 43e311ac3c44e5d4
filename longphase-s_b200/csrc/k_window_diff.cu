// k_window_diff.cu — the +-100 window comparison of a read with the reference around every covered tumor position:
// getWindowsDiffRef / getOrderWindowsDiffRef / processCigarOperation (reference src/somatic_haplotag/SomaticVarCaller.cpp:627-710),
// the hottest function of the tumor extract pass on the CPU (47 % of its non-I/O samples).
//
// The reference appends (offset, read_base) pairs to PosSomaticOffsetBase[allele]; its only consumer, the DenseAlt filter
// (:1160-1203), counts entries per offset.  The kernel therefore bins straight into window_hist[slot][allele][offset + 100].
//
// Mapping: ONE THREAD PER (alignment, tumor position) work item emitted by k_call_alleles' tumor dialect.
// The reference's scan is a sequential state machine with quirks (the budget is decremented BEFORE each step and the hop to the
// neighbouring CIGAR op happens at 0 or -1, so the backward scan skips the first base of every op; N / P / X ops consume
// iterations without moving; offsets are iteration indices, not base distances).  It is run as written, one iteration at a time:
// the 32 items of a warp walk their 100 + 100 iterations in lockstep, the hop (a few instructions, taken by a fifth of the lanes
// per iteration) is the only divergent part.  What made the scan slow on a GPU is where its operands live: a dependent 2-byte
// load per hop, a SEQ byte and a reference byte per step, all over the contig.  So the warp first stages, for each of its items,
// the neighbourhood of the tumor position in shared memory - 64 CIGAR ops around the covering op, 128 bytes of SEQ (256 bases)
// around the query index, 256 reference bases around the position: 32 units of 16 bytes, one coalesced 128-bit load per lane and
// item - and the scan runs on shared memory.  An access outside a window (a long deletion or insertion inside the +-100 window,
// more than 32 ops in one direction) falls back to global memory.
// History: one thread per item on global memory 0.61 ms for 302 k items; eight lanes per item 0.39 ms (198 GB/s of algorithmic
// bytes); eight lanes per item on staged windows 0.47 ms - four items per warp diverge at every segment, so every item paid for
// its own instruction stream; this layout: see profiles/.
#include <climits>
#include "lps_ctx.cuh"

namespace {

struct WdArgs {
    DevBatch b;
    const WdItem *items;
    unsigned long long n_items;
    const int32_t *vpos;        // variant positions
    const int32_t *tum_var;     // slot -> variant
    const char *ref;
    long long ref_len;          // 0 when the reference string is empty
    int32_t *window_hist;       // [n_tum][2][LPS_WINDOW_BINS]
};

constexpr int WD_TPB = 128;         // items (threads) per CTA
constexpr int WD_OPS = 64, WD_SEQ = 128, WD_REF = 256;
constexpr int WD_WORDS = (WD_OPS * 2 + WD_SEQ + WD_REF) / 4 + 1;   // 129 words per item: an odd stride keeps the items' windows on distinct banks

struct WdView {
    const DevBatch *b;
    const uint16_t *cig;            // the read's ops (global)
    uint64_t gop0;
    // each window holds the elements [x0 + lo, x0 + hi) of its array (the 16-byte units that lie inside the array); anything else is
    // read from global memory
    const uint16_t *s_ops; int op0, op_lo, op_hi;   // s_ops[k] = op op0 + k
    const uint8_t *seq; const uint8_t *s_seq; long long seq0; int seq_lo, seq_hi;   // s_seq[k] = byte seq0 + k of the read's SEQ
    const char *ref; const uint8_t *s_ref; long long ref0; int ref_lo, ref_hi; long long ref_len;
    __device__ __forceinline__ unsigned op_word(int ci) const {
        const int k = ci - op0;
        return (k >= op_lo && k < op_hi) ? (unsigned)s_ops[k] : (unsigned)cig[ci];
    }
    __device__ __forceinline__ int op_len(int ci, unsigned w) const {
        const unsigned len = w >> 4;
        if (len != 0xFFFu) return (int)len;
        uint32_t lo = 0, hi = b->n_long;
        const uint64_t gop = gop0 + (uint64_t)ci;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (b->long_at[mid] < gop) lo = mid + 1; else hi = mid;
        }
        return (lo < b->n_long && b->long_at[lo] == gop) ? (int)b->long_len[lo] : 0xFFF;
    }
    __device__ __forceinline__ char base(int rp) const {
        const long long byte = (long long)(rp >> 1) - seq0;
        const unsigned v = (byte >= seq_lo && byte < seq_hi) ? (unsigned)s_seq[byte] : (unsigned)seq[rp >> 1];
        return "=ACMGRSVTWYHKDBN"[(v >> ((~rp & 1) << 2)) & 0xfu];
    }
    __device__ __forceinline__ char ref_at(int fp) const {
        if ((long long)fp == ref_len) return '\0';                             // std::string::operator[](size())
        const long long k = (long long)fp - ref0;
        return (k >= ref_lo && k < ref_hi) ? (char)s_ref[k] : ref[fp];
    }
};

// processCigarOperation (:627-654)
__device__ __forceinline__ bool next_op(const WdView &v, int &ci, int ci_end, int dir, int &remaining, int &read_pos, int &ref_pos, int &op) {
    ci += dir;
    while (ci < ci_end && ci >= 0) {
        const unsigned w = v.op_word(ci);
        op = (int)(w & 15u);
        const int len = v.op_len(ci, w);
        if (op == 0 || op == 3 || op == 6 || op == 7 || op == 8) { remaining += len; return true; }
        else if (op == 1) read_pos += len * dir;
        else if (op == 2) ref_pos += len * dir;
        else return false;
        ci += dir;
    }
    return false;
}

// getOrderWindowsDiffRef (:655-686), one iteration at a time.  Every iteration decrements the budget first; a result of 0 or -1 hops to
// the neighbouring op (whose length is added to the budget) before the iteration is executed there; N / P / X ops consume the iteration
// without moving; a moving op steps both cursors and compares; the scan ends at the first position out of range (read == l_qseq is
// one past SEQ, undefined in the reference: ends the scan as well).
__device__ __forceinline__ void scan(const WdView &v, int ci, const int ncig, const int lq, int read_pos, int remaining, int ref_pos, const int dir,
                                     int32_t *__restrict__ hist, const bool live0) {
    int op = (int)(v.op_word(ci) & 15u);
    bool live = live0;
    for (int i = 1; i <= LPS_WINDOW; i++) {
        if (!__any_sync(0xffffffffu, live)) break;
        if (!live) continue;
        const bool hop = remaining == 1 || remaining == 0;
        remaining -= 1;
        if (hop && !next_op(v, ci, ncig, dir, remaining, read_pos, ref_pos, op)) { live = false; continue; }
        if (op == 2 || op == 1 || op == 3 || op == 6 || op == 8) continue;
        const int rp = read_pos + dir, fp = ref_pos + dir;
        bool ok;
        if (dir > 0) ok = rp <= lq - 1 && (long long)fp <= v.ref_len;
        else ok = !(read_pos > lq || (long long)ref_pos > v.ref_len + 1) && rp >= 0 && fp >= 0;
        if (!ok) { live = false; continue; }
        if (v.base(rp) != v.ref_at(fp)) atomicAdd(hist + i * dir + LPS_WINDOW, 1);
        read_pos = rp; ref_pos = fp;
    }
}

__global__ void __launch_bounds__(WD_TPB) k_window_diff(WdArgs a) {
    extern __shared__ __align__(16) uint32_t s_win[];          // [WD_TPB][WD_WORDS]
    __shared__ unsigned long long s_g0[WD_TPB][3];             // per item: first byte (global address) of its three windows
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, tid = threadIdx.x;
    const bool have = t < a.n_items;
    WdItem it;
    it.read = 0; it.slot2 = 0; it.opi = 0; it.qidx = 0; it.off = 0;
    if (have) it = a.items[t];
    const int r = (int)it.read;
    WdView v;
    v.b = &a.b;
    v.gop0 = have ? a.b.cigar_off[r] : 0;
    v.cig = a.b.cigar16 + v.gop0;
    v.seq = a.b.seq4 + (have ? a.b.seq_off[r] : 0);
    const int ncig = have ? (int)a.b.n_cigar[r] : 0, lq = have ? a.b.l_qseq[r] : 0;
    const int ci = (int)it.opi, off = (int)it.off, qidx = (int)it.qidx;
    const int var_pos = have ? a.vpos[a.tum_var[it.slot2 >> 1]] : 0;
    v.ref = a.ref; v.ref_len = a.ref_len;
    // ---- the three windows: 16-byte units aligned in the GLOBAL address space (a window may start a few elements before the wanted
    //      position; whatever lies there - the previous read's ops or bases - is never looked at).  A unit is loaded only when it
    //      lies inside [array start rounded up, array end rounded DOWN to 16 bytes): no byte outside the caller's arrays is touched.
    const uintptr_t ops_base = ((uintptr_t)a.b.cigar16 + 15u) & ~(uintptr_t)15, ops_lim = ((uintptr_t)a.b.cigar16 + (uintptr_t)a.b.cigar_len * 2u) & ~(uintptr_t)15;
    const uintptr_t seq_base = ((uintptr_t)a.b.seq4 + 15u) & ~(uintptr_t)15, seq_lim = ((uintptr_t)a.b.seq4 + (uintptr_t)a.b.seq_bytes) & ~(uintptr_t)15;
    const uintptr_t ref_base = ((uintptr_t)a.ref + 15u) & ~(uintptr_t)15, ref_lim = ((uintptr_t)a.ref + (uintptr_t)(a.ref_len > 0 ? a.ref_len : 0)) & ~(uintptr_t)15;
    const uintptr_t g_ops = ((uintptr_t)(v.cig + ci) - (uintptr_t)WD_OPS) & ~(uintptr_t)15;                 // WD_OPS / 2 ops before the covering op
    const uintptr_t g_seq = ((uintptr_t)(v.seq + (qidx >> 1)) - (uintptr_t)(WD_SEQ / 2)) & ~(uintptr_t)15;
    const uintptr_t g_ref = ((uintptr_t)(a.ref + var_pos) - (uintptr_t)(WD_REF / 2)) & ~(uintptr_t)15;
    s_g0[tid][0] = g_ops; s_g0[tid][1] = g_seq; s_g0[tid][2] = g_ref;
    __syncwarp();
    {
        // lane l loads unit l of every item of its warp: units 0..7 ops, 8..15 SEQ, 16..31 reference
        const int which = lane < 8 ? 0 : lane < 16 ? 1 : 2;
        const int unit = lane < 8 ? lane : lane < 16 ? lane - 8 : lane - 16;
        const uintptr_t base = which == 0 ? ops_base : which == 1 ? seq_base : ref_base, lim = which == 0 ? ops_lim : which == 1 ? seq_lim : ref_lim;
        const int dst_word = (which == 0 ? 0 : which == 1 ? WD_OPS * 2 / 4 : (WD_OPS * 2 + WD_SEQ) / 4) + unit * 4;
        const int w0 = tid & ~31;
        for (int k = 0; k < 32; k++) {
            if ((unsigned long long)blockIdx.x * blockDim.x + (unsigned)(w0 + k) >= a.n_items) break;
            const uintptr_t addr = (uintptr_t)s_g0[w0 + k][which] + 16u * (unsigned)unit;
            if (addr >= base && addr + 16 <= lim) {
                const uint4 q = *reinterpret_cast<const uint4 *>(addr);
                uint32_t *d = s_win + (size_t)(w0 + k) * WD_WORDS + dst_word;
                d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
            }
        }
    }
    __syncwarp();
    auto valid = [](uintptr_t g0, uintptr_t base, uintptr_t lim, int bytes, int &lo, int &hi) {
        const long long l = (long long)base - (long long)g0, h = (long long)lim - (long long)g0;
        lo = (int)(l < 0 ? 0 : (l > bytes ? bytes : l));
        hi = (int)(h < 0 ? 0 : (h > bytes ? bytes : h));
        if (hi < lo) hi = lo;
    };
    const uint8_t *mine = reinterpret_cast<const uint8_t *>(s_win + (size_t)tid * WD_WORDS);
    v.s_ops = reinterpret_cast<const uint16_t *>(mine); v.s_seq = mine + WD_OPS * 2; v.s_ref = mine + WD_OPS * 2 + WD_SEQ;
    valid(g_ops, ops_base, ops_lim, WD_OPS * 2, v.op_lo, v.op_hi);
    v.op0 = (int)(((long long)g_ops - (long long)(uintptr_t)v.cig) / 2); v.op_lo = (v.op_lo + 1) / 2; v.op_hi /= 2;      // bytes -> ops
    valid(g_seq, seq_base, seq_lim, WD_SEQ, v.seq_lo, v.seq_hi);
    v.seq0 = (long long)g_seq - (long long)(uintptr_t)v.seq;
    valid(g_ref, ref_base, ref_lim, WD_REF, v.ref_lo, v.ref_hi);
    v.ref0 = (long long)g_ref - (long long)(uintptr_t)a.ref;
    int32_t *hist = a.window_hist + (size_t)it.slot2 * LPS_WINDOW_BINS;
    // getWindowsDiffRef (:688-710): the op is an M/=/X op, never an insertion
    const int oplen = have ? v.op_len(ci, v.op_word(ci)) : 0;
    const int fwd = oplen - off > 0 ? oplen - off : 0, rev = off > 0 ? off : 0;
    scan(v, ci, ncig, lq, qidx, rev, var_pos, -1, hist, have);
    scan(v, ci, ncig, lq, qidx, fwd, var_pos, 1, hist, have);
}

}  // namespace

int lps_launch_window_diff(lps_ctx *ctx, int have_reference) {
    const unsigned long long n = ctx->n_wd_items;
    ctx->stats.ms_kernel_window_diff = 0.f;
    if (n == 0) return LPS_OK;
    WdArgs a;
    a.b = ctx->batch; a.items = ctx->d_wd_items.p; a.n_items = n; a.vpos = ctx->var.pos; a.tum_var = ctx->som.tum_var;
    a.ref = ctx->d_ref.p; a.ref_len = have_reference ? (long long)ctx->ref_len : 0; a.window_hist = ctx->som.window_hist;
    const int tb = WD_TPB;
    const size_t smem = (size_t)WD_TPB * WD_WORDS * sizeof(uint32_t);
    static thread_local int prepared_device = -1;
    if (prepared_device != ctx->device) {            // the opt-in belongs to the device, like k_call_alleles'
        LPS_CUDA(ctx, cudaFuncSetAttribute(k_window_diff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prepared_device = ctx->device;
    }
    cudaEventRecord(ctx->kev[4], ctx->stream);
    k_window_diff<<<(unsigned)((n + tb - 1) / tb), tb, smem, ctx->stream>>>(a);   // one thread per item
    cudaEventRecord(ctx->kev[5], ctx->stream);
    ctx->stats.kernel_launches++;
    LPS_CUDA(ctx, cudaGetLastError());
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.ms_kernel_window_diff, ctx->kev[4], ctx->kev[5]);
    return LPS_OK;
}
