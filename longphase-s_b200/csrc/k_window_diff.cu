// k_window_diff.cu — the +-100 window comparison of a read with the reference around every covered tumor position:
// getWindowsDiffRef / getOrderWindowsDiffRef / processCigarOperation (reference src/somatic_haplotag/SomaticVarCaller.cpp:627-710),
// the hottest function of the tumor extract pass on the CPU (47 % of its non-I/O samples).
//
// The reference appends (offset, read_base) pairs to PosSomaticOffsetBase[allele]; its only consumer, the DenseAlt filter
// (:1160-1203), counts entries per offset.  The kernel therefore bins straight into window_hist[slot][allele][offset + 100].
//
// Mapping: EIGHT LANES PER (alignment, tumor position) work item emitted by k_call_alleles' tumor dialect, four items per warp.
// The reference's scan is a sequential state machine with quirks (the budget is decremented BEFORE each step and the hop to the
// neighbouring CIGAR op happens at 0 or -1, so the backward scan skips the first base of every op; N / P / X ops consume
// iterations without moving; offsets are iteration indices, not base distances).  Its control flow only changes at hops, so the
// scan is cut into SEGMENTS between hops: the (cheap, sequential) hop logic is replayed redundantly by the eight lanes, and the
// iterations of a segment - consecutive read / reference bases - are compared eight at a time with coalesced byte loads.  The
// first version used one thread per item and was bound by LSU wavefronts (every lane reading its own read): 0.61 ms for 302 k
// items; this layout needs ~7x fewer wavefronts.
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

// Per work item the neighbourhood of the tumor position is staged in shared memory once - 64 CIGAR ops around the covering op,
// 128 bytes of SEQ (256 bases) around the query index, 320 reference bases around the position - by the item's eight lanes with
// coalesced loads; the scan (hop logic and base comparison) then runs on shared memory.  Before, every hop waited for a dependent
// 2-byte load from L2 and every segment for its SEQ / reference bytes: ~40 serialised L2 round trips per item.  An access outside
// a window (a long deletion or insertion inside the +-100 window, more than 32 ops in one direction) falls back to global memory.
constexpr int WD_ITEMS = 16;        // items per CTA (8 lanes each)
constexpr int WD_OPS = 64, WD_SEQ = 128, WD_REF = 320;

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

// getOrderWindowsDiffRef (:655-686), segment by segment.  `remaining` is the budget BEFORE the decrement of iteration i.
__device__ __forceinline__ void scan(const WdView &v, int ci, int ncig, int lq, int read_pos, int remaining, int ref_pos, const int dir,
                                     int32_t *__restrict__ hist, const int sub) {
    int op = (int)(v.op_word(ci) & 15u);
    int i = 1;
    while (i <= LPS_WINDOW) {
        int first = 0;
        if (remaining == 1 || remaining == 0) {            // the decrement of iteration i gives 0 or -1: hop before executing it
            remaining -= 1;
            if (!next_op(v, ci, ncig, dir, remaining, read_pos, ref_pos, op)) return;
            first = 1;                                     // iteration i runs in the new op without another decrement
        }
        // iterations that follow without a hop: until the decrement gives 0; a negative budget never hops again
        const int extra = remaining >= 2 ? remaining - 1 : (remaining < 0 ? LPS_WINDOW : 0);
        int run = first + extra;
        if (run > LPS_WINDOW + 1 - i) run = LPS_WINDOW + 1 - i;
        if (!(op == 2 || op == 1 || op == 3 || op == 6 || op == 8)) {
            // a moving op: iteration i + t compares read[read_pos + dir (t+1)] with ref[ref_pos + dir (t+1)]; the scan ends at the
            // first position out of range (read == l_qseq is one past SEQ, undefined in the reference: ends the scan as well)
            int n_ok;
            if (dir > 0) n_ok = min(lq - 1 - read_pos, (int)min((long long)INT_MAX, v.ref_len - (long long)ref_pos));
            else n_ok = (read_pos > lq || (long long)ref_pos > v.ref_len + 1) ? 0 : min(read_pos, ref_pos);
            if (n_ok < 0) n_ok = 0;
            const int n = min(run, n_ok);
            for (int t = sub; t < n; t += 8) {
                const int rp = read_pos + dir * (t + 1), fp = ref_pos + dir * (t + 1);
                if (v.base(rp) != v.ref_at(fp)) atomicAdd(hist + (i + t) * dir + LPS_WINDOW, 1);
            }
            if (n < run) return;
            read_pos += dir * n; ref_pos += dir * n;
        }
        remaining -= run - first;
        i += run;
    }
}

__global__ void __launch_bounds__(WD_ITEMS * 8) k_window_diff(WdArgs a) {
    __shared__ __align__(16) uint16_t s_ops[WD_ITEMS][WD_OPS];
    __shared__ __align__(16) uint8_t s_seq[WD_ITEMS][WD_SEQ];
    __shared__ __align__(16) uint8_t s_ref[WD_ITEMS][WD_REF];
    const unsigned long long t = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7, slot = threadIdx.x >> 3;
    if (t >= a.n_items) return;        // whole groups of eight lanes leave together; no block-wide barrier below
    const WdItem it = a.items[t];
    const int r = (int)it.read;
    WdView v;
    v.b = &a.b;
    v.gop0 = a.b.cigar_off[r];
    v.cig = a.b.cigar16 + v.gop0;
    v.seq = a.b.seq4 + a.b.seq_off[r];
    const int ncig = (int)a.b.n_cigar[r], lq = a.b.l_qseq[r];
    const int ci = (int)it.opi, off = (int)it.off, qidx = (int)it.qidx;
    const int var_pos = a.vpos[a.tum_var[it.slot2 >> 1]];
    v.ref = a.ref; v.ref_len = a.ref_len;
    // ---- stage the three windows: 16-byte units, aligned in the GLOBAL address space (so a window may start a few elements before
    //      the wanted position: whatever lies there - the previous read's ops or bases - is never looked at); units that would
    //      reach outside the arrays stay zero ----
    // a unit is loaded when it lies inside [base rounded up, end rounded DOWN to 16 bytes): no byte outside the caller's arrays is touched
    auto stage = [&](const void *array, uint64_t n_bytes, const void *want_ptr, int window_bytes, uint8_t *dst, long long &x0_bytes, int &lo, int &hi) {
        const uintptr_t base = ((uintptr_t)array + 15u) & ~(uintptr_t)15, lim = ((uintptr_t)array + (uintptr_t)n_bytes) & ~(uintptr_t)15;
        const uintptr_t g0 = ((uintptr_t)want_ptr - (uintptr_t)(window_bytes / 2)) & ~(uintptr_t)15;
        x0_bytes = (long long)g0;
        for (int u = sub; u < window_bytes / 16; u += 8) {
            const uintptr_t addr = g0 + 16u * (unsigned)u;
            if (addr >= base && addr + 16 <= lim) reinterpret_cast<uint4 *>(dst)[u] = *reinterpret_cast<const uint4 *>(addr);
        }
        const long long l = (long long)base - (long long)g0, h = (long long)lim - (long long)g0;
        lo = (int)(l < 0 ? 0 : (l > window_bytes ? window_bytes : l));
        hi = (int)(h < 0 ? 0 : (h > window_bytes ? window_bytes : h));
        if (hi < lo) hi = lo;
    };
    long long x0;
    stage(a.b.cigar16, a.b.cigar_len * 2ull, v.cig + ci, WD_OPS * 2, reinterpret_cast<uint8_t *>(s_ops[slot]), x0, v.op_lo, v.op_hi);
    v.op0 = (int)((x0 - (long long)(uintptr_t)v.cig) / 2); v.op_lo = (v.op_lo + 1) / 2; v.op_hi /= 2;      // bytes -> ops
    stage(a.b.seq4, a.b.seq_bytes, v.seq + (qidx >> 1), WD_SEQ, s_seq[slot], x0, v.seq_lo, v.seq_hi);
    v.seq0 = x0 - (long long)(uintptr_t)v.seq;
    stage(a.ref, (uint64_t)(a.ref_len > 0 ? a.ref_len : 0), a.ref + var_pos, WD_REF, s_ref[slot], x0, v.ref_lo, v.ref_hi);
    v.ref0 = x0 - (long long)(uintptr_t)a.ref;
    v.s_ops = s_ops[slot]; v.s_seq = s_seq[slot]; v.s_ref = s_ref[slot];
    __syncwarp();
    int32_t *hist = a.window_hist + (size_t)it.slot2 * LPS_WINDOW_BINS;
    // getWindowsDiffRef (:688-710): the op is an M/=/X op, never an insertion
    const int oplen = v.op_len(ci, v.op_word(ci));
    const int fwd = oplen - off > 0 ? oplen - off : 0, rev = off > 0 ? off : 0;
    scan(v, ci, ncig, lq, qidx, rev, var_pos, -1, hist, sub);
    scan(v, ci, ncig, lq, qidx, fwd, var_pos, 1, hist, sub);
}

}  // namespace

int lps_launch_window_diff(lps_ctx *ctx, int have_reference) {
    const unsigned long long n = ctx->n_wd_items;
    ctx->stats.ms_kernel_window_diff = 0.f;
    if (n == 0) return LPS_OK;
    WdArgs a;
    a.b = ctx->batch; a.items = ctx->d_wd_items.p; a.n_items = n; a.vpos = ctx->var.pos; a.tum_var = ctx->som.tum_var;
    a.ref = ctx->d_ref.p; a.ref_len = have_reference ? (long long)ctx->ref_len : 0; a.window_hist = ctx->som.window_hist;
    const int tb = WD_ITEMS * 8;
    cudaEventRecord(ctx->kev[4], ctx->stream);
    k_window_diff<<<(unsigned)((8 * n + tb - 1) / tb), tb, 0, ctx->stream>>>(a);   // 8 lanes per item
    cudaEventRecord(ctx->kev[5], ctx->stream);
    ctx->stats.kernel_launches++;
    LPS_CUDA(ctx, cudaGetLastError());
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.ms_kernel_window_diff, ctx->kev[4], ctx->kev[5]);
    return LPS_OK;
}
