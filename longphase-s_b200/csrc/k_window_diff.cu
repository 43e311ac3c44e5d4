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
    const uint16_t *s_ops; int op0;                 // s_ops[k] = op op0 + k
    const uint8_t *seq; const uint8_t *s_seq; long long seq0;   // s_seq[k] = seq byte seq0 + k (byte index inside the read's SEQ)
    const char *ref; const uint8_t *s_ref; long long ref0; long long ref_len;
    __device__ __forceinline__ unsigned op_word(int ci) const {
        const unsigned k = (unsigned)(ci - op0);
        return k < (unsigned)WD_OPS ? (unsigned)s_ops[k] : (unsigned)cig[ci];
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
        const unsigned v = (unsigned long long)byte < (unsigned long long)WD_SEQ ? (unsigned)s_seq[byte] : (unsigned)seq[rp >> 1];
        return "=ACMGRSVTWYHKDBN"[(v >> ((~rp & 1) << 2)) & 0xfu];
    }
    __device__ __forceinline__ char ref_at(int fp) const {
        if ((long long)fp == ref_len) return '\0';                             // std::string::operator[](size())
        const long long k = (long long)fp - ref0;
        return (unsigned long long)k < (unsigned long long)WD_REF ? (char)s_ref[k] : ref[fp];
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
    // ---- stage the three windows (each lane its share, the item's eight lanes are consecutive lanes of one warp) ----
    v.op0 = ci - WD_OPS / 2;
    for (int k = sub; k < WD_OPS; k += 8) {
        const int idx = v.op0 + k;
        s_ops[slot][k] = (idx >= 0 && idx < ncig) ? v.cig[idx] : (uint16_t)0;
    }
    const long long seq_bytes = ((long long)lq + 1) >> 1;
    v.seq0 = (long long)(qidx >> 1) - WD_SEQ / 2;
    for (int k = sub; k < WD_SEQ; k += 8) {
        const long long byte = v.seq0 + k;
        s_seq[slot][k] = (byte >= 0 && byte < seq_bytes) ? v.seq[byte] : (uint8_t)0;
    }
    v.ref0 = (long long)var_pos - WD_REF / 2;
    for (int k = sub; k < WD_REF; k += 8) {
        const long long fp = v.ref0 + k;
        s_ref[slot][k] = (fp >= 0 && fp < a.ref_len) ? (uint8_t)a.ref[fp] : (uint8_t)0;
    }
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
