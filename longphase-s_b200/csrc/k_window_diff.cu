// k_window_diff.cu — the +-100 window comparison of a read with the reference around every covered tumor position:
// getWindowsDiffRef / getOrderWindowsDiffRef / processCigarOperation (reference src/somatic_haplotag/SomaticVarCaller.cpp:627-710),
// the hottest function of the tumor extract pass on the CPU (47 % of its non-I/O samples).
//
// The reference appends (offset, read_base) pairs to PosSomaticOffsetBase[allele]; its only consumer, the DenseAlt filter
// (:1160-1203), counts entries per offset.  The kernel therefore bins straight into window_hist[slot][allele][offset + 100].
//
// Mapping: ONE LANE PER (work item, direction), sixteen items per warp (lanes 0-15 scan backwards, lanes 16-31 forwards), on data
// the warp staged in shared memory with coalesced 16-byte loads: per item 64 CIGAR ops around the covering op, 256 read bases
// (already turned into the reference's characters) around the query index and 320 reference bases around the position.
// The reference's scan is a sequential state machine with quirks (the budget is decremented BEFORE each step and the hop to the
// neighbouring CIGAR op happens at 0 or -1, so the backward scan skips the first base of every op; N / P / X ops consume
// iterations without moving; offsets are iteration indices, not base distances).  A lane replays it in TRIPS: a trip visits at most
// one CIGAR op of a hop and then compares up to WD_CHUNK bases of the current run, so that the lanes of a warp - whose runs have
// different lengths - stay in one straight-line loop body instead of diverging into nested loops.
// History: one thread per item on global memory 0.61 ms; eight lanes per item, hop logic replayed by all eight, 0.47 ms for 302 k
// items (389 M warp instructions, 62 % of them in the range-checked per-base accessors); this layout: see profiles/.
#include <climits>
#include "lps_ctx.cuh"
#include "lps_async.cuh"

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

constexpr int WD_WARPS = 4, WD_ITEMS = 16;       // warps per CTA, items per warp
constexpr int WD_OPS = 64, WD_SEQ = 128, WD_REF = 256;
#ifndef LPS_WD_CHUNK
#define LPS_WD_CHUNK 8
#endif
constexpr int WD_CHUNK = LPS_WD_CHUNK;

struct WdWarp {
    uint4 ops[WD_ITEMS][WD_OPS * 2 / 16];        // ops[j] as uint16_t[WD_OPS]: op (op0 + k) of item j
    uint4 rd[WD_ITEMS][WD_SEQ * 2 / 16];         // characters of the read bases (two per staged SEQ byte)
    uint4 rf[WD_ITEMS][WD_REF / 16];             // reference characters
    unsigned long long g[WD_ITEMS][3];           // global byte address of each window's first 16-byte unit (0: item absent)
};

// where a window of `window_bytes` around want_ptr starts (16-byte aligned in the GLOBAL address space, so it may begin a few
// elements before the wanted one) and which of its bytes [lo, hi) lie in units that are completely inside the array
__device__ __forceinline__ unsigned long long window_of(const void *array, uint64_t n_bytes, const void *want_ptr, int before_bytes, int window_bytes, int &lo, int &hi) {
    const uintptr_t base = ((uintptr_t)array + 15u) & ~(uintptr_t)15, lim = ((uintptr_t)array + (uintptr_t)n_bytes) & ~(uintptr_t)15;
    const uintptr_t g0 = ((uintptr_t)want_ptr - (uintptr_t)before_bytes) & ~(uintptr_t)15;      // the wanted element sits before_bytes .. before_bytes + 15 bytes into the window
    const long long l = (long long)base - (long long)g0, h = (long long)lim - (long long)g0;
    lo = (int)(l < 0 ? 0 : (l > window_bytes ? window_bytes : l));
    hi = (int)(h < 0 ? 0 : (h > window_bytes ? window_bytes : h));
    if (hi < lo) hi = lo;
    return (unsigned long long)g0;
}

__global__ void __launch_bounds__(WD_WARPS * 32) k_window_diff(WdArgs a) {
    __shared__ WdWarp s_w[WD_WARPS];
    __shared__ uint16_t s_two[256];              // SEQ byte -> its two bases as characters (first base in the low byte)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_two[i] = (uint16_t)((unsigned)(unsigned char)"=ACMGRSVTWYHKDBN"[i >> 4] | ((unsigned)(unsigned char)"=ACMGRSVTWYHKDBN"[i & 15] << 8));
    __syncthreads();
    WdWarp &S = s_w[wib];
    const int j = lane & (WD_ITEMS - 1), dir = lane < WD_ITEMS ? -1 : 1;
    const unsigned long long t = ((unsigned long long)blockIdx.x * WD_WARPS + wib) * WD_ITEMS + (unsigned long long)j;
    const bool present = t < a.n_items;
    WdItem it = {0u, 0u, 0u, 0u, 0u};
    if (present) it = a.items[t];
    const int r = (int)it.read;
    const uint64_t gop0 = present ? a.b.cigar_off[r] : 0;
    const uint16_t *cig = a.b.cigar16 + gop0;
    const uint8_t *seq = a.b.seq4 + (present ? a.b.seq_off[r] : 0);
    const int ncig = present ? (int)a.b.n_cigar[r] : 0, lq = present ? a.b.l_qseq[r] : 0;
    const int ci0 = (int)it.opi, off = (int)it.off, qidx = (int)it.qidx;
    const int var_pos = present ? a.vpos[a.tum_var[it.slot2 >> 1]] : 0;
    // ---- the three windows of the item (both lanes of an item compute the same; the backward lane publishes the addresses) ----
    int op_lo, op_hi, sq_lo, sq_hi, rf_lo, rf_hi;
    const unsigned long long g_ops = window_of(a.b.cigar16, a.b.cigar_len * 2ull, cig + ci0, WD_OPS - 8, WD_OPS * 2, op_lo, op_hi);
    const unsigned long long g_seq = window_of(a.b.seq4, a.b.seq_bytes, seq + (qidx >> 1), WD_SEQ / 2 - 8, WD_SEQ, sq_lo, sq_hi);
    const unsigned long long g_ref = window_of(a.ref, (uint64_t)(a.ref_len > 0 ? a.ref_len : 0), a.ref + var_pos, WD_REF / 2 - 8, WD_REF, rf_lo, rf_hi);
    const int op0 = (int)(((long long)g_ops - (long long)(uintptr_t)cig) / 2);            // S.ops[j][k] = op op0 + k
    op_lo = (op_lo + 1) / 2; op_hi /= 2;                                                 // bytes -> ops
    const int rd0 = (int)(2 * ((long long)g_seq - (long long)(uintptr_t)seq));            // S.rd[j][k] = base rd0 + k of the read
    sq_lo *= 2; sq_hi *= 2;                                                              // bytes -> bases
    const int rf0 = (int)((long long)g_ref - (long long)(uintptr_t)a.ref);                // S.rf[j][k] = reference base rf0 + k
    if (lane < WD_ITEMS) {
        S.g[j][0] = present ? g_ops : 0ull; S.g[j][1] = present ? g_seq : 0ull; S.g[j][2] = present ? g_ref : 0ull;
    }
    __syncwarp();
    {
        // units that are not completely inside their array are skipped (and never read back: the [lo, hi) ranges exclude them)
        const uintptr_t ops_b = ((uintptr_t)a.b.cigar16 + 15u) & ~(uintptr_t)15, ops_l = ((uintptr_t)a.b.cigar16 + (uintptr_t)(a.b.cigar_len * 2ull)) & ~(uintptr_t)15;
        // ops and reference go straight to shared memory (cp.async, all in flight at once); SEQ passes through registers for the
        // nibble -> character table
#pragma unroll
        for (int q = lane; q < WD_ITEMS * (WD_OPS * 2 / 16); q += 32) {
            const int jj = q / (WD_OPS * 2 / 16), u = q % (WD_OPS * 2 / 16);
            const uintptr_t addr = (uintptr_t)S.g[jj][0] + 16u * (unsigned)u;
            if (addr >= ops_b && addr + 16 <= ops_l) cp_async16(smem_u32(&S.ops[jj][u]), reinterpret_cast<const void *>(addr));
        }
        const uintptr_t ref_b = ((uintptr_t)a.ref + 15u) & ~(uintptr_t)15, ref_l = ((uintptr_t)a.ref + (uintptr_t)(a.ref_len > 0 ? a.ref_len : 0)) & ~(uintptr_t)15;
#pragma unroll
        for (int q = lane; q < WD_ITEMS * (WD_REF / 16); q += 32) {
            const int jj = q / (WD_REF / 16), u = q % (WD_REF / 16);
            const uintptr_t addr = (uintptr_t)S.g[jj][2] + 16u * (unsigned)u;
            if (addr >= ref_b && addr + 16 <= ref_l) cp_async16(smem_u32(&S.rf[jj][u]), reinterpret_cast<const void *>(addr));
        }
        cp_async_commit();
        const uintptr_t seq_b = ((uintptr_t)a.b.seq4 + 15u) & ~(uintptr_t)15, seq_l = ((uintptr_t)a.b.seq4 + (uintptr_t)a.b.seq_bytes) & ~(uintptr_t)15;
        for (int q = lane; q < WD_ITEMS * (WD_SEQ / 16); q += 32) {
            const int jj = q / (WD_SEQ / 16), u = q % (WD_SEQ / 16);
            const uintptr_t addr = (uintptr_t)S.g[jj][1] + 16u * (unsigned)u;
            if (addr >= seq_b && addr + 16 <= seq_l) {
                const uint4 x = *reinterpret_cast<const uint4 *>(addr);
                const unsigned w[4] = {x.x, x.y, x.z, x.w};
                unsigned o[8];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    o[2 * k] = (unsigned)s_two[w[k] & 0xFFu] | ((unsigned)s_two[(w[k] >> 8) & 0xFFu] << 16);
                    o[2 * k + 1] = (unsigned)s_two[(w[k] >> 16) & 0xFFu] | ((unsigned)s_two[w[k] >> 24] << 16);
                }
                S.rd[jj][2 * u] = make_uint4(o[0], o[1], o[2], o[3]);
                S.rd[jj][2 * u + 1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
        cp_async_wait<0>();
    }
    __syncwarp();
    if (!present) return;
    const uint16_t *s_ops = reinterpret_cast<const uint16_t *>(S.ops[j]);
    const uint8_t *s_rd = reinterpret_cast<const uint8_t *>(S.rd[j]), *s_rf = reinterpret_cast<const uint8_t *>(S.rf[j]);
    auto op_word = [&](int ci) -> unsigned {
        const int k = ci - op0;
        return (k >= op_lo && k < op_hi) ? (unsigned)s_ops[k] : (unsigned)cig[ci];
    };
    auto op_len = [&](int ci, unsigned w) -> int {
        const unsigned len = w >> 4;
        if (len != 0xFFFu) return (int)len;
        uint32_t lo = 0, hi = a.b.n_long;
        const uint64_t gop = gop0 + (uint64_t)ci;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (a.b.long_at[mid] < gop) lo = mid + 1; else hi = mid;
        }
        return (lo < a.b.n_long && a.b.long_at[lo] == gop) ? (int)a.b.long_len[lo] : 0xFFF;
    };
    const int ref_len = (int)min(a.ref_len, (long long)INT_MAX - 2);                           // positions are ints: nothing beyond is reachable
    // the rare accesses outside a staged window (a long insertion or deletion inside the +-100 iterations) read global memory
    auto read_char_far = [&](int rp) -> unsigned {
        const unsigned v = (unsigned)seq[rp >> 1];
        return (unsigned)(unsigned char)"=ACMGRSVTWYHKDBN"[(v >> ((~rp & 1) << 2)) & 0xfu];     // bam_seqi
    };
    auto ref_char_far = [&](int fp) -> unsigned {
        if (fp == ref_len) return 0u;                                                          // std::string::operator[](size())
        return (unsigned)(unsigned char)a.ref[fp];
    };
    const unsigned sq_span = (unsigned)(sq_hi - sq_lo), rf_span = (unsigned)(rf_hi - rf_lo);
    int32_t *hist = a.window_hist + (size_t)it.slot2 * LPS_WINDOW_BINS + LPS_WINDOW;
    // getWindowsDiffRef (:688-710): the covering op is an M/=/X op, never an insertion
    unsigned opw = op_word(ci0);
    const int oplen = op_len(ci0, opw);
    int remaining = dir < 0 ? (off > 0 ? off : 0) : (oplen - off > 0 ? oplen - off : 0);
    int ci = ci0, read_pos = qidx, ref_pos = var_pos, i = 1, first = 0;
    unsigned op = opw & 15u;
    bool hopping = false;
    // getOrderWindowsDiffRef (:655-686) + processCigarOperation (:627-654), trip by trip.  `remaining` is the budget BEFORE the
    // decrement of iteration i; a budget of 1 or 0 hops before the iteration runs (which then needs no further decrement: first = 1).
    while (i <= LPS_WINDOW) {
        if (!hopping && (unsigned)remaining <= 1u) { remaining -= 1; hopping = true; }
        if (hopping) {
            ci += dir;
            if (ci >= ncig || ci < 0) break;
            opw = op_word(ci);
            op = opw & 15u;
            const int len = op_len(ci, opw);
            if ((0x1C9u >> op) & 1u) { remaining += len; hopping = false; first = 1; }        // M N P = X
            else if (op == 1u) { read_pos += len * dir; continue; }
            else if (op == 2u) { ref_pos += len * dir; continue; }
            else break;
        }
        // the iterations that follow without a hop: until the decrement gives 0; a negative budget never hops again
        const int extra = remaining >= 2 ? remaining - 1 : (remaining < 0 ? LPS_WINDOW : 0);
        int n = first + extra;
        if (n > LPS_WINDOW + 1 - i) n = LPS_WINDOW + 1 - i;
        if (n > WD_CHUNK) n = WD_CHUNK;
        if (!((0x14Eu >> op) & 1u)) {                                                          // not I D N P X: a moving op
            // iteration i + t compares read[read_pos + dir (t+1)] with ref[ref_pos + dir (t+1)]; the scan ends at the first position
            // out of range (read == l_qseq is one past SEQ, undefined in the reference: ends the scan as well)
            int n_ok;
            if (dir > 0) n_ok = min(lq - 1 - read_pos, ref_len - ref_pos);
            else n_ok = (read_pos > lq || ref_pos > ref_len + 1) ? 0 : min(read_pos, ref_pos);
            if (n_ok < 0) n_ok = 0;
            const int nc = min(n, n_ok);
            // indices of the chunk's first and last base inside the staged windows; one range check for the chunk
            const int kr = read_pos - rd0 - sq_lo, kf = ref_pos - rf0 - rf_lo;
            const bool near = (unsigned)(kr + dir) < sq_span && (unsigned)(kr + dir * nc) < sq_span && (unsigned)(kf + dir) < rf_span &&
                              (unsigned)(kf + dir * nc) < rf_span;
            if (near) {
                // the chunk's bases as words: nc bytes upwards from index +1 (forward) or downwards from index -1 (backward) of both windows
                const unsigned ar = (unsigned)(sq_lo + kr + (dir > 0 ? 1 : -nc)), af = (unsigned)(rf_lo + kf + (dir > 0 ? 1 : -nc));
                const unsigned *wr = reinterpret_cast<const unsigned *>(s_rd) + (ar >> 2), *wf = reinterpret_cast<const unsigned *>(s_rf) + (af >> 2);
                const unsigned sr = (ar & 3u) * 8u, sf = (af & 3u) * 8u;
#pragma unroll
                for (int h = 0; h < WD_CHUNK / 4; h++) {
                    if (4 * h < nc) {
                        const unsigned x = __funnelshift_r(wr[h], wr[h + 1], sr) ^ __funnelshift_r(wf[h], wf[h + 1], sf);
                        const int m = min(nc - 4 * h, 4);                                  // bytes of this word that belong to the chunk
                        unsigned d = m == 4 ? x : (x & ((1u << (8 * m)) - 1u));
                        while (d) {
                            const int byte = (__ffs((int)d) - 1) >> 3;                     // a differing base: byte `byte` of word h
                            d &= ~(0xFFu << (8 * byte));
                            const int pos = 4 * h + byte;                                  // its rank inside the chunk, counted upwards
                            const int k = dir > 0 ? pos : nc - 1 - pos;
                            atomicAdd(hist + (i + k) * dir, 1);
                        }
                    }
                }
            } else {
                for (int k = 0; k < nc; k++) {
                    const int rp = read_pos + dir * (k + 1), fp = ref_pos + dir * (k + 1);
                    const int xr = rp - rd0 - sq_lo, xf = fp - rf0 - rf_lo;
                    const unsigned cr = (unsigned)xr < sq_span ? (unsigned)s_rd[sq_lo + xr] : read_char_far(rp);
                    const unsigned cf = (unsigned)xf < rf_span ? (unsigned)s_rf[rf_lo + xf] : ref_char_far(fp);
                    if (cr != cf) atomicAdd(hist + (i + k) * dir, 1);
                }
            }
            if (nc < n) break;
            read_pos += dir * n; ref_pos += dir * n;
        }
        remaining -= n - first;
        first = 0;
        i += n;
    }
}

}  // namespace

int lps_launch_window_diff(lps_ctx *ctx, int have_reference) {
    const unsigned long long n = ctx->n_wd_items;
    ctx->stats.ms_kernel_window_diff = 0.f;
    if (n == 0) return LPS_OK;
    WdArgs a;
    a.b = ctx->batch; a.items = ctx->d_wd_items.p; a.n_items = n; a.vpos = ctx->var.pos; a.tum_var = ctx->som.tum_var;
    a.ref = ctx->d_ref.p; a.ref_len = have_reference ? (long long)ctx->ref_len : 0; a.window_hist = ctx->som.window_hist;
    const unsigned long long per_cta = (unsigned long long)WD_WARPS * WD_ITEMS;
    cudaEventRecord(ctx->kev[4], ctx->stream);
    k_window_diff<<<(unsigned)((n + per_cta - 1) / per_cta), WD_WARPS * 32, 0, ctx->stream>>>(a);   // 2 lanes per item
    cudaEventRecord(ctx->kev[5], ctx->stream);
    ctx->stats.kernel_launches++;
    LPS_CUDA(ctx, cudaGetLastError());
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.ms_kernel_window_diff, ctx->kev[4], ctx->kev[5]);
    return LPS_OK;
}
