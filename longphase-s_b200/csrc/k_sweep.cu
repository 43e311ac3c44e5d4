// k_sweep.cu — VairiantGraph::edgeConnectResult on the device (reference src/phase/PhasingGraph.cpp:286-474, with
// VariantEdge::findBestEdgePair :166-228 already reduced to one vote byte per edge cell by k_fold_edges and Onelongcase :251-283).
//
// The sweep is a left-to-right chain: node k's haplotype comes from the votes of its <= W predecessors (two float sums whose
// order of addition is part of the result, plus Onelongcase's integer counters), then k votes on its <= W successors.  One
// sequential pass over 65 k nodes costs milliseconds on any single processor.  But the state that crosses a point of the chain
// is small: the haplotypes (up to one global flip - the rule is symmetric in the two haplotypes) and voted / skipped status of
// the last W voters, and `lastConnectPos`.  So the chain is cut into SEGMENTS of S nodes, one warp each, all running at once:
//   1. k_sweep_segments: segment p starts H nodes EARLY (a halo) from an empty state; by the time it reaches its own first
//      node the accumulators only hold votes of nodes it decided itself.  It records its decisions for its core nodes, and for
//      the last W halo nodes separately.
//   2. k_sweep_verify: for every boundary, the W halo decisions of segment p are compared with what segment p-1 decided for
//      the same nodes: same assigned / new-block / voted pattern and haplotypes equal up to ONE flip.  If so the state at the
//      boundary is the true one (by induction from segment 0, which starts at the true start), up to that flip, which is
//      resolved by a scan over the segments (a block start inside a segment re-anchors the orientation: hp = 1).
//   3. k_sweep_fallback: if any boundary disagrees, ONE warp redoes the whole chain sequentially (exact by construction) -
//      the launch is unconditional, the kernel returns at once when step 2 succeeded, so the host never has to look.
//   4. finish: block starts by a max-scan, single-member blocks dropped (:425), PS = position(block start) + 1, per VARIANT.
// Inside a warp node s lives in lane s & 31, slot (s >> 5) & 1 (the window of W <= 63 successors never holds two nodes of one
// slot), so no accumulator ever moves; the vote rows and the per-node meta words are staged through shared memory by bulk
// copies (cp.async.bulk on an mbarrier) one chunk ahead of the chain.
#include <climits>
#include <cub/cub.cuh>
#include "lps_ctx.cuh"
#include "lps_async.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SW_CH = 64;        // nodes per staged chunk
constexpr int SW_RS_MAX = 80;    // lps_vote_row_stride(63)
constexpr int SW_WARPS = 4;      // segments per CTA
#ifndef LPS_SW_SEG
#define LPS_SW_SEG 128
#endif
constexpr int SW_SEG = LPS_SW_SEG;      // core nodes per segment (a warp walks halo + core = ~270 nodes; 64 k nodes are 512 warps, one wave)
constexpr int SW_HALO_W = 4;     // halo = SW_HALO_W * W nodes

struct SweepArgs {
    const int32_t *n_nodes;      // device: number of graph nodes
    int W, RS, seg, halo, n_seg;
    const uint8_t *votes;        // [n_nodes][RS] vote bytes of voter k, byte j = vote on node 16 * ((k + 1) / 16) + j
    const uint16_t *meta;        // [n_nodes] node type | gap << 3 | (last_link + 1) << 8
    uint8_t *flags;              // [n_nodes] 1 assigned | 2 new block | (hp - 1) << 2 | 8 voted
    uint8_t *halo_flags;         // [n_seg][W]
    int32_t *first_nb;           // [n_seg] first block start inside the segment's core, INT_MAX if none
    uint8_t *seg_flip;           // [n_seg] flip of the core nodes before first_nb
    int32_t *all_ok;             // 1 when every boundary verified
};

struct alignas(16) SweepSmem {
    uint8_t rows[2][SW_CH * SW_RS_MAX];
    uint16_t meta[2][SW_CH];
    unsigned long long mbar[2];
    uint8_t flags[SW_SEG + SW_HALO_W * 64];          // decisions of a segment (halo + core), flushed once at its end
};

// What a voter adds to a successor for every value of a vote byte (5 bits), per (voter haplotype, voter type class): the two float
// increments and the packed Onelongcase increment.  Type classes: 0 SNP-like (weights 1 / 20, counted by Onelongcase), 1 indel
// (type 3: same weights, not counted), 2 danger indel (type 4: weight 0.1f).  float4 so that one 128-bit load fetches an entry.
constexpr int SW_LUT = 32 * 2 * 3;

// everything VariantEdge::findBestEdgePair / edgeConnectResult (:360-417) make of one vote byte, for a voter of haplotype `same`
__device__ __forceinline__ void vote_of(unsigned info, unsigned same, float w_lo, float w_hi, bool type_ok, float &d1, float &d2, unsigned &dp) {
    const unsigned link = info & 3u;
    const bool heavy = (info & 4u) != 0;
    const float wb = heavy ? w_hi : w_lo;                                    // 1 / 20 / 0.1f (:367-369)
    const bool tA = link == same, tB = link != 0u && !tA;
    d1 = tA ? wb : 0.f; d2 = tB ? wb : 0.f;                                   // + 0.0f is exact; the sums are never -0.0
    const bool single = (info & 8u) != 0;
    const bool qual = !single && (info & 16u) != 0 && type_ok;              // Onelongcase: weight >= 1, voter not an indel (:265)
    const unsigned wi = heavy ? 20u : 1u;
    dp = ((link != 0u && single) ? 1u : 0u) | ((qual && tA) ? (wi << 8) : 0u) | ((qual && tB) ? (wi << 20) : 0u);
}

__device__ __forceinline__ void build_vote_lut(float4 *lut) {
    for (int i = threadIdx.x; i < SW_LUT; i += blockDim.x) {
        const unsigned info = i & 31u, same = ((i >> 5) & 1u) ? 2u : 1u, cls = (unsigned)i >> 6;
        float d1, d2; unsigned dp;
        vote_of(info, same, cls == 2u ? (float)0.1 : 1.f, cls == 2u ? (float)0.1 : 20.f, cls == 0u, d1, d2, dp);
        lut[i] = make_float4(d1, d2, __uint_as_float(dp), 0.f);
    }
}

// the chain over nodes [k_begin, k_end), from an empty state at k_begin; decisions of nodes >= k_core go to flags[], those of the
// W nodes before k_core to halo_flags[seg]
__device__ __forceinline__ void sweep_range(const SweepArgs &a, SweepSmem &S, const float4 *__restrict__ lut, const int N, const int k_begin,
                                            const int k_core, const int k_end, const int seg, const int lane) {
    float w1a = 0.f, w2a = 0.f, w1b = 0.f, w2b = 0.f;
    unsigned pka = 0u, pkb = 0u;
    int last_connect = -1, first_nb = INT_MAX;
    const int W = a.W, RS = a.RS;
    const int k_stop = min(k_end, N - 1);            // the loop of :313 needs a successor
    const bool staged = seg >= 0 && k_stop - k_begin <= (int)sizeof(S.flags);
    if (k_begin < k_stop) {
        int c = k_begin / SW_CH;
        const int c_last = (k_stop - 1) / SW_CH;
        auto issue = [&](int chunk, int b) {
            if (lane == 0) {
                const int k0 = chunk * SW_CH;
                const int rows = min(SW_CH, N - k0);
                const uint32_t b_rows = (uint32_t)rows * (uint32_t)RS, b_meta = (uint32_t)((rows * 2 + 15) & ~15);
                const uint32_t bar = smem_u32(&S.mbar[b]);
                mbar_expect_tx(bar, b_rows + b_meta);
                bulk_g2s(smem_u32(&S.rows[b][0]), a.votes + (size_t)k0 * RS, b_rows, bar);
                bulk_g2s(smem_u32(&S.meta[b][0]), a.meta + k0, b_meta, bar);
            }
        };
        int buf = 0;
        uint32_t phase = 0u;
        issue(c, 0);
        for (; c <= c_last; c++) {
            if (c + 1 <= c_last) issue(c + 1, buf ^ 1);
            mbar_wait(smem_u32(&S.mbar[buf]), (phase >> buf) & 1u);
            phase ^= 1u << buf;
            const int k0 = c * SW_CH;
            const int ka = max(k0, k_begin), kb = min(k0 + SW_CH, k_stop);
            for (int k = ka; k < kb; k++) {
                const unsigned m = S.meta[buf][k - k0];
                const bool gap = (m & 8u) != 0;                                 // |pos[k+1] - pos[k]| > distance (:318-320)
                const unsigned type = m & 7u;
                const int Lp1 = (int)(m >> 8);
                const int owner = k & 31;
                const bool sb = ((k >> 5) & 1) != 0;
                float h1 = sb ? w1b : w1a, h2 = sb ? w2b : w2a;
                const unsigned pk = sb ? pkb : pka;
                if (lane == owner) { if (sb) { w1b = 0.f; w2b = 0.f; pkb = 0u; } else { w1a = 0.f; w2a = 0.f; pka = 0u; } }   // the slot now belongs to node k + 64
                unsigned f = 0u;
                bool votes = false;
                int hp = 1;
                if (!gap) {
                    const int sg = (int)(pk & 0xFFu), a1 = (int)((pk >> 8) & 0xFFFu), a2 = (int)(pk >> 20);
                    if (sg > 3 && (a1 | a2) != 0) { h1 = (float)a1; h2 = (float)a2; }                                      // Onelongcase :276-281
                    unsigned bits = (h1 == h2 ? 1u : 0u) | (h1 > h2 ? 2u : 0u);
                    bits = __shfl_sync(FULL, bits, owner);
                    bool assigned = true, nb = false;
                    if (bits & 1u) {
                        if (last_connect >= 0 && k < last_connect) assigned = false;                                       // :340-342
                        else nb = true;                                                                                     // new block, hp = 1
                    } else hp = (bits & 2u) ? 1 : 2;
                    if (assigned) {
                        votes = Lp1 != 0;
                        f = 1u | (nb ? 2u : 0u) | ((unsigned)(hp - 1) << 2) | (votes ? 8u : 0u);
                        if (nb && k >= k_core && first_nb == INT_MAX) first_nb = k;
                    }
                }
                if (lane == 0) {
                    if (staged) S.flags[k - k_begin] = (uint8_t)f;
                    else if (k >= k_core) a.flags[k] = (uint8_t)f;
                }
                if (votes) {
                    const int base = (k + 1) & ~15;
                    const uint8_t *row = S.rows[buf] + (k - k0) * RS;
                    const float4 *__restrict__ L = lut + (((type == 4u ? 2u : type == 3u ? 1u : 0u) << 6) | ((unsigned)(hp - 1) << 5));
                    const int j0 = (lane - base) & 31;
                    const bool slot = (((base + j0) >> 5) & 1) != 0;           // slot of the node byte j0 votes on; byte j0 + 32 votes on the other slot
                    float4 c0 = L[row[j0]], c1 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j0 + 32 < RS) c1 = L[row[j0 + 32]];
                    if (RS > 64 && j0 + 64 < RS) { const float4 c2 = L[row[j0 + 64]]; c0.x += c2.x; c0.y += c2.y; c0.z = __uint_as_float(__float_as_uint(c0.z) + __float_as_uint(c2.z)); }
                    const float4 ca = slot ? c1 : c0, cb = slot ? c0 : c1;
                    w1a += ca.x; w2a += ca.y; pka += __float_as_uint(ca.z);
                    w1b += cb.x; w2b += cb.y; pkb += __float_as_uint(cb.z);
                    last_connect = k + Lp1;
                }
            }
            __syncwarp();
            buf ^= 1;
        }
    }
    if (staged && k_begin < k_stop) {
        __syncwarp();
        for (int k = k_begin + lane; k < k_stop; k += 32) {
            const uint8_t f = S.flags[k - k_begin];
            if (k >= k_core) a.flags[k] = f;
            else if (k >= k_core - W) a.halo_flags[(size_t)seg * W + (k - (k_core - W))] = f;
        }
    }
    if (seg >= 0 && lane == 0) a.first_nb[seg] = first_nb;
}

__global__ void __launch_bounds__(SW_WARPS * 32) k_sweep_segments(SweepArgs a) {
    __shared__ SweepSmem s_all[SW_WARPS];
    __shared__ float4 s_lut[SW_LUT];
    build_vote_lut(s_lut);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    SweepSmem &S = s_all[wib];
    if (lane == 0) { mbar_init(smem_u32(&S.mbar[0]), 1u); mbar_init(smem_u32(&S.mbar[1]), 1u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int N = *a.n_nodes;
    const int p = blockIdx.x * SW_WARPS + wib;
    if (p >= a.n_seg) return;
    const int b = p * a.seg, e = min(N, b + a.seg);
    // halo nodes without a decision (before node 0) read as "nothing there" in the comparison
    if (lane == 0) for (int i = 0; i < a.W; i++) a.halo_flags[(size_t)p * a.W + i] = 0xFFu;
    __syncwarp();
    if (b >= N) { if (lane == 0) a.first_nb[p] = INT_MAX; return; }
    sweep_range(a, S, s_lut, N, max(0, b - a.halo), b, e, p, lane);
}

// boundary checks, one warp per boundary over as many CTAs as it takes (a single CTA took 28 us for 506 boundaries: 16 dependent rounds
// per warp); the flips relative to the truth are a scan over the segments, done at the start of k_sweep_fallback
__global__ void __launch_bounds__(256) k_sweep_verify(SweepArgs a) {
    // one warp per boundary, lane i compares the i-th node of the window (W <= 63: two rounds at most).  The verdict of boundary p
    // is left in seg_flip[p] as bits: 1 = flipped relative to segment p - 1, 2 = a block start inside the core re-anchors the
    // orientation, 4 = the window does not agree up to ONE flip.  k_sweep_fallback turns them into the absolute flips.
    const int N = *a.n_nodes;
    const int lane = threadIdx.x & 31, p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= a.n_seg) return;
    const int b = p * a.seg, e = min(N, b + a.seg);
    bool bad = false;
    unsigned flip_yes = 0u, flip_no = 0u;               // assigned nodes whose haplotype differs / agrees
    if (p > 0 && b < N - 1 && b - a.halo > 0) {          // a halo that starts at node 0 starts from the true state: nothing to check
        for (int i = lane; i < a.W; i += 32) {
            const int kk = b - a.W + i;
            if (kk < 0) continue;
            const unsigned t = a.flags[kk], h = a.halo_flags[(size_t)p * a.W + i];
            if (h == 0xFFu || ((t ^ h) & 0xBu)) bad = true;
            else if (t & 1u) { if (((t ^ h) >> 2) & 1u) flip_yes = 1u; else flip_no = 1u; }
        }
    }
    const bool any_bad = __any_sync(0xffffffffu, bad), any_yes = __any_sync(0xffffffffu, flip_yes != 0u),
               any_no = __any_sync(0xffffffffu, flip_no != 0u);
    if (lane == 0)
        a.seg_flip[p] = (uint8_t)((any_yes ? 1u : 0u) | ((b < N && a.first_nb[p] < e) ? 2u : 0u) | ((any_bad || (any_yes && any_no)) ? 4u : 0u));
}

// the exact sequential chain, only when a boundary did not verify
__global__ void __launch_bounds__(32) k_sweep_fallback(SweepArgs a, int forced) {
    __shared__ SweepSmem S;
    __shared__ float4 s_lut[SW_LUT];
    __shared__ uint8_t s_pk[4096];
    if (!forced) {
        // the scan over the boundaries' verdicts: flip in force inside segment p = its relative flip XOR the flip at the end of p - 1
        // (a block start inside a core re-anchors the orientation)
        // carry(p) = XOR of the relative flips of the segments after the last re-anchoring one, up to p: 32 segments per step with ballots
        const int ln = threadIdx.x;
        const unsigned below = (1u << ln) - 1u;            // lanes 0 .. ln-1
        unsigned carry = 0u;
        int ok = 1;
        for (int t0 = 0; t0 < a.n_seg; t0 += 4096) {        // a tile of verdicts at a time through shared memory: independent loads
            const int tn = min(4096, a.n_seg - t0);
            for (int i = ln; i < tn; i += 32) s_pk[i] = a.seg_flip[t0 + i];
            __syncwarp();
            for (int p0 = t0; p0 < t0 + tn; p0 += 32) {
                const int p = p0 + ln;
                const unsigned pk = p < t0 + tn ? (unsigned)s_pk[p - t0] : 0u;
                const unsigned R = __ballot_sync(0xffffffffu, (pk & 1u) != 0u), A = __ballot_sync(0xffffffffu, (pk & 2u) != 0u);
                if (__any_sync(0xffffffffu, (pk & 4u) != 0u)) ok = 0;
                const unsigned Ap = A & below;
                unsigned cprev;                                 // carry at the end of segment p - 1
                if (Ap) { const int la = 31 - __clz((int)Ap); cprev = (unsigned)__popc(R & below & ~((2u << la) - 1u)) & 1u; }
                else cprev = ((unsigned)__popc(R & below) & 1u) ^ carry;
                if (p < t0 + tn) s_pk[p - t0] = (uint8_t)(p == 0 ? 0u : ((pk & 1u) ^ cprev));
                if (A) { const int la = 31 - __clz((int)A); carry = (unsigned)__popc(R & ~((2u << la) - 1u)) & 1u; }
                else carry = ((unsigned)__popc(R) & 1u) ^ carry;
            }
            __syncwarp();
            for (int i = ln; i < tn; i += 32) a.seg_flip[t0 + i] = s_pk[i];
            __syncwarp();
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ln == 0) *a.all_ok = ok;
        if (ok) return;
    }
    build_vote_lut(s_lut);
    __syncwarp();
    const int lane = threadIdx.x;
    if (lane == 0) { mbar_init(smem_u32(&S.mbar[0]), 1u); mbar_init(smem_u32(&S.mbar[1]), 1u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int N = *a.n_nodes;
    sweep_range(a, S, s_lut, N, 0, 0, N, -1, lane);
    for (int p = lane; p < a.n_seg; p += 32) a.seg_flip[p] = 0;
}

__global__ void k_sweep_block_start(int n_upper, const int32_t *__restrict__ n_nodes, const uint8_t *__restrict__ flags, int32_t *__restrict__ start) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_upper) return;
    start[k] = (k < *n_nodes && (flags[k] & 3u) == 3u) ? k : -1;
}

// a block is kept when it has a second member (:425): any assigned node that is not a block start marks its block
__global__ void k_sweep_mark_multi(int n_upper, const int32_t *__restrict__ n_nodes, const uint8_t *__restrict__ flags, const int32_t *__restrict__ start,
                                   uint8_t *__restrict__ multi) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_upper || k >= *n_nodes) return;
    if ((flags[k] & 3u) == 1u && start[k] >= 0) multi[start[k]] = 1;
}

// per VARIANT: phase set (block start position + 1) and REF-allele haplotype after the sweep
__global__ void k_sweep_finish(int nv, const int32_t *__restrict__ node_of_var, const uint8_t *__restrict__ flags, const int32_t *__restrict__ start,
                               const uint8_t *__restrict__ multi, const int32_t *__restrict__ node_pos, const int32_t *__restrict__ first_nb,
                               const uint8_t *__restrict__ seg_flip, int seg, int32_t *__restrict__ ps, int8_t *__restrict__ hap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const int k = node_of_var[i];
    int out_ps = 0, out_hap = -1;
    if (k >= 0) {
        const unsigned f = flags[k];
        const int st = (f & 1u) ? start[k] : -1;
        if (st >= 0 && !(st == k && !multi[k])) {
            const int p = k / seg;
            const unsigned flip = k < first_nb[p] ? (unsigned)seg_flip[p] : 0u;
            out_ps = node_pos[st] + 1;
            out_hap = (int)(((f >> 2) & 1u) ^ flip);
        }
    }
    ps[i] = out_ps;
    hap[i] = (int8_t)out_hap;
}

struct MaxOp { __device__ __forceinline__ int32_t operator()(int32_t x, int32_t y) const { return x > y ? x : y; } };

}  // namespace

// d_vote_info / d_sweep_meta / d_node_pos / d_node_of_var hold the graph of this contig; n_upper >= number of nodes (the host may
// not know the exact count).  Writes the sweep result per variant into d_ps / d_hap_ref.  No host synchronisation.
int lps_launch_sweep(lps_ctx *ctx, const lps_phase_params *p, int n_upper) {
    cudaStream_t st = ctx->stream;
    const int nv = ctx->var.n, W = ctx->window;
    (void)p;
    LPS_CUDA(ctx, ctx->d_ps.reserve((size_t)nv + 1));
    LPS_CUDA(ctx, ctx->d_hap_ref.reserve((size_t)nv + 1));
    if (n_upper <= 0 || nv <= 0) {
        if (nv > 0) {
            LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_ps.p, 0, 4 * (size_t)nv, st));
            LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_hap_ref.p, 0xFF, (size_t)nv, st));
        }
        return LPS_OK;
    }
    SweepArgs a;
    a.n_nodes = ctx->d_n_nodes.p;
    a.W = W; a.RS = lps_vote_row_stride(W); a.seg = SW_SEG; a.halo = SW_HALO_W * W;
    if (a.seg < 2 * W) a.seg = 2 * W;
    a.n_seg = (n_upper + a.seg - 1) / a.seg;
    LPS_CUDA(ctx, ctx->d_sweep_flags.reserve((size_t)n_upper + 16));
    LPS_CUDA(ctx, ctx->d_sweep_halo.reserve((size_t)a.n_seg * W + 16));
    LPS_CUDA(ctx, ctx->d_sweep_first_nb.reserve((size_t)a.n_seg + 1));
    LPS_CUDA(ctx, ctx->d_sweep_flip.reserve((size_t)a.n_seg + 1));
    LPS_CUDA(ctx, ctx->d_sweep_ok.reserve(1));
    LPS_CUDA(ctx, ctx->d_sweep_start.reserve((size_t)n_upper + 1));
    LPS_CUDA(ctx, ctx->d_sweep_multi.reserve((size_t)n_upper + 1));
    a.votes = ctx->d_vote_info.p; a.meta = ctx->d_sweep_meta.p; a.flags = ctx->d_sweep_flags.p; a.halo_flags = ctx->d_sweep_halo.p;
    a.first_nb = ctx->d_sweep_first_nb.p; a.seg_flip = ctx->d_sweep_flip.p; a.all_ok = ctx->d_sweep_ok.p;
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_sweep_flags.p, 0, (size_t)n_upper + 16, st));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_sweep_multi.p, 0, (size_t)n_upper + 1, st));
    const char *force = getenv("LPS_SWEEP_SEQUENTIAL");      // A/B and tests: skip the speculation, run the exact chain on one warp
    const int forced = (force && force[0] == '1') ? 1 : 0;
    if (forced) {
        LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_sweep_ok.p, 0, 4, st));
        LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_sweep_first_nb.p, 0, 4 * (size_t)a.n_seg, st));
    } else {
        k_sweep_segments<<<(a.n_seg + SW_WARPS - 1) / SW_WARPS, SW_WARPS * 32, 0, st>>>(a);
        k_sweep_verify<<<(a.n_seg + 7) / 8, 256, 0, st>>>(a);
        ctx->stats.kernel_launches += 2;
    }
    k_sweep_fallback<<<1, 32, 0, st>>>(a, forced);
    const int tb = 256, gb = (n_upper + tb - 1) / tb;
    k_sweep_block_start<<<gb, tb, 0, st>>>(n_upper, ctx->d_n_nodes.p, ctx->d_sweep_flags.p, ctx->d_sweep_start.p);
    size_t tmp = 0;
    cub::DeviceScan::InclusiveScan(nullptr, tmp, ctx->d_sweep_start.p, ctx->d_sweep_start.p, MaxOp(), n_upper, st);
    LPS_CUDA(ctx, ctx->d_cub_tmp.reserve(tmp + 256));
    cub::DeviceScan::InclusiveScan(ctx->d_cub_tmp.p, tmp, ctx->d_sweep_start.p, ctx->d_sweep_start.p, MaxOp(), n_upper, st);
    k_sweep_mark_multi<<<gb, tb, 0, st>>>(n_upper, ctx->d_n_nodes.p, ctx->d_sweep_flags.p, ctx->d_sweep_start.p, ctx->d_sweep_multi.p);
    k_sweep_finish<<<(nv + tb - 1) / tb, tb, 0, st>>>(nv, ctx->d_node_of_var.p, ctx->d_sweep_flags.p, ctx->d_sweep_start.p, ctx->d_sweep_multi.p,
                                                      ctx->d_node_pos.p, ctx->d_sweep_first_nb.p, ctx->d_sweep_flip.p, a.seg, ctx->d_ps.p, ctx->d_hap_ref.p);
    ctx->stats.kernel_launches += 5;
    LPS_CUDA(ctx, cudaGetLastError());
    return LPS_OK;
}
