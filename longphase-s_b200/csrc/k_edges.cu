// k_edges.cu — KERNEL 2: graph construction of VairiantGraph::addEdge (reference
// src/phase/PhasingGraph.cpp:795-888) and the ordered float fold of SubEdge::addSubEdge (:25-70).
//
// The reference folds `float` edge weights read by read in std::map<std::string,...> (lexicographic
// read NAME) order: `x++` for a high-quality pair, `x = x + 0.1(double)` otherwise.  The result
// depends on the order, so an unordered atomic add cannot be bit-exact.  Instead of sorting the
// ~35x larger stream of pair contributions, the device
//   1. lays the surviving calls out as "merged reads" M in name-rank order (alignments of one name
//      concatenated, sorted by position) — 4 bytes per call: node << 2 | allele << 1 | q_hi;
//   2. builds, with ONE stable radix sort of the calls by node, the list of calls at each node in
//      name-rank order;
//   3. runs one warp per node `a`: it walks that list sequentially (the fold order) and, for each
//      call, the lanes handle the next <= connectAdjacent calls of the same merged read in parallel,
//      updating a [window][4] float tile kept in shared memory.  Distinct lanes hit distinct cells
//      (positions inside a merged read are distinct, duplicates are serialised), so the per-cell
//      order is exactly the reference's.  The tile is written once, coalesced.
// Cell (a, b) is indexed by node distance d = b - a - 1 < window, nodes being the variants with at
// least one surviving call (the keys of totalVariantInfo): that is the only part of the table the
// sweep ever reads (:360-417); contributions to farther pairs are only counted.
#include <algorithm>
#include <cub/cub.cuh>
#include "lps_ctx.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// per read: alive call count, node marking (last writer decides the node type, :803-832)
__global__ void k_mark_nodes(int n_reads, const uint64_t *__restrict__ call_off, const lps_call *__restrict__ calls,
                             const uint8_t *__restrict__ read_dead, const uint8_t *__restrict__ call_erased,
                             const int32_t *__restrict__ name_rank, unsigned long long *__restrict__ var_lastw,
                             uint32_t *__restrict__ alive_cnt, uint64_t *__restrict__ aln_keys) {
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (wid >= n_reads) return;
    int r = (int)wid;
    uint64_t c0 = call_off[r], c1 = call_off[r + 1];
    int alive = 0;
    if (!read_dead[r]) {
        for (uint64_t c = c0 + lane; c < c1; c += 32) {
            if (call_erased && call_erased[c]) continue;
            lps_call cl = calls[c];
            unsigned type = cl.quality == -4 ? 3u : (cl.quality == -5 ? 4u : 0u);
            atomicMax(&var_lastw[cl.var], ((unsigned long long)(r + 1) << 3) | type);
            alive++;
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) alive += __shfl_xor_sync(FULL, alive, d);
    if (lane == 0) {
        alive_cnt[r] = (uint32_t)alive;
        aln_keys[r] = alive ? (((uint64_t)(uint32_t)name_rank[r] << 32) | (uint32_t)r) : ~0ull;
    }
}

__global__ void k_node_flags(int nv, const unsigned long long *__restrict__ var_lastw, int32_t *__restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nv) flag[i] = var_lastw[i] != 0;
}

__global__ void k_fill_nodes(int nv, const unsigned long long *__restrict__ var_lastw, const int32_t *__restrict__ vpos,
                             int32_t *__restrict__ node_of_var, int32_t *__restrict__ node_var, int32_t *__restrict__ node_pos,
                             uint8_t *__restrict__ node_type) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    if (var_lastw[i] != 0) {
        int k = node_of_var[i];
        node_var[k] = i;
        node_pos[k] = vpos[i];
        node_type[k] = (uint8_t)(var_lastw[i] & 7ull);
    } else node_of_var[i] = -1;
}

__global__ void k_sorted_counts(int n, const uint64_t *__restrict__ keys_sorted, const uint32_t *__restrict__ alive_cnt,
                                uint64_t *__restrict__ cnt_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint64_t k = i < n ? keys_sorted[i] : ~0ull;
    cnt_out[i] = (k == ~0ull) ? 0 : alive_cnt[(uint32_t)k];
}

// one warp per sorted alignment: append its alive calls to M
__global__ void k_fill_merged(int n_aln, const uint64_t *__restrict__ keys_sorted, const uint64_t *__restrict__ grp_off,
                              const uint64_t *__restrict__ call_off, const lps_call *__restrict__ calls,
                              const uint8_t *__restrict__ call_erased, const int32_t *__restrict__ node_of_var, int base_quality,
                              uint32_t *__restrict__ M, uint32_t *__restrict__ M_gend, uint32_t *__restrict__ node_cnt) {
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (wid >= n_aln) return;
    int i = (int)wid;
    uint64_t key = keys_sorted[i];
    uint32_t rank = (uint32_t)(key >> 32);
    int r = (int)(uint32_t)key;
    // end of the merged group: first later alignment with another name
    int j = i + 1;
    while (j < n_aln && (uint32_t)(keys_sorted[j] >> 32) == rank) j++;
    uint32_t gend = (uint32_t)grp_off[j];
    uint64_t c0 = call_off[r], c1 = call_off[r + 1];
    uint64_t w = grp_off[i];
    for (uint64_t cb = c0; cb < c1; cb += 32) {
        uint64_t c = cb + lane;
        bool ok = c < c1 && !(call_erased && call_erased[c]);
        unsigned m = __ballot_sync(FULL, ok);
        if (ok) {
            lps_call cl = calls[c];
            int q = cl.quality < 0 ? 60 : cl.quality;                       // -4/-5 -> 60 (:820-828)
            int node = node_of_var[cl.var];
            uint64_t dst = w + __popc(m & ((1u << lane) - 1u));
            M[dst] = ((uint32_t)node << 2) | ((uint32_t)cl.allele << 1) | (q >= base_quality ? 1u : 0u);
            M_gend[dst] = gend;
            atomicAdd(&node_cnt[node], 1u);
        }
        w += __popc(m);
    }
}

// merged reads made of several alignments: order by position (== node index).  The runs are already
// sorted, so an in-place insertion sort moves little.  Equal positions keep their concatenation order,
// which is what std::sort's insertion-sort branch (n <= 16) does; larger groups with ties are flagged
// for the host, which replays the same std::sort as the reference (ReadVariant::sort, Util.cpp:3-5).
__global__ void k_sort_multi_groups(int n_aln, const uint64_t *__restrict__ keys_sorted, const uint64_t *__restrict__ grp_off,
                                    uint32_t *__restrict__ M, uint32_t *__restrict__ M_unsorted, uint2 *__restrict__ tie_groups,
                                    uint32_t tie_cap, unsigned int *__restrict__ n_tie) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_aln) return;
    uint32_t rank = (uint32_t)(keys_sorted[i] >> 32);
    if (i > 0 && (uint32_t)(keys_sorted[i - 1] >> 32) == rank) return;          // not a group head
    int j = i + 1;
    while (j < n_aln && (uint32_t)(keys_sorted[j] >> 32) == rank) j++;
    if (j == i + 1) return;                                                      // single alignment
    uint64_t g0 = grp_off[i], g1 = grp_off[j];
    for (uint64_t a = g0; a < g1; a++) M_unsorted[a] = M[a];                     // concatenation order, for the host replay
    bool tie = false;
    for (uint64_t a = g0 + 1; a < g1; a++) {
        uint32_t v = M[a];
        uint64_t b = a;
        while (b > g0 && (M[b - 1] >> 2) > (v >> 2)) { M[b] = M[b - 1]; b--; }
        if (b > g0 && (M[b - 1] >> 2) == (v >> 2)) tie = true;
        M[b] = v;
    }
    if (tie && g1 - g0 > 16) {
        unsigned k = atomicAdd(n_tie, 1u);
        if (k < tie_cap) tie_groups[k] = make_uint2((uint32_t)g0, (uint32_t)g1);
    }
}

// staging of the tied groups for the host std::sort replay: one warp per group
__global__ void k_tie_copy(int n_groups, const uint2 *__restrict__ groups, const uint32_t *__restrict__ stage_off,
                           uint32_t *__restrict__ M, uint32_t *__restrict__ stage, int to_stage) {
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (wid >= n_groups) return;
    uint2 g = groups[wid];
    uint32_t o = stage_off[wid];
    for (uint32_t a = g.x + lane; a < g.y; a += 32) {
        if (to_stage) stage[o + (a - g.x)] = M[a]; else M[a] = stage[o + (a - g.x)];
    }
}

__global__ void k_split_merged(uint64_t n, const uint32_t *__restrict__ M, uint32_t *__restrict__ node, uint32_t *__restrict__ idx) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { node[i] = M[i] >> 2; idx[i] = (uint32_t)i; }
}

__global__ void k_widen_u32b(int n, const uint32_t *__restrict__ in, uint64_t *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = i < n ? in[i] : 0;
}

// the fold: one warp per node
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_fold_edges(int n_nodes, int W, double edge_weight,
                                                          const uint64_t *__restrict__ node_off,
                                                          const uint32_t *__restrict__ list, const uint32_t *__restrict__ M,
                                                          const uint32_t *__restrict__ M_gend, float *__restrict__ weights,
                                                          double edge_threshold, uint8_t *__restrict__ vote_info,
                                                          unsigned long long *__restrict__ counters, uint64_t n_merged,
                                                          int8_t *__restrict__ last_link, int RS) {
    extern __shared__ float s_acc[];                       // [WARPS][W*4]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wid = (long long)blockIdx.x * WARPS + wib;
    if (wid >= n_nodes) return;
    const int a = (int)wid;
    float *acc = s_acc + (size_t)wib * W * 4;
    for (int i = lane; i < W * 4; i += 32) acc[i] = 0.0f;
    __syncwarp();
    const uint64_t l0 = node_off[a], l1 = node_off[a + 1];
    unsigned long long contrib = 0, far = 0;
    if (W <= 64) {
        // Software pipeline: the words of call t+1 (its own entry, the end of its merged read, the <= W entries that follow it in
        // the read) are requested before call t is folded, so the fold never waits on a dependent load chain.  Lane l owns the
        // successors l and 32 + l of a call (W <= 64): one round per call instead of two.
        const int W2 = W > 32 ? W - 32 : 0;
        const uint32_t nm = (uint32_t)n_merged;            // merged entries are indexed with 32 bits (d_M_idx is uint32)
        auto fetch = [&](uint32_t m, uint32_t &ea, uint32_t &gend, uint32_t &e0, uint32_t &e1) {
            ea = M[m]; gend = M_gend[m];
            const uint32_t i0 = m + 1u + (uint32_t)lane, i1 = i0 + 32u;
            e0 = (lane < W && i0 < nm) ? M[i0] : 0xffffffffu;
            e1 = (lane < W2 && i1 < nm) ? M[i1] : 0xffffffffu;
        };
        uint32_t m_cur = 0, m_nxt = 0, ea = 0, gend = 0, e0 = 0xffffffffu, e1 = 0xffffffffu;
        if (l0 < l1) { m_cur = list[l0]; fetch(m_cur, ea, gend, e0, e1); }
        if (l0 + 1 < l1) m_nxt = list[l0 + 1];
        unsigned c32 = 0, f32 = 0;                         // per-lane counts of this node (<= 2 per call), widened once at the end
        const float wlo = (float)edge_weight;
        (void)wlo;
        for (uint64_t t = l0; t < l1; t++) {
            const uint32_t m = m_cur, ea_c = ea, gend_c = gend;
            uint32_t eb0 = e0, eb1 = e1;
            if (t + 1 < l1) {
                m_cur = m_nxt;
                fetch(m_cur, ea, gend, e0, e1);
                if (t + 2 < l1) m_nxt = list[t + 2];
            }
            const unsigned al_a = (ea_c >> 1) & 1u, hi_a = ea_c & 1u;
            const bool act0 = lane < W && m + 1u + (uint32_t)lane < gend_c;
            const bool act1 = lane < W2 && m + 33u + (uint32_t)lane < gend_c;
            // the read reaches beyond 32 successors (warp-uniform).  Rare for reads that carry ~20 calls, so everything that concerns
            // the second slot of a lane sits behind this one branch instead of riding along predicated in every iteration.
            const bool second = __any_sync(FULL, act1);
            if (!act0) eb0 = 0xffffffffu;
            const int nb0 = (int)(eb0 >> 2);
            // duplicates (two calls of one merged read at the same position) are adjacent in the read's sorted list
            const uint32_t p0 = __shfl_up_sync(FULL, eb0 >> 2, 1);
            bool dup = act0 && lane > 0 && p0 == (uint32_t)nb0;
            const int d0 = nb0 - a - 1;
            const bool dense0 = act0 && d0 >= 0 && d0 < W;
            c32 += (unsigned)dense0;
            f32 += (unsigned)(act0 && !dense0);
            float *cell0 = dense0 ? &acc[d0 * 4 + (int)(al_a * 2u + ((eb0 >> 1) & 1u))] : nullptr;
            const bool high0 = hi_a && (eb0 & 1u);
            if (!second) {
                if (!__any_sync(FULL, dup)) {
                    // distinct successors of one read hit distinct cells
                    if (dense0) { float x = *cell0; x = high0 ? x + 1.0f : (float)((double)x + edge_weight); *cell0 = x; }   // SubEdge::addSubEdge :40-43, :62-65
                } else {
                    // keep the read's own order
                    for (int l = 0; l < 32; l++) {
                        if (lane == l && dense0) { float x = *cell0; x = high0 ? x + 1.0f : (float)((double)x + edge_weight); *cell0 = x; }
                        __syncwarp();
                    }
                }
            } else {
                if (!act1) eb1 = 0xffffffffu;
                const int nb1 = (int)(eb1 >> 2);
                const uint32_t p1 = __shfl_up_sync(FULL, eb1 >> 2, 1), last0 = __shfl_sync(FULL, eb0 >> 2, 31);
                dup = dup || (act1 && (lane > 0 ? p1 : last0) == (uint32_t)nb1);
                const bool any_dup = __any_sync(FULL, dup);
                const int d1 = nb1 - a - 1;
                const bool dense1 = act1 && d1 >= 0 && d1 < W;
                c32 += (unsigned)dense1;
                f32 += (unsigned)(act1 && !dense1);
                float *cell1 = dense1 ? &acc[d1 * 4 + (int)(al_a * 2u + ((eb1 >> 1) & 1u))] : nullptr;
                const bool high1 = hi_a && (eb1 & 1u);
                if (!any_dup) {
                    if (dense0) { float x = *cell0; x = high0 ? x + 1.0f : (float)((double)x + edge_weight); *cell0 = x; }
                    if (dense1) { float x = *cell1; x = high1 ? x + 1.0f : (float)((double)x + edge_weight); *cell1 = x; }
                } else {
                    // keep the read's own order: successors 0..31, then 32..W-1
                    for (int l = 0; l < 32; l++) {
                        if (lane == l && dense0) { float x = *cell0; x = high0 ? x + 1.0f : (float)((double)x + edge_weight); *cell0 = x; }
                        __syncwarp();
                    }
                    for (int l = 0; l < W2; l++) {
                        if (lane == l && dense1) { float x = *cell1; x = high1 ? x + 1.0f : (float)((double)x + edge_weight); *cell1 = x; }
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
        }
        contrib += c32; far += f32;
    } else {
        // connect_adjacent > 64: plain rounds of 32 successors
        uint32_t m_next = l0 < l1 ? list[l0] : 0;
        for (uint64_t t = l0; t < l1; t++) {
            const uint32_t m = m_next;
            if (t + 1 < l1) m_next = list[t + 1];
            const uint32_t ea = M[m];
            const uint32_t gend = M_gend[m];
            const unsigned al_a = (ea >> 1) & 1u, hi_a = ea & 1u;
            for (int j0 = 0; j0 < W; j0 += 32) {
                const int j = j0 + lane;
                const uint64_t idx = (uint64_t)m + 1 + (uint64_t)j;
                const bool active = j < W && idx < gend;
                const uint32_t eb = active ? M[idx] : 0xffffffffu;
                const int nb = (int)(eb >> 2);
                const uint32_t nb_prev = __shfl_up_sync(FULL, eb >> 2, 1);
                const bool dup = active && lane > 0 && nb_prev == (uint32_t)nb;
                const bool any_dup = __any_sync(FULL, dup);
                const int d = nb - a - 1;
                const bool dense = active && d >= 0 && d < W;
                if (active) { if (dense) contrib++; else far++; }
                float *cell = dense ? &acc[d * 4 + (int)(al_a * 2u + ((eb >> 1) & 1u))] : nullptr;
                const bool high = hi_a && (eb & 1u);
                if (!any_dup) {
                    if (dense) {
                        float x = *cell;
                        x = high ? x + 1.0f : (float)((double)x + edge_weight);      // SubEdge::addSubEdge :40-43, :62-65
                        *cell = x;
                    }
                } else {
                    // two calls of one merged read at the same position: keep the read's own order
                    for (int l = 0; l < 32; l++) {
                        if (lane == l && dense) {
                            float x = *cell;
                            x = high ? x + 1.0f : (float)((double)x + edge_weight);
                            *cell = x;
                        }
                        __syncwarp();
                    }
                }
                __syncwarp();
                if (!__any_sync(FULL, active)) break;
            }
        }
    }
    __syncwarp();
    float *out = weights + (size_t)a * W * 4;
    for (int i = lane; i < W * 4; i += 32) out[i] = acc[i];
    // epilogue: everything VariantEdge::findBestEdgePair (:166-228) derives from a cell, one byte per successor:
    //   bits 0-1 link (1 same haplotype, 2 opposite, 0 none), bit 2 weight-20 rule, bit 3 (para+cross) <= 1,
    //   bit 4 edgeSimilarRatio < 0.2.  The host sweep (host_phase.cpp) only reads these bytes.
    // Row layout (see lps_host_sweep): RS = lps_vote_row_stride(W) bytes per voter, byte j = vote on node 16*((a+1)/16) + j, so that
    // the host adds whole aligned 16-node blocks; bytes without a vote are written as 0 (the buffer is never memset).
    int last = -1;
    const int shift = (a + 1) & 15;
    for (int j = lane; j < RS; j += 32) {
        const int d = j - shift;
        unsigned info = 0;
        if (d >= 0 && d < W && a + 1 + d < n_nodes) {
            const float rr = acc[d * 4 + 0], ra = acc[d * 4 + 1], ar = acc[d * 4 + 2], aa = acc[d * 4 + 3];
            const float para = rr + aa, cross = ar + ra;
            const double esr = (double)fminf(para, cross) / (double)fmaxf(para, cross);
            unsigned link = 0;
            if (rr + aa > ra + ar) link = 1; else if (rr + aa < ra + ar) link = 2;
            if (esr > edge_threshold) link = 0;
            info = link;
            if ((esr <= 0.1 && (rr + aa + ra + ar) >= 1) || ((rr + aa) < 1 && (ra + ar) >= 1) || ((rr + aa) >= 1 && (ra + ar) < 1)) info |= 4u;
            if ((para + cross) <= 1) info |= 8u;
            if (esr < 0.2) info |= 16u;
            if (info & 3u) last = max(last, d);
        }
        vote_info[(size_t)a * RS + j] = (uint8_t)info;
    }
#pragma unroll
    for (int dd = 16; dd; dd >>= 1) last = max(last, __shfl_xor_sync(FULL, last, dd));
    if (lane == 0) last_link[a] = (int8_t)last;
#pragma unroll
    for (int dd = 16; dd; dd >>= 1) {
        contrib += __shfl_xor_sync(FULL, contrib, dd);
        far += __shfl_xor_sync(FULL, far, dd);
    }
    if (lane == 0) {
        if (contrib) atomicAdd(&counters[0], contrib);
        if (far) atomicAdd(&counters[1], far);
    }
}

int bits_for(uint32_t n) { int b = 1; while (b < 32 && (1ull << b) < (uint64_t)n + 1) b++; return b; }

}  // namespace

// host fix-up of merged reads with tied positions and more than 16 calls (see k_sort_multi_groups)
int lps_host_fix_tie_groups(lps_ctx *ctx, int n_groups);

// Merged reads with tied positions and more than 16 calls: libstdc++'s std::sort is not stable there, so the
// reference's order of the tied calls is whatever its introsort leaves (ReadVariant::sort, Util.cpp:3-5).  The groups
// (in concatenation order, saved by k_sort_multi_groups) are staged to the host in one copy, sorted with the very same
// std::sort and comparator (position order == node order), and written back.
namespace {
struct SortRec { int position; uint32_t packed; };
struct ByPosition { bool operator()(const SortRec &a, const SortRec &b) const { return a.position < b.position; } };
}
int lps_host_fix_tie_groups(lps_ctx *ctx, int n_groups) {
    cudaStream_t st = ctx->stream;
    std::vector<uint2> groups((size_t)n_groups);
    LPS_CUDA(ctx, cudaMemcpy(groups.data(), ctx->d_tie_groups.p, sizeof(uint2) * (size_t)n_groups, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> off((size_t)n_groups + 1, 0);
    for (int k = 0; k < n_groups; k++) off[(size_t)k + 1] = off[(size_t)k] + (groups[(size_t)k].y - groups[(size_t)k].x);
    const size_t total = off[(size_t)n_groups];
    LPS_CUDA(ctx, ctx->d_tie_off.reserve((size_t)n_groups + 1));
    LPS_CUDA(ctx, ctx->d_tie_stage.reserve(total + 1));
    LPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_tie_off.p, off.data(), 4 * ((size_t)n_groups + 1), cudaMemcpyHostToDevice, st));
    const int tb = 256;
    const unsigned grid = (unsigned)(((long long)n_groups * 32 + tb - 1) / tb);
    k_tie_copy<<<grid, tb, 0, st>>>(n_groups, ctx->d_tie_groups.p, ctx->d_tie_off.p, ctx->d_M_unsorted.p, ctx->d_tie_stage.p, 1);
    std::vector<uint32_t> stage(total);
    LPS_CUDA(ctx, cudaMemcpyAsync(stage.data(), ctx->d_tie_stage.p, 4 * total, cudaMemcpyDeviceToHost, st));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<SortRec> recs;
    for (int k = 0; k < n_groups; k++) {
        const size_t o = off[(size_t)k], m = off[(size_t)k + 1] - o;
        recs.resize(m);
        for (size_t i = 0; i < m; i++) { recs[i].position = (int)(stage[o + i] >> 2); recs[i].packed = stage[o + i]; }
        std::sort(recs.begin(), recs.end(), ByPosition());
        for (size_t i = 0; i < m; i++) stage[o + i] = recs[i].packed;
    }
    LPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_tie_stage.p, stage.data(), 4 * total, cudaMemcpyHostToDevice, st));
    k_tie_copy<<<grid, tb, 0, st>>>(n_groups, ctx->d_tie_groups.p, ctx->d_tie_off.p, ctx->d_M.p, ctx->d_tie_stage.p, 0);
    ctx->stats.kernel_launches += 2;
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    return LPS_OK;
}

int lps_launch_build_edges(lps_ctx *ctx, const lps_phase_params *p) {
    cudaStream_t st = ctx->stream;
    const int n = ctx->batch.n_reads, nv = ctx->var.n, W = p->connect_adjacent;
    const int tb = 256;
    if (W < 1 || W > 127) return ctx->fail(LPS_E_ARG, "connect_adjacent must be in [1,127]");   // last_link is an int8, the sweep packs 12-bit sums
    LPS_CUDA(ctx, ctx->d_var_lastw.reserve((size_t)nv + 1));
    LPS_CUDA(ctx, ctx->d_alive_cnt.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_aln_keys.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_aln_keys_sorted.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_node_of_var.reserve((size_t)nv + 2));
    LPS_CUDA(ctx, ctx->d_node_var.reserve((size_t)nv + 1));
    LPS_CUDA(ctx, ctx->d_node_type.reserve((size_t)nv + 1));
    LPS_CUDA(ctx, ctx->d_node_pos.reserve((size_t)nv + 66));
    LPS_CUDA(ctx, ctx->d_grp_off.reserve((size_t)n + 2));
    LPS_CUDA(ctx, ctx->d_edge_counters.reserve(4));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_var_lastw.p, 0, 8 * ((size_t)nv + 1), st));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_edge_counters.p, 0, 32, st));
    const uint8_t *erased = ctx->have_erased ? ctx->d_call_erased.p : nullptr;

    // 1. alive calls per read, node marking, alignment keys
    if (n > 0) {
        k_mark_nodes<<<(unsigned)(((long long)n * 32 + tb - 1) / tb), tb, 0, st>>>(
            n, ctx->d_call_off.p, ctx->d_calls.p, ctx->d_read_dead.p, erased, ctx->batch.name_rank,
            (unsigned long long *)ctx->d_var_lastw.p, ctx->d_alive_cnt.p, ctx->d_aln_keys.p);
        ctx->stats.kernel_launches++;
    }
    // 2. node numbering
    size_t cub_bytes = 0, need = 0;
    k_node_flags<<<(nv + tb) / tb, tb, 0, st>>>(nv, (const unsigned long long *)ctx->d_var_lastw.p, ctx->d_node_of_var.p);
    cub::DeviceScan::ExclusiveSum(nullptr, need, ctx->d_node_of_var.p, ctx->d_node_of_var.p, nv + 1, st);
    cub_bytes = need;
    cub::DeviceRadixSort::SortKeys(nullptr, need, ctx->d_aln_keys.p, ctx->d_aln_keys_sorted.p, n, 0, 64, st);
    if (need > cub_bytes) cub_bytes = need;
    cub::DeviceScan::ExclusiveSum(nullptr, need, ctx->d_grp_off.p, ctx->d_grp_off.p, n + 1, st);
    if (need > cub_bytes) cub_bytes = need;
    LPS_CUDA(ctx, ctx->d_cub_tmp.reserve(cub_bytes + 256));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_node_of_var.p + nv, 0, 4, st));
    cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, cub_bytes, ctx->d_node_of_var.p, ctx->d_node_of_var.p, nv + 1, st);
    int32_t n_nodes = 0;
    LPS_CUDA(ctx, cudaMemcpyAsync(&n_nodes, ctx->d_node_of_var.p + nv, 4, cudaMemcpyDeviceToHost, st));
    k_fill_nodes<<<(nv + tb) / tb, tb, 0, st>>>(nv, (const unsigned long long *)ctx->d_var_lastw.p, ctx->var.pos, ctx->d_node_of_var.p,
                                                ctx->d_node_var.p, ctx->d_node_pos.p, ctx->d_node_type.p);
    ctx->stats.kernel_launches += 3;
    // 3. alignments in (name rank, BAM order); offsets of their calls inside M
    int rank_bits = 32 + 32;
    cub::DeviceRadixSort::SortKeys(ctx->d_cub_tmp.p, cub_bytes, ctx->d_aln_keys.p, ctx->d_aln_keys_sorted.p, n, 0, rank_bits, st);
    k_sorted_counts<<<(n + 1 + tb) / tb, tb, 0, st>>>(n, ctx->d_aln_keys_sorted.p, ctx->d_alive_cnt.p, ctx->d_grp_off.p);
    cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, cub_bytes, ctx->d_grp_off.p, ctx->d_grp_off.p, n + 1, st);
    ctx->stats.kernel_launches += 3;
    uint64_t n_merged = 0;
    LPS_CUDA(ctx, cudaMemcpyAsync(&n_merged, ctx->d_grp_off.p + n, 8, cudaMemcpyDeviceToHost, st));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->n_nodes = n_nodes; ctx->window = W; ctx->n_merged = n_merged;
    // number of alive alignments = sorted keys that are not the sentinel: they are a prefix; count via grp_off on host is
    // avoidable: dead entries own zero calls, the kernels below simply skip them (their key is ~0).
    LPS_CUDA(ctx, ctx->d_M.reserve((size_t)n_merged + 1));
    LPS_CUDA(ctx, ctx->d_M_gend.reserve((size_t)n_merged + 1));
    LPS_CUDA(ctx, ctx->d_M_node.reserve((size_t)n_merged + 1));
    LPS_CUDA(ctx, ctx->d_M_node_sorted.reserve((size_t)n_merged + 1));
    LPS_CUDA(ctx, ctx->d_M_idx.reserve((size_t)n_merged + 1));
    LPS_CUDA(ctx, ctx->d_M_idx_sorted.reserve((size_t)n_merged + 1));
    LPS_CUDA(ctx, ctx->d_node_cnt.reserve((size_t)n_nodes + 2));
    LPS_CUDA(ctx, ctx->d_node_off.reserve((size_t)n_nodes + 2));
    LPS_CUDA(ctx, ctx->d_weights.reserve((size_t)n_nodes * (size_t)W * 4 + 4));
    LPS_CUDA(ctx, ctx->d_vote_info.reserve(((size_t)n_nodes + 2) * (size_t)lps_vote_row_stride(W) + 64));
    LPS_CUDA(ctx, ctx->d_last_link.reserve((size_t)n_nodes + 16));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_node_cnt.p, 0, 4 * ((size_t)n_nodes + 2), st));

    // alive alignments: keys != ~0 are a prefix of the sorted array
    int n_aln = 0;
    {
        // binary search on the device array would need a kernel; the host already knows which reads have calls:
        // count = reads that are not dead and have >= 1 surviving call == entries with alive_cnt > 0.
        std::vector<uint32_t> cnt((size_t)n);
        LPS_CUDA(ctx, cudaMemcpy(cnt.data(), ctx->d_alive_cnt.p, 4 * (size_t)n, cudaMemcpyDeviceToHost));
        ctx->stats.d2h_bytes += 4ull * (uint64_t)n;
        for (int r = 0; r < n; r++) if (cnt[r]) n_aln++;
    }
    if (n_aln > 0 && n_merged > 0) {
        k_fill_merged<<<(unsigned)(((long long)n_aln * 32 + tb - 1) / tb), tb, 0, st>>>(
            n_aln, ctx->d_aln_keys_sorted.p, ctx->d_grp_off.p, ctx->d_call_off.p, ctx->d_calls.p, erased, ctx->d_node_of_var.p,
            p->base_quality, ctx->d_M.p, ctx->d_M_gend.p, ctx->d_node_cnt.p);
        // multi-alignment merged reads
        DevBuf<uint2> &tie_groups = ctx->d_tie_groups;
        DevBuf<unsigned int> &n_tie = ctx->d_n_tie;
        const uint32_t tie_cap = 1u << 18;
        LPS_CUDA(ctx, tie_groups.reserve(tie_cap));
        LPS_CUDA(ctx, n_tie.reserve(1));
        LPS_CUDA(ctx, ctx->d_M_unsorted.reserve((size_t)n_merged + 1));
        LPS_CUDA(ctx, cudaMemsetAsync(n_tie.p, 0, 4, st));
        k_sort_multi_groups<<<(n_aln + tb - 1) / tb, tb, 0, st>>>(n_aln, ctx->d_aln_keys_sorted.p, ctx->d_grp_off.p, ctx->d_M.p,
                                                                  ctx->d_M_unsorted.p, tie_groups.p, tie_cap, n_tie.p);
        ctx->stats.kernel_launches += 2;
        unsigned int h_tie = 0;
        LPS_CUDA(ctx, cudaMemcpyAsync(&h_tie, n_tie.p, 4, cudaMemcpyDeviceToHost, st));
        LPS_CUDA(ctx, cudaStreamSynchronize(st));
        if (h_tie > tie_cap) return ctx->fail(LPS_E_NOMEM, "too many tied merged reads");
        if (h_tie) {
            int rc = lps_host_fix_tie_groups(ctx, (int)h_tie);
            if (rc) return rc;
        }

        // 4. per-node call lists in merged (= name rank) order: one stable radix sort by node
        k_split_merged<<<(unsigned)((n_merged + tb - 1) / tb), tb, 0, st>>>(n_merged, ctx->d_M.p, ctx->d_M_node.p, ctx->d_M_idx.p);
        size_t sort_bytes = 0;
        const int nb = bits_for((uint32_t)n_nodes);
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, ctx->d_M_node.p, ctx->d_M_node_sorted.p, ctx->d_M_idx.p,
                                        ctx->d_M_idx_sorted.p, (int)n_merged, 0, nb, st);
        LPS_CUDA(ctx, ctx->d_cub_tmp.reserve(sort_bytes + 256));
        cub::DeviceRadixSort::SortPairs(ctx->d_cub_tmp.p, sort_bytes, ctx->d_M_node.p, ctx->d_M_node_sorted.p, ctx->d_M_idx.p,
                                        ctx->d_M_idx_sorted.p, (int)n_merged, 0, nb, st);
        k_widen_u32b<<<(n_nodes + 1 + tb) / tb, tb, 0, st>>>(n_nodes, ctx->d_node_cnt.p, ctx->d_node_off.p);
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, ctx->d_node_off.p, ctx->d_node_off.p, n_nodes + 1, st);
        LPS_CUDA(ctx, ctx->d_cub_tmp.reserve(scan_bytes + 256));
        cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, scan_bytes, ctx->d_node_off.p, ctx->d_node_off.p, n_nodes + 1, st);
        ctx->stats.kernel_launches += 4;
    }
    // 5. the fold
    if (n_nodes > 0) {
        if (n_merged == 0) {
            LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_weights.p, 0, 4 * (size_t)n_nodes * W * 4, st));
            LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_vote_info.p, 0, (size_t)n_nodes * lps_vote_row_stride(W), st));
            LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_last_link.p, 0xFF, (size_t)n_nodes, st));
        } else {
            constexpr int WARPS = 8;
            size_t smem = (size_t)WARPS * W * 4 * sizeof(float);
            cudaEventRecord(ctx->kev[2], st);
            k_fold_edges<WARPS><<<(n_nodes + WARPS - 1) / WARPS, WARPS * 32, smem, st>>>(
                n_nodes, W, p->edge_weight, ctx->d_node_off.p, ctx->d_M_idx_sorted.p, ctx->d_M.p, ctx->d_M_gend.p, ctx->d_weights.p,
                p->edge_threshold, ctx->d_vote_info.p, (unsigned long long *)ctx->d_edge_counters.p, (uint64_t)n_merged, ctx->d_last_link.p, lps_vote_row_stride(W));
            cudaEventRecord(ctx->kev[3], st);
            ctx->stats.kernel_launches++;
        }
    }
    LPS_CUDA(ctx, cudaGetLastError());
    unsigned long long hc[2] = {0, 0};
    LPS_CUDA(ctx, cudaMemcpyAsync(hc, ctx->d_edge_counters.p, 16, cudaMemcpyDeviceToHost, st));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    if (n_nodes > 0 && n_merged > 0) cudaEventElapsedTime(&ctx->stats.ms_kernel_fold_edges, ctx->kev[2], ctx->kev[3]);
    ctx->n_contrib = hc[0]; ctx->n_contrib_far = hc[1];
    ctx->have_graph = true;
    return LPS_OK;
}
