// k_sweep.cu — VairiantGraph::edgeConnectResult on the device (reference
// src/phase/PhasingGraph.cpp:286-474 with findBestEdgePair :166-228 and Onelongcase :251-283).
//
// The sweep is a left-to-right chain: node k's haplotype comes from the weighted votes of its <= W
// predecessors, and k then votes on its W successors.  Nothing about the votes themselves depends on
// the chain, so k_fold_edges' epilogue already reduced every (node, successor) cell to one byte
// (link direction, weight class, "single read" flag, "ESR < 0.2" flag).  What is left is inherently
// sequential, so ONE WARP walks the nodes:
//   * the vote accumulators of the next 64 nodes live in REGISTERS (lane t%32 owns nodes t and t+32 of
//     a 64-slot ring), so a step is: 5 shuffles to broadcast node k's accumulators, a few compares,
//     then every lane adds node k's vote to the one or two successors it owns — no shared memory, no
//     barrier; the vote bytes of node k+1 are already in flight;
//   * hpCountMap2 is a float sum in voter order (weights 1, 20 and 0.1f): each successor receives its
//     votes one voter at a time in ascending voter order, exactly like the reference;
//   * block bookkeeping collapses: inside a block the REF-allele haplotype telescopes to hp[k]-1
//     (the first member of a block always has hp 1), and PS = position(block start) + 1 unless the
//     block has a single member.
// Keeping the sweep on the device means the [nodes][W][4] float table never crosses PCIe.
#include "lps_ctx.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct Acc { float w1, w2, s1, s2; int singles; };

__device__ __forceinline__ void acc_clear(Acc &a) { a.w1 = a.w2 = a.s1 = a.s2 = 0.f; a.singles = 0; }

__device__ __forceinline__ void acc_vote(Acc &a, unsigned info, int hp, unsigned type) {
    const unsigned link = info & 3u;
    if (!link) return;
    const float weight = type == 4u ? (float)0.1 : ((info & 4u) ? 20.f : 1.f);
    const bool to_h1 = (hp == 1) == (link == 1u);
    if (to_h1) a.w1 += weight; else a.w2 += weight;
    if (info & 8u) a.singles++;                                               // (para + cross) <= 1
    else if ((info & 16u) && weight >= 1.f && type != 3u) {                   // Onelongcase's second branch
        if (to_h1) a.s1 += weight; else a.s2 += weight;
    }
}

__global__ void __launch_bounds__(32) k_sweep(int N, int W, int distance, const int32_t *__restrict__ node_var,
                                              const int32_t *__restrict__ vpos, const uint8_t *__restrict__ node_type,
                                              const uint8_t *__restrict__ vote_info, int32_t *__restrict__ node_ps,
                                              int8_t *__restrict__ node_hap) {
    const int lane = threadIdx.x;
    for (int k = lane; k < N; k += 32) { node_ps[k] = 0; node_hap[k] = -1; }
    __syncwarp();
    if (N < 2) return;
    Acc A, B;                       // ring slots `lane` and `lane + 32`
    acc_clear(A); acc_clear(B);
    int block_start = -1, block_first_size = 0;   // current block and its member count
    int last_connect = -1;

    // vote bytes of node k for the one or two successors this lane owns
    auto load_info = [&](int k, unsigned &i1, unsigned &i2, int &t1) {
        t1 = k + 1 + ((lane - (k + 1)) & 31);
        const int d1 = t1 - k - 1, d2 = d1 + 32;
        i1 = (d1 < W && t1 < N) ? vote_info[(size_t)k * W + d1] : 0u;
        i2 = (d2 < W && t1 + 32 < N) ? vote_info[(size_t)k * W + d2] : 0u;
    };
    unsigned ni1, ni2; int nt1;
    load_info(0, ni1, ni2, nt1);
    int pos_k = vpos[node_var[0]], pos_k1 = vpos[node_var[1]];
    unsigned type_k = node_type[0];
    for (int k = 0; k + 1 < N; k++) {
        const unsigned i1 = ni1, i2 = ni2;
        const int t1 = nt1;
        const int cur_pos = pos_k, nxt_pos = pos_k1;
        const unsigned type = type_k;
        if (k + 2 < N) {
            load_info(k + 1, ni1, ni2, nt1);
            pos_k = pos_k1; pos_k1 = vpos[node_var[k + 2]];
            type_k = node_type[k + 1];
        }
        // ---- node k's accumulated votes: owned by lane k%32, slot A or B ----
        const int src = k & 31;
        const bool useB = (k & 32) != 0;
        float h1 = __shfl_sync(FULL, useB ? B.w1 : A.w1, src);
        float h2 = __shfl_sync(FULL, useB ? B.w2 : A.w2, src);
        const float s1 = __shfl_sync(FULL, useB ? B.s1 : A.s1, src);
        const float s2 = __shfl_sync(FULL, useB ? B.s2 : A.s2, src);
        const int singles = __shfl_sync(FULL, useB ? B.singles : A.singles, src);
        if (lane == src) { if (useB) acc_clear(B); else acc_clear(A); }      // the slot now belongs to node k + 64
        const int gap = nxt_pos - cur_pos;
        bool enter = !((gap < 0 ? -gap : gap) > distance);                     // :318-320
        int hp = 0;
        if (enter) {
            if (!(singles <= 3 || (s1 == 0.f && s2 == 0.f))) { h1 = s1; h2 = s2; }   // Onelongcase :276-281
            if (h1 == h2) {
                if (last_connect >= 0 && k < last_connect) enter = false;       // :340-342 (positions ascend with k)
                else {
                    // close the previous block: a block with a single member is not a block (:425)
                    if (lane == 0 && block_start >= 0 && block_first_size == 1) { node_ps[block_start] = 0; node_hap[block_start] = -1; }
                    block_start = k; block_first_size = 0; hp = 1;
                }
            } else hp = h1 > h2 ? 1 : 2;
        }
        if (enter) {
            block_first_size++;
            if (lane == 0) {
                node_ps[k] = vpos[node_var[block_start]] + 1;
                node_hap[k] = (int8_t)(hp - 1);
            }
            // ---- vote on the successors this lane owns ----
            if ((t1 & 32) == 0) acc_vote(A, i1, hp, type); else acc_vote(B, i1, hp, type);
            if (((t1 + 32) & 32) == 0) acc_vote(A, i2, hp, type); else acc_vote(B, i2, hp, type);
            // lastConnectPos = the farthest successor this node connected to (assigned in ascending order)
            int far = -1;
            if (i1 & 3u) far = t1;
            if (i2 & 3u) far = t1 + 32;
#pragma unroll
            for (int d = 16; d; d >>= 1) far = max(far, __shfl_xor_sync(FULL, far, d));
            if (far >= 0) last_connect = far;
        }
    }
    if (lane == 0 && block_start >= 0 && block_first_size == 1) { node_ps[block_start] = 0; node_hap[block_start] = -1; }
}

__global__ void k_expand_nodes(int nv, int N, const int32_t *__restrict__ node_of_var, const int32_t *__restrict__ node_ps,
                               const int8_t *__restrict__ node_hap, int32_t *__restrict__ ps, int8_t *__restrict__ hap) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    int k = node_of_var[i];
    ps[i] = k >= 0 ? node_ps[k] : 0;
    hap[i] = k >= 0 ? node_hap[k] : (int8_t)-1;
}

}  // namespace

// writes the sweep result per VARIANT into d_ps / d_hap_ref
int lps_launch_sweep(lps_ctx *ctx, const lps_phase_params *p) {
    cudaStream_t st = ctx->stream;
    const int N = ctx->n_nodes, W = ctx->window, nv = ctx->var.n;
    if (W > 63) return ctx->fail(LPS_E_ARG, "the device sweep supports connect_adjacent <= 63");
    LPS_CUDA(ctx, ctx->d_node_ps.reserve((size_t)N + 1));
    LPS_CUDA(ctx, ctx->d_node_hap.reserve((size_t)N + 1));
    LPS_CUDA(ctx, ctx->d_ps.reserve((size_t)nv + 1));
    LPS_CUDA(ctx, ctx->d_hap_ref.reserve((size_t)nv + 1));
    cudaEventRecord(ctx->kev[4], st);
    k_sweep<<<1, 32, 0, st>>>(N, W, p->distance, ctx->d_node_var.p, ctx->var.pos, ctx->d_node_type.p, ctx->d_vote_info.p,
                              ctx->d_node_ps.p, ctx->d_node_hap.p);
    cudaEventRecord(ctx->kev[5], st);
    if (nv > 0)
        k_expand_nodes<<<(nv + 255) / 256, 256, 0, st>>>(nv, N, ctx->d_node_of_var.p, ctx->d_node_ps.p, ctx->d_node_hap.p, ctx->d_ps.p,
                                                         ctx->d_hap_ref.p);
    ctx->stats.kernel_launches += 2;
    LPS_CUDA(ctx, cudaGetLastError());
    return LPS_OK;
}
