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
    // sixteen lanes per read (a read carries ~18 calls); both reads of a warp stay to the end (full-mask shuffles)
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    int lane = threadIdx.x & 15;
    const bool present = wid < n_reads;
    int r = present ? (int)wid : 0;
    uint64_t c0 = present ? call_off[r] : 0, c1 = present ? call_off[r + 1] : 0;
    int alive = 0;
    if (present && !read_dead[r]) {
        for (uint64_t c = c0 + lane; c < c1; c += 16) {
            if (call_erased && call_erased[c]) continue;
            lps_call cl = calls[c];
            unsigned type = cl.quality == -4 ? 3u : (cl.quality == -5 ? 4u : 0u);
            atomicMax(&var_lastw[cl.var], ((unsigned long long)(r + 1) << 3) | type);
            alive++;
        }
    }
#pragma unroll
    for (int d = 8; d; d >>= 1) alive += __shfl_xor_sync(FULL, alive, d);
    if (present && lane == 0) {
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
    // sixteen lanes per alignment (~18 calls each); the two alignments of a warp loop in lockstep (full-mask ballots)
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15, hshift = threadIdx.x & 16;
    int i = wid < n_aln ? (int)wid : 0;
    uint64_t key = wid < n_aln ? keys_sorted[i] : ~0ull;
    const bool live = key != ~0ull;                                             // alignments without an alive call sort to the end
    uint32_t rank = (uint32_t)(key >> 32);
    int r = live ? (int)(uint32_t)key : 0;
    // end of the merged group: first later alignment with another name
    int j = i + 1;
    while (live && j < n_aln && (uint32_t)(keys_sorted[j] >> 32) == rank) j++;
    uint32_t gend = live ? (uint32_t)grp_off[j] : 0u;
    uint64_t c0 = live ? call_off[r] : 0, c1 = live ? call_off[r + 1] : 0;
    uint64_t w = live ? grp_off[i] : 0;
    for (uint64_t cb = c0; __any_sync(FULL, cb < c1); cb += 16) {
        uint64_t c = cb + lane;
        bool ok = c < c1 && !(call_erased && call_erased[c]);
        unsigned m = (__ballot_sync(FULL, ok) >> hshift) & 0xFFFFu;             // this alignment's sixteen lanes
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
// ---- libstdc++'s std::sort, replayed: merged reads with tied positions and more than 16 calls -------------------------------
// ReadVariant::sort (Util.cpp:3-5) is std::sort by position; with equal positions the order of the tied calls is whatever
// introsort leaves, and the edge weights are folded in that order.  The algorithm is deterministic, so it is restated here step for
// step (GCC's bits/stl_algo.h: __introsort_loop with the median-of-three pivot moved to the front and the unguarded partition,
// then __final_insertion_sort: guarded insertion over the first 16, unguarded over the rest).  The heapsort branch behind the
// depth limit 2 * floor(log2(n)) is not restated: a group that reaches it is reported to the host instead (never seen so far).
// Keys are the merged entries themselves (node << 2 | allele << 1 | q_hi); the comparator looks at the node only.
__device__ __forceinline__ bool ss_less(uint32_t a, uint32_t b) { return (a >> 2) < (b >> 2); }
__device__ __forceinline__ void ss_swap(uint32_t *v, long long i, long long j) { const uint32_t t = v[i]; v[i] = v[j]; v[j] = t; }

__device__ bool std_sort_replay(uint32_t *v, long long n) {
    if (n < 2) return true;
    // __introsort_loop, the recursion on the right part turned into an explicit stack of (first, last, depth_limit)
    long long stk_first[48], stk_last[48];
    int stk_depth[48];
    int top = 0;
    int lg = 0;
    for (long long t = n; t > 1; t >>= 1) lg++;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * lg; top = 1;
    while (top > 0) {
        top--;
        long long first = stk_first[top], last = stk_last[top];
        int depth = stk_depth[top];
        // the reference implementation recurses into [cut, last) FIRST and then loops on [first, cut); the parts are disjoint, so the
        // order in which they are finished does not change the result
        while (last - first > 16) {
            if (depth == 0) return false;                       // would switch to heapsort
            depth--;
            // __unguarded_partition_pivot
            const long long mid = first + (last - first) / 2;
            {   // __move_median_to_first(first, first + 1, mid, last - 1)
                const long long a = first + 1, b = mid, c = last - 1;
                if (ss_less(v[a], v[b])) {
                    if (ss_less(v[b], v[c])) ss_swap(v, first, b);
                    else if (ss_less(v[a], v[c])) ss_swap(v, first, c);
                    else ss_swap(v, first, a);
                } else if (ss_less(v[a], v[c])) ss_swap(v, first, a);
                else if (ss_less(v[b], v[c])) ss_swap(v, first, c);
                else ss_swap(v, first, b);
            }
            long long lo = first + 1, hi = last;
            const long long pivot = first;
            while (true) {                                      // __unguarded_partition(first + 1, last, first)
                while (ss_less(v[lo], v[pivot])) lo++;
                hi--;
                while (ss_less(v[pivot], v[hi])) hi--;
                if (!(lo < hi)) break;
                ss_swap(v, lo, hi);
                lo++;
            }
            const long long cut = lo;
            if (top >= 48) return false;
            stk_first[top] = cut; stk_last[top] = last; stk_depth[top] = depth; top++;
            last = cut;
        }
    }
    // __final_insertion_sort
    auto unguarded_linear_insert = [&](long long lastp) {
        const uint32_t val = v[lastp];
        long long next = lastp - 1;
        while (ss_less(val, v[next])) { v[lastp] = v[next]; lastp = next; next--; }
        v[lastp] = val;
    };
    auto insertion_sort = [&](long long f, long long l) {
        if (f == l) return;
        for (long long i = f + 1; i != l; i++) {
            if (ss_less(v[i], v[f])) {
                const uint32_t val = v[i];
                for (long long k = i; k > f; k--) v[k] = v[k - 1];    // move_backward(first, i, i + 1)
                v[f] = val;
            } else unguarded_linear_insert(i);
        }
    };
    if (n > 16) {
        insertion_sort(0, 16);
        for (long long i = 16; i != n; i++) unguarded_linear_insert(i);
    } else insertion_sort(0, n);
    return true;
}

constexpr int SMG_WARPS = 8, SMG_CAP = 1024;      // merged reads of up to SMG_CAP calls are sorted in shared memory

// in-place insertion sort of concatenated sorted runs (equal positions keep their concatenation order)
__device__ __forceinline__ void insertion_sort_runs(uint32_t *v, long long n) {
    for (long long a = 1; a < n; a++) {
        const uint32_t x = v[a];
        long long b = a;
        while (b > 0 && (v[b - 1] >> 2) > (x >> 2)) { v[b] = v[b - 1]; b--; }
        v[b] = x;
    }
}

// position of every alive alignment in the (name rank, BAM order) list
__global__ void k_sorted_position(int n, const uint64_t *__restrict__ keys_sorted, int32_t *__restrict__ pos_of_read) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = keys_sorted[i];
    if (k != ~0ull) pos_of_read[(uint32_t)k] = i;
}

// One WARP per read name that owns several alignments of the batch (the groups the host indexed at submit time,
// lps_host_index_names): its alive alignments are neighbours in the sorted list, so the merged read is M[grp_off[first] ..
// grp_off[first + alive]).  The group is copied to shared memory by the whole warp (it is small and the work on it is a dependent
// chain: in global memory every step waited for L2), lane 0 orders it, the warp writes it back.
__global__ void __launch_bounds__(SMG_WARPS * 32) k_sort_multi_groups(int n_groups, const int32_t *__restrict__ group_off,
                                    const int32_t *__restrict__ members, const uint32_t *__restrict__ alive_cnt,
                                    const int32_t *__restrict__ pos_of_read, const uint64_t *__restrict__ grp_off,
                                    uint32_t *__restrict__ M, uint32_t *__restrict__ M_unsorted,
                                    uint2 *__restrict__ tie_groups, uint32_t tie_cap, unsigned int *__restrict__ n_tie, int record_only) {
    __shared__ uint32_t s_buf[SMG_WARPS][SMG_CAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = blockIdx.x * SMG_WARPS + wib;
    if (g >= n_groups) return;
    // first position and number of the group's alive alignments
    int first = INT_MAX, alive = 0;
    for (int m = group_off[g] + lane; m < group_off[g + 1]; m += 32) {
        const int r = members[m];
        if (alive_cnt[r]) { alive++; first = min(first, pos_of_read[r]); }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { alive += __shfl_xor_sync(FULL, alive, d); first = min(first, __shfl_xor_sync(FULL, first, d)); }
    if (alive < 2) return;                                                       // a single alignment is sorted already
    const uint64_t g0 = grp_off[first], g1 = grp_off[first + alive];
    const long long n = (long long)(g1 - g0);
    if (n < 2) return;
    const bool in_smem = n <= SMG_CAP;
    uint32_t *v = in_smem ? s_buf[wib] : M + g0;
    if (!record_only) for (long long a = lane; a < n; a += 32) M_unsorted[g0 + a] = M[g0 + a];      // concatenation order, for the replays
    if (in_smem) for (long long a = lane; a < n; a += 32) v[a] = M[g0 + a];
    __syncwarp();
    bool report = false;
    if (in_smem) {
        // Stable rank sort by position, all lanes: an element's place is the number of smaller keys plus the equal keys in front of
        // it (= what the insertion sort of the concatenated runs leaves).  One lane's insertion sort took up to 58 us per contig: a
        // supplementary alignment that maps in front of its primary moves every call past every other one, a shared-memory round trip each.
        bool tie_l = false;
        for (long long a = lane; a < n; a += 32) {
            const uint32_t x = v[a], key = x >> 2;
            int less = 0, eq_before = 0;
            for (long long c = 0; c < n; c++) {
                const uint32_t kc = v[c] >> 2;
                less += kc < key ? 1 : 0;
                eq_before += (kc == key && c < a) ? 1 : 0;
            }
            if (eq_before) tie_l = true;                       // a tie = two calls of the merged read at one position
            if (!record_only) M[g0 + less + eq_before] = x;
        }
        const bool tie = __any_sync(FULL, tie_l);
        if (tie && n > 16) {
            // more than 16 calls with a tie: the order of the tied calls is std::sort's - replay it from the concatenation order
            int done = 0;
            if (!record_only) {
                __syncwarp();
                for (long long a = lane; a < n; a += 32) v[a] = M_unsorted[g0 + a];
                __syncwarp();
                if (lane == 0) done = std_sort_replay(v, n) ? 1 : 0;
                done = __shfl_sync(FULL, done, 0);
                __syncwarp();
                if (done) for (long long a = lane; a < n; a += 32) M[g0 + a] = v[a];
                // else M keeps the position-sorted order for the host replay's write-back
            }
            report = !done;
        }
    } else if (lane == 0) {
        if (!record_only) insertion_sort_runs(v, n);
        // a tie = two calls of the merged read at one position: adjacent after the sort
        bool tie = false;
        for (long long a = 1; a < n && !tie; a++) tie = (v[a - 1] >> 2) == (v[a] >> 2);
        if (tie && n > 16) {
            bool done = false;
            if (!record_only) {
                for (long long a = 0; a < n; a++) v[a] = M_unsorted[g0 + a];
                done = std_sort_replay(v, n);
                if (!done) {                                     // leave the group position-sorted for the host replay's write-back
                    for (long long a = 0; a < n; a++) v[a] = M_unsorted[g0 + a];
                    insertion_sort_runs(v, n);
                }
            }
            report = !done;
        }
    }
    __syncwarp();
    report = __any_sync(FULL, report);
    if (lane == 0 && report) {
        const unsigned k = atomicAdd(n_tie, 1u);
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

// n_upper >= number of merged calls (read from device memory): the tail gets the key `pad_node`, which sorts behind every node
__global__ void k_split_merged(uint64_t n_upper, const uint64_t *__restrict__ n_merged, uint32_t pad_node, const uint32_t *__restrict__ M,
                               uint32_t *__restrict__ node, uint32_t *__restrict__ idx) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_upper) return;
    node[i] = i < *n_merged ? (M[i] >> 2) : pad_node;
    idx[i] = (uint32_t)i;
}

__global__ void k_widen_u32b(int n, const uint32_t *__restrict__ in, uint64_t *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = i < n ? in[i] : 0;
}

// the fold: one warp per node
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_fold_edges(const int32_t *__restrict__ n_nodes_ptr, int W, double edge_weight,
                                                          const uint64_t *__restrict__ node_off,
                                                          const uint32_t *__restrict__ list, const uint32_t *__restrict__ M,
                                                          const uint32_t *__restrict__ M_gend, float *__restrict__ weights,
                                                          double edge_threshold, uint8_t *__restrict__ vote_info,
                                                          unsigned long long *__restrict__ counters, const uint64_t *__restrict__ n_merged_ptr,
                                                          int8_t *__restrict__ last_link, int RS, const int32_t *__restrict__ node_pos,
                                                          const uint8_t *__restrict__ node_type, int distance, uint16_t *__restrict__ sweep_meta) {
    extern __shared__ float s_acc[];                       // [WARPS][W*4]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wid = (long long)blockIdx.x * WARPS + wib;
    const int n_nodes = *n_nodes_ptr;
    const uint64_t n_merged = *n_merged_ptr;
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
        const uint32_t t0 = (uint32_t)l0, t1 = (uint32_t)l1;       // 32-bit loop indices: the 64-bit index arithmetic was a fifth of the kernel's instructions
        const uint32_t nm = (uint32_t)n_merged;            // merged entries are indexed with 32 bits (d_M_idx is uint32)
        auto fetch = [&](uint32_t m, uint32_t &ea, uint32_t &gend, uint32_t &e0, uint32_t &e1) {
            ea = M[m]; gend = M_gend[m];
            const uint32_t i0 = m + 1u + (uint32_t)lane, i1 = i0 + 32u;
            e0 = (lane < W && i0 < nm) ? M[i0] : 0xffffffffu;
            e1 = (lane < W2 && i1 < nm) ? M[i1] : 0xffffffffu;
        };
        unsigned c32 = 0, f32 = 0;                         // per-lane counts of this node (<= 2 per call), widened once at the end
        const float wlo = (float)edge_weight;
        (void)wlo;
        auto fold_call = [&](const uint32_t m, const uint32_t ea_c, const uint32_t gend_c, uint32_t eb0, uint32_t eb1) {
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
            const int cell0 = d0 * 4 + (int)(al_a * 2u + ((eb0 >> 1) & 1u));        // offset into the tile (only used when dense0)
            const bool high0 = hi_a && (eb0 & 1u);
            if (!second) {
                if (!__any_sync(FULL, dup)) {
                    // distinct successors of one read hit distinct cells
                    if (dense0) { float x = acc[cell0]; x = high0 ? x + 1.0f : (float)((double)x + edge_weight); acc[cell0] = x; }   // SubEdge::addSubEdge :40-43, :62-65
                } else {
                    // keep the read's own order
                    for (int l = 0; l < 32; l++) {
                        if (lane == l && dense0) { float x = acc[cell0]; x = high0 ? x + 1.0f : (float)((double)x + edge_weight); acc[cell0] = x; }
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
                const int cell1 = d1 * 4 + (int)(al_a * 2u + ((eb1 >> 1) & 1u));
                const bool high1 = hi_a && (eb1 & 1u);
                if (!any_dup) {
                    if (dense0) { float x = acc[cell0]; x = high0 ? x + 1.0f : (float)((double)x + edge_weight); acc[cell0] = x; }
                    if (dense1) { float x = acc[cell1]; x = high1 ? x + 1.0f : (float)((double)x + edge_weight); acc[cell1] = x; }
                } else {
                    // keep the read's own order: successors 0..31, then 32..W-1
                    for (int l = 0; l < 32; l++) {
                        if (lane == l && dense0) { float x = acc[cell0]; x = high0 ? x + 1.0f : (float)((double)x + edge_weight); acc[cell0] = x; }
                        __syncwarp();
                    }
                    for (int l = 0; l < W2; l++) {
                        if (lane == l && dense1) { float x = acc[cell1]; x = high1 ? x + 1.0f : (float)((double)x + edge_weight); acc[cell1] = x; }
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
        };
        // (unrolling by two with two register sets, to save the moves of the hand-over, changed nothing: 0.223 ms either way)
        uint32_t mA = 0, eaA = 0, gendA = 0, e0A = 0xffffffffu, e1A = 0xffffffffu, mB = 0;
        if (t0 < t1) { mA = list[t0]; fetch(mA, eaA, gendA, e0A, e1A); }
        if (t0 + 1u < t1) mB = list[t0 + 1u];
        for (uint32_t t = t0; t < t1; t++) {
            const uint32_t m = mA, ea_c = eaA, gend_c = gendA, b0 = e0A, b1 = e1A;
            if (t + 1u < t1) {
                mA = mB;
                fetch(mA, eaA, gendA, e0A, e1A);
                if (t + 2u < t1) mB = list[t + 2u];
            }
            fold_call(m, ea_c, gend_c, b0, b1);
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
    if (lane == 0) {
        last_link[a] = (int8_t)last;
        // what the sweep needs to know about node a besides its votes: type, "the next node is farther than `distance`" (:318-320), last link
        int gap = 0;
        if (a + 1 < n_nodes) { const int d = node_pos[a + 1] - node_pos[a]; gap = (d < 0 ? -d : d) > distance; }
        sweep_meta[a] = (uint16_t)((unsigned)node_type[a] | (gap ? 8u : 0u) | ((unsigned)(last + 1) << 8));
    }
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

// The fold with SIXTEEN lanes per node, two nodes per warp (connect_adjacent <= 48, i.e. at most three rounds of 16 successors per
// call).  A merged read carries ~18 calls, so a call has ~9 successors on average: with 32 lanes per node two thirds of the lanes
// idled and the kernel, which is bound by instruction issue, paid a full warp instruction stream per node.  Same order of the float
// additions as k_fold_edges: calls of a node in name-rank order (one per loop trip), the successors of one call hit distinct cells
// unless the read has two calls at one position (then in the read's own order, lane by lane).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_fold_edges16(const int32_t *__restrict__ n_nodes_ptr, int W, double edge_weight,
                                                            const uint64_t *__restrict__ node_off,
                                                            const uint32_t *__restrict__ list, const uint32_t *__restrict__ M,
                                                            const uint32_t *__restrict__ M_gend, float *__restrict__ weights,
                                                            double edge_threshold, uint8_t *__restrict__ vote_info,
                                                            unsigned long long *__restrict__ counters, const uint64_t *__restrict__ n_merged_ptr,
                                                            int8_t *__restrict__ last_link, int RS, const int32_t *__restrict__ node_pos,
                                                            const uint8_t *__restrict__ node_type, int distance, uint16_t *__restrict__ sweep_meta) {
    extern __shared__ float s_acc[];                       // [WARPS * 2][W*4]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
    const int n_nodes = *n_nodes_ptr;
    const uint32_t nm = (uint32_t)*n_merged_ptr;           // merged entries are indexed with 32 bits (d_M_idx is uint32)
    const long long first = ((long long)blockIdx.x * WARPS + wib) * 2;
    if (first >= n_nodes) return;                          // the whole warp
    const bool valid = first + half < n_nodes;
    const int a = valid ? (int)(first + half) : 0;
    float *acc = s_acc + ((size_t)wib * 2 + half) * W * 4;
    for (int i = hl; i < W * 4; i += 16) acc[i] = 0.0f;
    __syncwarp();
    uint32_t t = valid ? (uint32_t)node_off[a] : 0u;
    const uint32_t t1 = valid ? (uint32_t)node_off[a + 1] : 0u;
    const unsigned hshift = 16u * (unsigned)half;
    auto fetch = [&](uint32_t m, uint32_t &ea, uint32_t &gend, uint32_t &e0, uint32_t &e1) {
        ea = M[m]; gend = M_gend[m];
        const uint32_t i0 = m + 1u + (uint32_t)hl, i1 = i0 + 16u;
        e0 = (hl < W && i0 < nm) ? M[i0] : 0xffffffffu;
        e1 = (16 + hl < W && i1 < nm) ? M[i1] : 0xffffffffu;
    };
    auto add = [&](int cell, bool high) {                  // SubEdge::addSubEdge :40-43, :62-65
        float x = acc[cell];
        x = high ? x + 1.0f : (float)((double)x + edge_weight);
        acc[cell] = x;
    };
    unsigned c32 = 0, f32 = 0;                             // per-lane counts of this node, widened once at the end
    uint32_t mA = 0, eaA = 0, gendA = 0, e0A = 0xffffffffu, e1A = 0xffffffffu, mB = 0;
    bool live = t < t1;
    if (live) { mA = list[t]; fetch(mA, eaA, gendA, e0A, e1A); }
    if (t + 1u < t1) mB = list[t + 1u];
    while (__any_sync(FULL, live)) {
        const uint32_t m = mA, ea_c = eaA, gend_c = gendA;
        uint32_t eb0 = e0A, eb1 = e1A;
        if (live && t + 1u < t1) {                         // the words of the next call are requested before this one is folded
            mA = mB;
            fetch(mA, eaA, gendA, e0A, e1A);
            if (t + 2u < t1) mB = list[t + 2u];
        }
        const unsigned al_a = (ea_c >> 1) & 1u, hi_a = ea_c & 1u;
        const bool act0 = live && hl < W && m + 1u + (uint32_t)hl < gend_c;
        const bool act1 = live && 16 + hl < W && m + 17u + (uint32_t)hl < gend_c;
        const bool act2 = live && 32 + hl < W && m + 33u + (uint32_t)hl < gend_c;
        if (!act0) eb0 = 0xffffffffu;
        const int nb0 = (int)(eb0 >> 2);
        // duplicates (two calls of one merged read at the same position) are adjacent in the read's sorted list
        const uint32_t p0 = __shfl_up_sync(FULL, eb0 >> 2, 1, 16);
        const bool dup0 = act0 && hl > 0 && p0 == (uint32_t)nb0;
        const int d0 = nb0 - a - 1;
        const bool dense0 = act0 && d0 >= 0 && d0 < W;
        c32 += (unsigned)dense0;
        f32 += (unsigned)(act0 && !dense0);
        const int cell0 = d0 * 4 + (int)(al_a * 2u + ((eb0 >> 1) & 1u));
        const bool high0 = hi_a && (eb0 & 1u);
        // warp-uniform: does any call of the two reach beyond 16 successors, does any have a duplicate among its first 16?
        const bool beyond = __any_sync(FULL, act1);
        const unsigned dupb0 = __ballot_sync(FULL, dup0);
        if (!beyond && dupb0 == 0u) {
            if (dense0) add(cell0, high0);
        } else {
            if (!act1) eb1 = 0xffffffffu;
            const int nb1 = (int)(eb1 >> 2);
            const uint32_t p1 = __shfl_up_sync(FULL, eb1 >> 2, 1, 16), l0 = __shfl_sync(FULL, eb0 >> 2, 15, 16);
            const bool dup1 = act1 && (hl > 0 ? p1 : l0) == (uint32_t)nb1;
            const int d1 = nb1 - a - 1;
            const bool dense1 = act1 && d1 >= 0 && d1 < W;
            c32 += (unsigned)dense1;
            f32 += (unsigned)(act1 && !dense1);
            const int cell1 = d1 * 4 + (int)(al_a * 2u + ((eb1 >> 1) & 1u));
            const bool high1 = hi_a && (eb1 & 1u);
            const unsigned dupb1 = dupb0 | __ballot_sync(FULL, dup1);
            if (!__any_sync(FULL, act2) && dupb1 == 0u) {
                // 17..32 successors, all distinct: distinct cells
                if (dense0) add(cell0, high0);
                if (dense1) add(cell1, high1);
            } else {
                uint32_t eb2 = 0xffffffffu;
                if (act2) eb2 = M[m + 33u + (uint32_t)hl];     // successors 33.. of a call: rare, loaded on demand
                const int nb2 = (int)(eb2 >> 2);
                const uint32_t p2 = __shfl_up_sync(FULL, eb2 >> 2, 1, 16), l1 = __shfl_sync(FULL, eb1 >> 2, 15, 16);
                const bool dup2 = act2 && (hl > 0 ? p2 : l1) == (uint32_t)nb2;
                const int d2 = nb2 - a - 1;
                const bool dense2 = act2 && d2 >= 0 && d2 < W;
                c32 += (unsigned)dense2;
                f32 += (unsigned)(act2 && !dense2);
                const int cell2 = d2 * 4 + (int)(al_a * 2u + ((eb2 >> 1) & 1u));
                const bool high2 = hi_a && (eb2 & 1u);
                const unsigned dupb = dupb1 | __ballot_sync(FULL, dup2);
                const bool my_dup = ((dupb >> hshift) & 0xFFFFu) != 0u;      // this node's call has a duplicate
                if (!my_dup) {
                    if (dense0) add(cell0, high0);
                    if (dense1) add(cell1, high1);
                    if (dense2) add(cell2, high2);
                }
                if (dupb != 0u) {
                    // keep the read's own order: successors 0..15, 16..31, 32..
                    for (int l = 0; l < 16; l++) { if (my_dup && hl == l && dense0) add(cell0, high0); __syncwarp(); }
                    for (int l = 0; l < 16; l++) { if (my_dup && hl == l && dense1) add(cell1, high1); __syncwarp(); }
                    for (int l = 0; l < 16; l++) { if (my_dup && hl == l && dense2) add(cell2, high2); __syncwarp(); }
                }
            }
        }
        __syncwarp();
        if (live) { t++; live = t < t1; }
    }
    unsigned long long contrib = c32, far = f32;
    __syncwarp();
    if (valid) {
        float *out = weights + (size_t)a * W * 4;
        for (int i = hl; i < W * 4; i += 16) out[i] = acc[i];
    }
    // epilogue: everything VariantEdge::findBestEdgePair (:166-228) derives from a cell, one byte per successor (see k_fold_edges)
    int last = -1;
    const int shift = (a + 1) & 15;
    if (valid) {
        for (int j = hl; j < RS; j += 16) {
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
    }
#pragma unroll
    for (int dd = 8; dd; dd >>= 1) last = max(last, __shfl_xor_sync(FULL, last, dd));      // inside the 16 lanes of the node
    if (valid && hl == 0) {
        last_link[a] = (int8_t)last;
        int gap = 0;
        if (a + 1 < n_nodes) { const int d = node_pos[a + 1] - node_pos[a]; gap = (d < 0 ? -d : d) > distance; }
        sweep_meta[a] = (uint16_t)((unsigned)node_type[a] | (gap ? 8u : 0u) | ((unsigned)(last + 1) << 8));
    }
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

// first / last called position of the listed reads (what the overlap filter looks at)
namespace {
__global__ void k_first_last(int m, const int32_t *__restrict__ reads, const uint64_t *__restrict__ call_off,
                             const lps_call *__restrict__ calls, const int32_t *__restrict__ vpos, int32_t *__restrict__ first_pos,
                             int32_t *__restrict__ last_pos, uint32_t *__restrict__ ncalls) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int r = reads[i];
    uint64_t c0 = call_off[r], c1 = call_off[r + 1];
    first_pos[i] = c1 > c0 ? vpos[calls[c0].var] : -1;
    last_pos[i] = c1 > c0 ? vpos[calls[c1 - 1].var] : -1;
    ncalls[i] = (uint32_t)(c1 - c0);
}

// The overlap filter at the head of VairiantGraph::addEdge (PhasingGraph.cpp:707-781) among the alignments of ONE read name: a
// small state machine per name, so one thread per name that owns more than one alignment of the batch (the names are grouped on
// the host at submit time, lps_host_index_names).  alignRange[name] is inserted as {0, 0} before the find() (:712-716), so its
// .first is always 0; `kept` (readIdxVec[name]) is a stack in the scratch array parallel to the members.
__global__ void k_overlap_filter(int n_groups, const int32_t *__restrict__ group_off, const int32_t *__restrict__ members,
                                 const int32_t *__restrict__ first_pos, const int32_t *__restrict__ last_pos,
                                 const uint32_t *__restrict__ ncalls, double overlap_threshold, int32_t *__restrict__ kept,
                                 uint8_t *__restrict__ read_dead) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int b0 = group_off[g], b1 = group_off[g + 1];
    int top = b0;                        // kept[b0 .. top) is the stack
    int range_end = 0;
    for (int m = b0; m < b1; m++) {
        if (!ncalls[m]) continue;        // alignments without calls never reach addEdge
        const int first = first_pos[m], last = last_pos[m];
        bool drop_cur = false;
        while (0 <= first && first <= range_end) {
            if (last < range_end) { drop_cur = true; break; }
            if (top == b0) break;
            const int prev = kept[top - 1];
            const int prev_start = first_pos[prev], prev_end = last_pos[prev];
            const double ov_start = (double)max(prev_start, first), ov_end = (double)min(prev_end, last);
            if (ov_start > ov_end) break;
            const double ov_len = ov_end - ov_start + 1;
            const double span = (double)max(prev_end, last) - (double)min(prev_start, first) + 1;
            if (ov_len / span >= overlap_threshold) {
                const int len_prev = prev_end - prev_start + 1, len_cur = last - first + 1;
                if (len_cur <= len_prev) { drop_cur = true; break; }
                read_dead[members[prev]] = 1;
                top--;
                range_end = top == b0 ? first : last_pos[kept[top - 1]];
            } else break;
        }
        range_end = last;
        if (drop_cur) read_dead[members[m]] = 1;
        else kept[top++] = m;
    }
}
}  // namespace

// device half of the filters at the head of addEdge: d_read_dead for the batch.  No host synchronisation.
int lps_launch_overlap_filter(lps_ctx *ctx, const lps_phase_params *p) {
    cudaStream_t st = ctx->stream;
    const int n = ctx->batch.n_reads;
    const int m = (int)ctx->h_multi_members.size(), ng = (int)ctx->h_multi_group_off.size() - 1;
    LPS_CUDA(ctx, ctx->d_read_dead.reserve((size_t)n + 1));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_read_dead.p, 0, (size_t)n + 1, st));
    if (m <= 0 || ng <= 0) return LPS_OK;
    LPS_CUDA(ctx, ctx->d_first_pos.reserve((size_t)m + 1));
    LPS_CUDA(ctx, ctx->d_last_pos.reserve((size_t)m + 1));
    LPS_CUDA(ctx, ctx->d_multi_ncalls.reserve((size_t)m + 1));
    LPS_CUDA(ctx, ctx->d_multi_kept.reserve((size_t)m + 1));
    k_first_last<<<(m + 255) / 256, 256, 0, st>>>(m, ctx->d_multi_members.p, ctx->d_call_off.p, ctx->d_calls.p, ctx->var.pos,
                                                  ctx->d_first_pos.p, ctx->d_last_pos.p, ctx->d_multi_ncalls.p);
    k_overlap_filter<<<(ng + 127) / 128, 128, 0, st>>>(ng, ctx->d_multi_group_off.p, ctx->d_multi_members.p, ctx->d_first_pos.p, ctx->d_last_pos.p,
                                                       ctx->d_multi_ncalls.p, p->overlap_threshold, ctx->d_multi_kept.p, ctx->d_read_dead.p);
    ctx->stats.kernel_launches += 2;
    LPS_CUDA(ctx, cudaGetLastError());
    return LPS_OK;
}

// Graph construction + ordered fold.  The host knows upper bounds only (calls of the batch >= merged calls, variants >= nodes); the
// real counts stay in device memory (d_n_nodes, d_n_merged) and every kernel reads them there, so nothing here waits for the
// device - except, with sync_ties, the replay of tied merged reads on the host (std::sort of > 16 calls with equal positions).
// Without sync_ties the number of such groups is left in d_n_tie for the caller to check at the end of the contig.
int lps_launch_build_edges(lps_ctx *ctx, const lps_phase_params *p, bool sync_ties) {
    cudaStream_t st = ctx->stream;
    const int n = ctx->batch.n_reads, nv = ctx->var.n, W = p->connect_adjacent;
    const int tb = 256;
    if (W < 1 || W > 127) return ctx->fail(LPS_E_ARG, "connect_adjacent must be in [1,127]");   // last_link is an int8, the sweep packs 12-bit sums
    const size_t MU = (size_t)ctx->n_calls;      // upper bound of the merged calls
    const int NU = nv;                           // upper bound of the nodes
    const int RS = lps_vote_row_stride(W);
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
    LPS_CUDA(ctx, ctx->d_M.reserve(MU + 1));
    LPS_CUDA(ctx, ctx->d_M_gend.reserve(MU + 1));
    LPS_CUDA(ctx, ctx->d_M_node.reserve(MU + 1));
    LPS_CUDA(ctx, ctx->d_M_node_sorted.reserve(MU + 1));
    LPS_CUDA(ctx, ctx->d_M_idx.reserve(MU + 1));
    LPS_CUDA(ctx, ctx->d_M_idx_sorted.reserve(MU + 1));
    LPS_CUDA(ctx, ctx->d_M_unsorted.reserve(MU + 1));
    LPS_CUDA(ctx, ctx->d_node_cnt.reserve((size_t)NU + 2));
    LPS_CUDA(ctx, ctx->d_node_off.reserve((size_t)NU + 2));
    LPS_CUDA(ctx, ctx->d_weights.reserve((size_t)NU * (size_t)W * 4 + 4));
    LPS_CUDA(ctx, ctx->d_vote_info.reserve(((size_t)NU + 2) * (size_t)RS + 64));
    LPS_CUDA(ctx, ctx->d_last_link.reserve((size_t)NU + 16));
    LPS_CUDA(ctx, ctx->d_sweep_meta.reserve((size_t)NU + 16));
    LPS_CUDA(ctx, ctx->d_n_tie.reserve(1));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_var_lastw.p, 0, 8 * ((size_t)nv + 1), st));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_edge_counters.p, 0, 32, st));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_node_cnt.p, 0, 4 * ((size_t)NU + 2), st));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_n_tie.p, 0, 4, st));
    const uint8_t *erased = ctx->have_erased ? ctx->d_call_erased.p : nullptr;
    ctx->window = W;

    // 1. alive calls per read, node marking, alignment keys
    if (n > 0) {
        k_mark_nodes<<<(unsigned)(((long long)n * 16 + tb - 1) / tb), tb, 0, st>>>(
            n, ctx->d_call_off.p, ctx->d_calls.p, ctx->d_read_dead.p, erased, ctx->batch.name_rank,
            (unsigned long long *)ctx->d_var_lastw.p, ctx->d_alive_cnt.p, ctx->d_aln_keys.p);
        ctx->stats.kernel_launches++;
    }
    // 2. node numbering; the node count stays on the device behind the last element of the scan
    size_t cub_bytes = 0, need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, ctx->d_node_of_var.p, ctx->d_node_of_var.p, nv + 1, st);
    cub_bytes = need;
    // name ranks are < n: the key (rank << 32 | read) only has to be ordered by rank - the sort is stable and the keys arrive in read
    // order - so only the rank's bits are sorted
    const int rank_bits = bits_for((uint32_t)(n > 0 ? n : 1)) + 1;
    cub::DeviceRadixSort::SortKeys(nullptr, need, ctx->d_aln_keys.p, ctx->d_aln_keys_sorted.p, n, 32, std::min(64, 32 + rank_bits), st);
    if (need > cub_bytes) cub_bytes = need;
    cub::DeviceScan::ExclusiveSum(nullptr, need, ctx->d_grp_off.p, ctx->d_grp_off.p, n + 1, st);
    if (need > cub_bytes) cub_bytes = need;
    const int node_bits = bits_for((uint32_t)NU + 1);
    if (MU > 0) {
        cub::DeviceRadixSort::SortPairs(nullptr, need, ctx->d_M_node.p, ctx->d_M_node_sorted.p, ctx->d_M_idx.p, ctx->d_M_idx_sorted.p, (int)MU, 0, node_bits, st);
        if (need > cub_bytes) cub_bytes = need;
    }
    cub::DeviceScan::ExclusiveSum(nullptr, need, ctx->d_node_off.p, ctx->d_node_off.p, NU + 1, st);
    if (need > cub_bytes) cub_bytes = need;
    LPS_CUDA(ctx, ctx->d_cub_tmp.reserve(cub_bytes + 256));
    k_node_flags<<<(nv + tb) / tb, tb, 0, st>>>(nv, (const unsigned long long *)ctx->d_var_lastw.p, ctx->d_node_of_var.p);
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_node_of_var.p + nv, 0, 4, st));
    cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, cub_bytes, ctx->d_node_of_var.p, ctx->d_node_of_var.p, nv + 1, st);
    // the scan's last element IS the node count; copy it out of the array k_fill_nodes rewrites
    LPS_CUDA(ctx, ctx->d_n_nodes.reserve(2));
    LPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_n_nodes.p, ctx->d_node_of_var.p + nv, 4, cudaMemcpyDeviceToDevice, st));
    k_fill_nodes<<<(nv + tb) / tb, tb, 0, st>>>(nv, (const unsigned long long *)ctx->d_var_lastw.p, ctx->var.pos, ctx->d_node_of_var.p,
                                                ctx->d_node_var.p, ctx->d_node_pos.p, ctx->d_node_type.p);
    ctx->stats.kernel_launches += 3;
    // 3. alignments in (name rank, BAM order); offsets of their calls inside M.  Alignments without an alive call carry the key ~0: they
    //    sort to the end and own zero calls, the kernels below skip them.
    if (n > 0) cub::DeviceRadixSort::SortKeys(ctx->d_cub_tmp.p, cub_bytes, ctx->d_aln_keys.p, ctx->d_aln_keys_sorted.p, n, 32, std::min(64, 32 + rank_bits), st);
    k_sorted_counts<<<(n + 1 + tb) / tb, tb, 0, st>>>(n, ctx->d_aln_keys_sorted.p, ctx->d_alive_cnt.p, ctx->d_grp_off.p);
    cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, cub_bytes, ctx->d_grp_off.p, ctx->d_grp_off.p, n + 1, st);
    ctx->stats.kernel_launches += 3;
    const uint64_t *d_n_merged = ctx->d_grp_off.p + n;      // the scan's last element: merged calls
    if (n > 0 && MU > 0) {
        k_fill_merged<<<(unsigned)(((long long)n * 16 + tb - 1) / tb), tb, 0, st>>>(
            n, ctx->d_aln_keys_sorted.p, ctx->d_grp_off.p, ctx->d_call_off.p, ctx->d_calls.p, erased, ctx->d_node_of_var.p,
            p->base_quality, ctx->d_M.p, ctx->d_M_gend.p, ctx->d_node_cnt.p);
        // multi-alignment merged reads
        LPS_CUDA(ctx, ctx->d_tie_groups.reserve(std::max<size_t>(ctx->d_tie_groups.cap, 4096)));
        const int n_groups = (int)ctx->h_multi_group_off.size() - 1;
        if (n_groups > 0) {
            LPS_CUDA(ctx, ctx->d_pos_of_read.reserve((size_t)n + 1));
            k_sorted_position<<<(n + tb - 1) / tb, tb, 0, st>>>(n, ctx->d_aln_keys_sorted.p, ctx->d_pos_of_read.p);
            k_sort_multi_groups<<<(n_groups + SMG_WARPS - 1) / SMG_WARPS, SMG_WARPS * 32, 0, st>>>(
                n_groups, ctx->d_multi_group_off.p, ctx->d_multi_members.p, ctx->d_alive_cnt.p, ctx->d_pos_of_read.p, ctx->d_grp_off.p, ctx->d_M.p,
                ctx->d_M_unsorted.p, ctx->d_tie_groups.p, (uint32_t)std::min<size_t>(ctx->d_tie_groups.cap, 0xFFFFFFFFu), ctx->d_n_tie.p, 0);
            ctx->stats.kernel_launches += 2;
        }
        ctx->stats.kernel_launches += 1;
        if (sync_ties) {
            unsigned int h_tie = 0;
            LPS_CUDA(ctx, cudaMemcpyAsync(&h_tie, ctx->d_n_tie.p, 4, cudaMemcpyDeviceToHost, st));
            LPS_CUDA(ctx, cudaStreamSynchronize(st));
            if (h_tie > ctx->d_tie_groups.cap) {
                // more tied groups than the list holds: grow it and list them again (the merged reads are sorted by now)
                LPS_CUDA(ctx, ctx->d_tie_groups.reserve((size_t)h_tie + 1024));
                LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_n_tie.p, 0, 4, st));
                k_sort_multi_groups<<<(n_groups + SMG_WARPS - 1) / SMG_WARPS, SMG_WARPS * 32, 0, st>>>(
                    n_groups, ctx->d_multi_group_off.p, ctx->d_multi_members.p, ctx->d_alive_cnt.p, ctx->d_pos_of_read.p, ctx->d_grp_off.p, ctx->d_M.p,
                    ctx->d_M_unsorted.p, ctx->d_tie_groups.p, (uint32_t)std::min<size_t>(ctx->d_tie_groups.cap, 0xFFFFFFFFu), ctx->d_n_tie.p, 1);
                ctx->stats.kernel_launches++;
                LPS_CUDA(ctx, cudaMemcpyAsync(&h_tie, ctx->d_n_tie.p, 4, cudaMemcpyDeviceToHost, st));
                LPS_CUDA(ctx, cudaStreamSynchronize(st));
                if (h_tie > ctx->d_tie_groups.cap) return ctx->fail(LPS_E_NOMEM, "tie group list sizing did not converge");
            }
            if (h_tie) {
                int rc = lps_host_fix_tie_groups(ctx, (int)h_tie);
                if (rc) return rc;
                LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_n_tie.p, 0, 4, st));
            }
        }
        // 4. per-node call lists in merged (= name rank) order: one stable radix sort by node
        k_split_merged<<<(unsigned)((MU + tb - 1) / tb), tb, 0, st>>>((uint64_t)MU, d_n_merged, (uint32_t)NU, ctx->d_M.p, ctx->d_M_node.p, ctx->d_M_idx.p);
        cub::DeviceRadixSort::SortPairs(ctx->d_cub_tmp.p, cub_bytes, ctx->d_M_node.p, ctx->d_M_node_sorted.p, ctx->d_M_idx.p, ctx->d_M_idx_sorted.p, (int)MU, 0,
                                        node_bits, st);
        ctx->stats.kernel_launches += 2;
    }
    k_widen_u32b<<<(NU + 1 + tb) / tb, tb, 0, st>>>(NU, ctx->d_node_cnt.p, ctx->d_node_off.p);
    cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, cub_bytes, ctx->d_node_off.p, ctx->d_node_off.p, NU + 1, st);
    ctx->stats.kernel_launches += 2;
    // 5. the fold (nodes without calls in M - there are none - would simply write an empty row)
    if (NU > 0) {
        constexpr int WARPS = 8;
        cudaEventRecord(ctx->kev[2], st);
        static const bool wide_fold = getenv("LPS_FOLD_WIDE") != nullptr;      // A/B switch: the 32-lanes-per-node kernel
        if (W <= 48 && !wide_fold) {
            const size_t smem = (size_t)WARPS * 2 * W * 4 * sizeof(float);
            const int per_cta = WARPS * 2;
            k_fold_edges16<WARPS><<<(NU + per_cta - 1) / per_cta, WARPS * 32, smem, st>>>(
                ctx->d_n_nodes.p, W, p->edge_weight, ctx->d_node_off.p, ctx->d_M_idx_sorted.p, ctx->d_M.p, ctx->d_M_gend.p, ctx->d_weights.p,
                p->edge_threshold, ctx->d_vote_info.p, (unsigned long long *)ctx->d_edge_counters.p, d_n_merged, ctx->d_last_link.p, RS, ctx->d_node_pos.p,
                ctx->d_node_type.p, p->distance, ctx->d_sweep_meta.p);
        } else {
            const size_t smem = (size_t)WARPS * W * 4 * sizeof(float);
            k_fold_edges<WARPS><<<(NU + WARPS - 1) / WARPS, WARPS * 32, smem, st>>>(
                ctx->d_n_nodes.p, W, p->edge_weight, ctx->d_node_off.p, ctx->d_M_idx_sorted.p, ctx->d_M.p, ctx->d_M_gend.p, ctx->d_weights.p,
                p->edge_threshold, ctx->d_vote_info.p, (unsigned long long *)ctx->d_edge_counters.p, d_n_merged, ctx->d_last_link.p, RS, ctx->d_node_pos.p,
                ctx->d_node_type.p, p->distance, ctx->d_sweep_meta.p);
        }
        cudaEventRecord(ctx->kev[3], st);
        ctx->stats.kernel_launches++;
    }
    LPS_CUDA(ctx, cudaGetLastError());
    ctx->have_graph = true;
    return LPS_OK;
}

// counts of the graph (nodes, merged calls, pair contributions) to the host; waits for the stream
int lps_fetch_graph_counts(lps_ctx *ctx) {
    cudaStream_t st = ctx->stream;
    const int n = ctx->batch.n_reads;
    int32_t n_nodes = 0;
    uint64_t n_merged = 0;
    unsigned long long hc[2] = {0, 0};
    LPS_CUDA(ctx, cudaMemcpyAsync(&n_nodes, ctx->d_n_nodes.p, 4, cudaMemcpyDeviceToHost, st));
    LPS_CUDA(ctx, cudaMemcpyAsync(&n_merged, ctx->d_grp_off.p + n, 8, cudaMemcpyDeviceToHost, st));
    LPS_CUDA(ctx, cudaMemcpyAsync(hc, ctx->d_edge_counters.p, 16, cudaMemcpyDeviceToHost, st));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    if (ctx->var.n > 0) cudaEventElapsedTime(&ctx->stats.ms_kernel_fold_edges, ctx->kev[2], ctx->kev[3]);
    ctx->n_nodes = n_nodes; ctx->n_merged = n_merged;
    ctx->n_contrib = hc[0]; ctx->n_contrib_far = hc[1];
    return LPS_OK;
}
