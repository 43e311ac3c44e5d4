// host_phase.cpp — host-side steps of the `phase` pipeline that sit BETWEEN the kernels.
//
// They are inherently sequential or tiny (O(#alignments sharing a name), O(#clip positions),
// O(#variants * window)) and stay on the host by design, but they
// must be bit-exact because kernel 2 consumes the filters and kernel 3a consumes the sweep:
//   * overlap filter among the alignments of one read name   reference PhasingGraph.cpp:707-781
//   * Clip::getCNVInterval (run twice) + CNV mismatch filter  reference PhasingGraph.cpp:520-692, 1103-1227
//   * edgeConnectResult / findBestEdgePair / Onelongcase       reference PhasingGraph.cpp:166-228, 251-283, 286-474
//   * std::sort replay for merged reads with tied positions    reference Util.cpp:3-5
#include <algorithm>
#include <cmath>
#include <map>
#include "lps_ctx.cuh"

// --------------------------------------------------------------------------------------------
// overlap filter.  Works on alignments that produced >= 1 call (after filterSNP); only names that
// own more than one such alignment can be affected, so the state machine runs on those alone.
// Sets ctx->h_read_dead.
// --------------------------------------------------------------------------------------------
int lps_host_overlap_filter(lps_ctx *ctx, const lps_phase_params *p, const std::vector<int32_t> &first_pos,
                            const std::vector<int32_t> &last_pos, const std::vector<uint32_t> &ncalls) {
    const int n = (int)ncalls.size();
    ctx->h_read_dead.assign((size_t)n, 0);
    const std::vector<int32_t> &rank = ctx->h_name_rank;
    int max_rank = -1;
    for (int r = 0; r < n; r++) if (ncalls[r] && rank[r] > max_rank) max_rank = rank[r];
    if (max_rank < 0) return LPS_OK;
    // bucket the alignments with calls by name rank, BAM order preserved inside a bucket
    std::vector<int32_t> bucket_off((size_t)max_rank + 2, 0);
    for (int r = 0; r < n; r++) if (ncalls[r]) bucket_off[(size_t)rank[r] + 1]++;
    for (int k = 0; k <= max_rank; k++) bucket_off[(size_t)k + 1] += bucket_off[(size_t)k];
    std::vector<int32_t> members((size_t)bucket_off[(size_t)max_rank + 1]);
    {
        std::vector<int32_t> cursor(bucket_off.begin(), bucket_off.end() - 1);
        for (int r = 0; r < n; r++) if (ncalls[r]) members[(size_t)cursor[(size_t)rank[r]]++] = r;
    }
    std::vector<int32_t> kept;   // readIdxVec[name]
    for (int k = 0; k <= max_rank; k++) {
        const int b0 = bucket_off[(size_t)k], b1 = bucket_off[(size_t)k + 1];
        if (b1 - b0 < 2) continue;
        kept.clear();
        // alignRange[name]: inserted as {0,0} before the find(), so .first is always 0 (:712-716)
        int range_end = 0;
        for (int m = b0; m < b1; m++) {
            const int cur = members[(size_t)m];
            const int first = first_pos[(size_t)cur], last = last_pos[(size_t)cur];
            bool drop_cur = false;
            while (0 <= first && first <= range_end) {
                if (last < range_end) { drop_cur = true; break; }
                if (kept.empty()) break;
                const int prev = kept.back();
                const int prev_start = first_pos[(size_t)prev], prev_end = last_pos[(size_t)prev];
                const double ov_start = std::max(prev_start, first), ov_end = std::min(prev_end, last);
                if (ov_start > ov_end) break;
                const double ov_len = ov_end - ov_start + 1;
                const double span = (double)std::max(prev_end, last) - (double)std::min(prev_start, first) + 1;
                if (ov_len / span >= p->overlap_threshold) {
                    const int len_prev = prev_end - prev_start + 1, len_cur = last - first + 1;
                    if (len_cur <= len_prev) { drop_cur = true; break; }
                    ctx->h_read_dead[(size_t)prev] = 1;
                    kept.pop_back();
                    range_end = kept.empty() ? first : last_pos[(size_t)kept.back()];
                } else break;
            }
            range_end = last;
            if (drop_cur) ctx->h_read_dead[(size_t)cur] = 1;
            else kept.push_back(cur);
        }
    }
    return LPS_OK;
}

// --------------------------------------------------------------------------------------------
// Clip::getCNVInterval state machine over the clipCount map (ascending positions).
// --------------------------------------------------------------------------------------------
namespace {
struct CnvState {
    bool push = false, slow_up = false, slow_down = false;
    int curr = 0, reject = 0, pull_down = 0, slow_down_count = 0, cand_start = -1, cand_end = -1;
    void reset() { *this = CnvState(); }
    void thresholds(int up) {   // Clip::updateThreshold :1112-1126
        reject = up;
        if (up >= 20) { pull_down = up / 2; slow_down_count = 5; }
        else if (up >= 10) { pull_down = up / 2; slow_down_count = up / 4; }
        else { pull_down = 5; slow_down_count = 2; }
    }
    void arm(int pos, int up, int down, int area) {
        push = true; slow_up = false; slow_down = true;
        curr = up - down; cand_start = pos; cand_end = pos + area;
        thresholds(up);
    }
};
}  // namespace

void lps_host_cnv_intervals(const std::vector<int32_t> &pos, const std::vector<int32_t> &front,
                            const std::vector<int32_t> &back, std::vector<int32_t> &cs, std::vector<int32_t> &ce) {
    const int area = 30000;
    const size_t n = pos.size();
    if (n == 0) return;   // the reference crashes here (rbegin() of an empty map); we return no interval
    CnvState s;
    for (size_t k = 0; k <= n; k++) {
        const bool sentinel = k == n;   // copy of the last entry, AreaSize to the right (:1134)
        const int at = sentinel ? pos[n - 1] + area : pos[k];
        const int up = front[sentinel ? n - 1 : k], down = back[sentinel ? n - 1 : k];
        const bool idle = !s.push && !s.slow_down && !s.slow_up;
        if (idle) {
            if (up >= 5 && s.curr == 0) s.arm(at, up, down, area);
            else if (up > down && s.curr == 0) {
                s.push = false; s.slow_up = true; s.slow_down = false;
                s.curr = up - down; s.cand_start = at; s.cand_end = at + area;
            }
        } else if (s.push && s.slow_down) {
            if (up > s.reject) {
                s.thresholds(up);
                s.cand_start = at; s.cand_end = at + area;
            }
            s.curr += up - down;
            if (s.curr > 30) s.cand_end = at + area;
            if (down >= s.pull_down || (s.curr <= s.slow_down_count && at <= s.cand_end)) {
                cs.push_back(s.cand_start); ce.push_back(at);
                s.reset();
            }
            if (at > s.cand_end || s.curr <= 0 || at - s.cand_start >= 200000) s.reset();
        } else if (s.slow_up) {
            const bool closes = s.curr > 20 ? down >= s.curr / 4 : down >= 5;
            if (closes) { cs.push_back(s.cand_start); ce.push_back(at); s.reset(); }
            else if (up >= 5) s.arm(at, up, down, area);
            else {
                s.curr += up - down;
                if (s.curr > 30) s.cand_end = at + area;
                if (at > s.cand_end || s.curr <= 0 || at - s.cand_start >= 200000) s.reset();
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// CNV mismatch filter (rare: needs clip pile-ups).  Operates on host copies of the calls of the
// surviving alignments; fills `erased` (one byte per call of the batch CSR).
// The interval vector is the state machine's output TWICE (Clip ctor + PhasingProcess.cpp:148), hence
// not sorted; the index walking below follows the reference literally for that reason.
// --------------------------------------------------------------------------------------------
int lps_host_cnv_filter(lps_ctx *ctx, std::vector<uint8_t> &erased) {
    const std::vector<int32_t> &cs = ctx->h_cnv_start, &ce = ctx->h_cnv_end;
    const size_t ncnv = cs.size();
    const int n = (int)ctx->h_status.size();
    erased.assign(ctx->h_calls.size(), 0);
    if (ncnv == 0) return LPS_OK;
    const std::vector<uint64_t> &off = ctx->h_call_off;
    const std::vector<lps_call> &calls = ctx->h_calls;
    const std::vector<int32_t> &vpos = ctx->h_vpos;
    auto inside = [](int p, int s, int e) { return p >= s && p <= e; };
    std::vector<int> alns;
    for (int r = 0; r < n; r++) if (off[(size_t)r + 1] > off[(size_t)r] && !ctx->h_read_dead[(size_t)r]) alns.push_back(r);

    // calculateCnvMismatchRate: ALT calls of each alignment inside each interval, keyed by interval start
    std::vector<std::map<int, int>> alt_in_cnv(alns.size());
    size_t ci = 0;
    for (size_t a = 0; a < alns.size(); a++) {
        const uint64_t c0 = off[(size_t)alns[a]], c1 = off[(size_t)alns[a] + 1];
        const int rs = vpos[(size_t)calls[c0].var], re = vpos[(size_t)calls[c1 - 1].var];
        while (ci > 0 && cs[ci] > rs) ci--;
        size_t i = ci;
        for (; i < ncnv && cs[i] <= re; i++)
            for (uint64_t c = c0; c < c1; c++) {
                const int vp = vpos[(size_t)calls[c].var];
                if (vp > ce[i]) break;
                if (inside(vp, cs[i], ce[i]) && calls[c].allele == 1) alt_in_cnv[a][cs[i]]++;
            }
        ci = i > 0 ? i - 1 : 0;
    }
    // aggregateCnvReadMismatchRate: per (variant, allele) the list of those per-read counts
    std::map<int, std::map<int, std::vector<int>>> per_variant;   // variant index -> allele -> counts
    ci = 0;
    for (size_t a = 0; a < alns.size(); a++) {
        const uint64_t c0 = off[(size_t)alns[a]], c1 = off[(size_t)alns[a] + 1];
        const int rs = vpos[(size_t)calls[c0].var], re = vpos[(size_t)calls[c1 - 1].var];
        while (ci > 0 && cs[ci] > rs) ci--;
        size_t i = ci;
        for (; i < ncnv && cs[i] <= re; i++)
            for (uint64_t c = c0; c < c1; c++) {
                const int vp = vpos[(size_t)calls[c].var];
                if (vp > ce[i]) break;
                auto hit = alt_in_cnv[a].find(cs[i]);
                if (inside(vp, cs[i], ce[i]) && hit != alt_in_cnv[a].end()) per_variant[calls[c].var][calls[c].allele].push_back(hit->second);
            }
        ci = i > 0 ? i - 1 : 0;
    }
    // calculateAverageMismatchRate
    auto mean = [](const std::vector<int> &v) { double s = 0.0; for (int x : v) s += x; return v.empty() ? 0.0 : s / (double)v.size(); };
    std::map<int, double> miss;
    for (const auto &kv : per_variant) {
        const int vp = vpos[(size_t)kv.first];
        for (size_t i = 0; i < ncnv; i++) {   // the reference never advances its start index in this function
            if (cs[i] > vp) break;
            if (!inside(vp, cs[i], ce[i])) continue;
            auto r = kv.second.find(0), al = kv.second.find(1);
            if (r == kv.second.end() || al == kv.second.end()) continue;
            const double mr = mean(r->second), ma = mean(al->second);
            if (mr != 0 && ma != 0) miss[kv.first] = ma / (mr + ma);
        }
    }
    if (miss.empty()) return LPS_OK;
    // filterHighMismatchVariants
    ci = 0;
    for (size_t a = 0; a < alns.size(); a++) {
        const uint64_t c0 = off[(size_t)alns[a]], c1 = off[(size_t)alns[a] + 1];
        const int rs = vpos[(size_t)calls[c0].var];
        while (ci > 0 && cs[ci] > rs) ci--;
        for (uint64_t c = c0; c < c1; c++) {
            const int vp = vpos[(size_t)calls[c].var];
            size_t i = ci;
            for (; i < ncnv && cs[i] <= vp; i++) {
                if (!inside(vp, cs[i], ce[i])) continue;
                auto m = miss.find(calls[c].var);
                if (m != miss.end() && m->second >= 0.7) { erased[c] = 1; break; }
            }
            ci = i > 0 ? i - 1 : 0;
        }
    }
    return LPS_OK;
}

// --------------------------------------------------------------------------------------------
// the sweep: every node receives weighted votes from up to `window` predecessors and, in turn, votes
// on its `window` successors.  Votes for node t can only come from nodes t-window..t-1, so a ring of
// window+1 slots holds all live vote lists.
// --------------------------------------------------------------------------------------------
namespace {
struct Vote { int voter; float para, cross, weight; int hap; double esr; };
}

void lps_host_sweep(const lps_phase_params *p, int32_t N, int32_t W, const int32_t *node_pos, const uint8_t *node_type,
                    const float *weights, int32_t *node_ps, int8_t *node_hap_ref) {
    for (int k = 0; k < N; k++) { node_ps[k] = 0; node_hap_ref[k] = -1; }
    if (N < 2) return;
    const int R = W + 1;
    std::vector<std::vector<Vote>> ring((size_t)R);
    std::vector<float> w1((size_t)R, 0.0f), w2((size_t)R, 0.0f);   // hpCountMap2[node][1|2]
    std::vector<int8_t> hp((size_t)N, 0);                           // hpResult
    std::vector<int32_t> block_of((size_t)N, -2);                   // -2: never entered a block
    int block_start = -1, last_connect = -1;
    for (int k = 0; k + 1 < N; k++) {
        const int slot = k % R;
        std::vector<Vote> &my_votes = ring[(size_t)slot];
        float h1 = w1[(size_t)slot], h2 = w2[(size_t)slot];
        const bool skip = std::abs(node_pos[k + 1] - node_pos[k]) > p->distance;   // :318-320
        bool enter = !skip;
        if (enter) {
            // Onelongcase (:251-283): many single-read votes -> trust only consistent multi-read non-indel ones
            int singles = 0;
            float s1 = 0, s2 = 0;
            for (const Vote &v : my_votes) {
                if ((v.para + v.cross) <= 1) singles++;
                else if (v.esr < 0.2 && v.weight >= 1 && node_type[v.voter] != 3) {
                    if (v.hap == 1) s1 += v.weight; else if (v.hap == 2) s2 += v.weight;
                }
            }
            if (!(singles <= 3 || (s1 == 0 && s2 == 0))) { h1 = s1; h2 = s2; }
            if (h1 == h2) {
                if (last_connect >= 0 && node_pos[k] < node_pos[last_connect]) enter = false;   // :340-342
                else { block_start = k; block_of[(size_t)k] = k; hp[(size_t)k] = 1; }
            } else {
                hp[(size_t)k] = h1 > h2 ? 1 : 2;
                block_of[(size_t)k] = block_start;
            }
        }
        if (enter) {
            const float *row = weights + (size_t)k * (size_t)W * 4;
            for (int d = 0; d < W && k + 1 + d < N; d++) {
                const int t = k + 1 + d;
                const float rr = row[d * 4 + 0], ra = row[d * 4 + 1], ar = row[d * 4 + 2], aa = row[d * 4 + 3];
                // findBestEdgePair (:166-228)
                const float para = rr + aa, cross = ar + ra;
                const double esr = (double)std::min(para, cross) / (double)std::max(para, cross);
                int link = 0;                            // 1: same haplotype, 2: opposite, 0: no connection
                if (rr + aa > ra + ar) link = 1; else if (rr + aa < ra + ar) link = 2;
                if (esr > p->edge_threshold) link = 0;
                Vote v;
                v.voter = k; v.weight = 1; v.hap = 0;
                if ((esr <= 0.1 && (rr + aa + ra + ar) >= 1) || ((rr + aa) < 1 && (ra + ar) >= 1) || ((rr + aa) >= 1 && (ra + ar) < 1))
                    v.weight = 20;
                v.para = rr + aa; v.cross = ra + ar; v.esr = esr;
                if (node_type[k] == 4) v.weight = (float)0.1;                   // danger indel voter (:367-369)
                if (link) {
                    const bool to_h1 = (hp[(size_t)k] == 1) == (link == 1);
                    const int ts = t % R;
                    if (to_h1) { w1[(size_t)ts] += v.weight; v.hap = 1; } else { w2[(size_t)ts] += v.weight; v.hap = 2; }
                    ring[(size_t)ts].push_back(v);
                    last_connect = t;
                }
            }
        }
        // slot k is recycled for node k + R
        my_votes.clear(); w1[(size_t)slot] = 0.0f; w2[(size_t)slot] = 0.0f;
    }
    // blocks -> PS and haplotype of the REF allele (:423-467); one-node blocks are dropped
    int prev = -1, prev_block = -3;
    for (int k = 0; k < N; k++) {
        const int b = block_of[(size_t)k];
        if (b == -2) continue;
        if (b != prev_block) { prev_block = b; prev = k; continue; }
        const int ps = node_pos[b] + 1;
        if (node_ps[prev] == 0) { node_ps[prev] = ps; if (node_hap_ref[prev] < 0) node_hap_ref[prev] = 0; }
        node_ps[k] = ps;
        node_hap_ref[k] = (int8_t)(hp[(size_t)prev] == hp[(size_t)k] ? node_hap_ref[prev] : 1 - node_hap_ref[prev]);
        prev = k;
    }
}

// --------------------------------------------------------------------------------------------
// merged reads with tied positions and > 16 calls: replay std::sort like ReadVariant::sort()
// --------------------------------------------------------------------------------------------
namespace {
struct SortRec { int position; uint32_t packed; };
struct ByPosition { bool operator()(const SortRec &a, const SortRec &b) const { return a.position < b.position; } };
}

int lps_host_fix_tie_groups(lps_ctx *ctx, const std::vector<uint32_t> &heads, const std::vector<uint64_t> &keys_sorted,
                            const std::vector<uint64_t> &grp_off, int base_quality) {
    const int n = ctx->batch.n_reads, nv = ctx->var.n;
    // host copies of what the groups are made of
    std::vector<uint64_t> off((size_t)n + 1);
    LPS_CUDA(ctx, cudaMemcpy(off.data(), ctx->d_call_off.p, 8 * ((size_t)n + 1), cudaMemcpyDeviceToHost));
    std::vector<int32_t> node_of((size_t)nv);
    LPS_CUDA(ctx, cudaMemcpy(node_of.data(), ctx->d_node_of_var.p, 4 * (size_t)nv, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> erased;
    if (ctx->have_erased) {
        erased.resize((size_t)ctx->n_calls);
        LPS_CUDA(ctx, cudaMemcpy(erased.data(), ctx->d_call_erased.p, (size_t)ctx->n_calls, cudaMemcpyDeviceToHost));
    }
    const int n_aln = (int)keys_sorted.size();
    std::vector<lps_call> buf;
    std::vector<SortRec> recs;
    for (uint32_t head : heads) {
        const uint32_t rank = (uint32_t)(keys_sorted[head] >> 32);
        int j = (int)head;
        recs.clear();
        for (; j < n_aln && (uint32_t)(keys_sorted[(size_t)j] >> 32) == rank; j++) {
            const int r = (int)(uint32_t)keys_sorted[(size_t)j];
            const uint64_t c0 = off[(size_t)r], c1 = off[(size_t)r + 1];
            buf.resize((size_t)(c1 - c0));
            if (c1 > c0) LPS_CUDA(ctx, cudaMemcpy(buf.data(), ctx->d_calls.p + c0, sizeof(lps_call) * (size_t)(c1 - c0), cudaMemcpyDeviceToHost));
            for (uint64_t c = c0; c < c1; c++) {
                if (!erased.empty() && erased[(size_t)c]) continue;
                const lps_call &cl = buf[(size_t)(c - c0)];
                const int q = cl.quality < 0 ? 60 : cl.quality;
                SortRec s;
                s.position = ctx->h_vpos[(size_t)cl.var];
                s.packed = ((uint32_t)node_of[(size_t)cl.var] << 2) | ((uint32_t)cl.allele << 1) | (q >= base_quality ? 1u : 0u);
                recs.push_back(s);
            }
        }
        std::sort(recs.begin(), recs.end(), ByPosition());
        const uint64_t g0 = grp_off[head], g1 = grp_off[(size_t)j];
        if (recs.size() != (size_t)(g1 - g0)) return ctx->fail(LPS_E_STATE, "merged group size mismatch in tie fix-up");
        std::vector<uint32_t> packed(recs.size());
        for (size_t i = 0; i < recs.size(); i++) packed[i] = recs[i].packed;
        LPS_CUDA(ctx, cudaMemcpy(ctx->d_M.p + g0, packed.data(), 4 * packed.size(), cudaMemcpyHostToDevice));
    }
    return LPS_OK;
}
