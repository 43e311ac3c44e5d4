// k_read_correction.cu — KERNEL 3a: VairiantGraph::readCorrection (reference
// src/phase/PhasingGraph.cpp:891-1029) and the per-variant part of exportResult (:1049-1077).
//
//   pass 1  one thread per surviving alignment: haplotype vote over its calls that sit in a phased
//           block (double adds of 1 / 0.1 IN CALL ORDER — the order is part of the result), read
//           haplotype if max/(sum) > readConfidence && sum > 1, then integer atomics
//           count[hp][variant][allele]++ for every call of a tagged read;
//   pass 2  one thread per variant: r1 = c[0][ref]+c[1][alt], r2 = c[1][ref]+c[0][alt]; keep the
//           orientation with confidence > snpConfidence, otherwise the variant is un-phased.
// Integer atomics only, so the result does not depend on scheduling.
#include "lps_ctx.cuh"

namespace {

// (r02: eight lanes per alignment with the votes added in order through shuffles measured 44 - 72 us against 40 us for this
// one-thread-per-alignment form: the kernel is bound by the integer atomics on the counters of neighbouring alignments' shared variants)
__global__ void k_read_vote(int n_reads, const uint64_t *__restrict__ call_off, const lps_call *__restrict__ calls,
                            const uint8_t *__restrict__ read_dead, const uint8_t *__restrict__ call_erased,
                            const unsigned long long *__restrict__ var_lastw, const int32_t *__restrict__ ps_sweep,
                            const int8_t *__restrict__ hap_sweep, double read_confidence, int8_t *__restrict__ read_hp,
                            int32_t *__restrict__ hp_counts) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    uint64_t c0 = call_off[r], c1 = call_off[r + 1];
    if (c0 == c1 || read_dead[r]) { read_hp[r] = -2; return; }
    double rc = 0.0, ac = 0.0;
    for (uint64_t c = c0; c < c1; c++) {
        if (call_erased && call_erased[c]) continue;
        lps_call cl = calls[c];
        if (ps_sweep[cl.var] == 0) continue;                                 // not in bkResult
        int h = cl.allele == 0 ? hap_sweep[cl.var] : 1 - hap_sweep[cl.var];   // subNodeHP[(pos, allele+1)]
        unsigned type = (unsigned)(var_lastw[cl.var] & 7ull);
        if (type == 0) { if (h == 0) rc++; else ac++; }                       // SNP: +1
        else { if (h == 0) rc += 0.1; else ac += 0.1; }                       // indel / danger indel: +0.1
    }
    double mx = rc > ac ? rc : ac;
    if (mx / (rc + ac) > read_confidence && (rc + ac) > 1) {
        int bh = rc > ac ? 0 : 1;
        read_hp[r] = (int8_t)bh;
        for (uint64_t c = c0; c < c1; c++) {
            if (call_erased && call_erased[c]) continue;
            lps_call cl = calls[c];
            atomicAdd(&hp_counts[(size_t)cl.var * 4 + (size_t)bh * 2 + (size_t)cl.allele], 1);
        }
    } else read_hp[r] = -1;
}

__global__ void k_variant_decide(int nv, const unsigned long long *__restrict__ var_lastw, const int32_t *__restrict__ hp_counts,
                                 const int32_t *__restrict__ ps_sweep, double snp_confidence, int32_t *__restrict__ ps,
                                 int8_t *__restrict__ hap_ref) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    int h = -1;
    if (var_lastw[i] != 0) {
        const int32_t *c = hp_counts + (size_t)i * 4;
        double r1 = (double)c[0] + (double)c[3], r2 = (double)c[2] + (double)c[1];
        double conf = (r1 > r2 ? r1 : r2) / (r1 + r2);                        // 0/0 = NaN -> un-phased
        if (conf > snp_confidence) { if (r1 > r2) h = 0; else if (r1 < r2) h = 1; }
    }
    hap_ref[i] = (int8_t)h;
    ps[i] = h >= 0 ? ps_sweep[i] : 0;
}

}  // namespace

// d_ps / d_hap_ref hold the sweep result per VARIANT on entry and the final result on exit
int lps_launch_read_correction(lps_ctx *ctx, const lps_phase_params *p) {
    cudaStream_t st = ctx->stream;
    const int n = ctx->batch.n_reads, nv = ctx->var.n;
    const int tb = 256;
    LPS_CUDA(ctx, ctx->d_read_hp.reserve((size_t)n + 1));
    LPS_CUDA(ctx, ctx->d_hp_counts.reserve((size_t)nv * 4 + 4));
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_hp_counts.p, 0, 16 * (size_t)nv + 16, st));
    const uint8_t *erased = ctx->have_erased ? ctx->d_call_erased.p : nullptr;
    DevBuf<int32_t> &ps_final = ctx->d_ps_final;
    DevBuf<int8_t> &hap_final = ctx->d_hap_final;
    LPS_CUDA(ctx, ps_final.reserve((size_t)nv + 1));
    LPS_CUDA(ctx, hap_final.reserve((size_t)nv + 1));
    if (n > 0) {
        k_read_vote<<<(n + tb - 1) / tb, tb, 0, st>>>(n, ctx->d_call_off.p, ctx->d_calls.p, ctx->d_read_dead.p, erased,
                                                      (const unsigned long long *)ctx->d_var_lastw.p, ctx->d_ps.p, ctx->d_hap_ref.p,
                                                      p->read_confidence, ctx->d_read_hp.p, ctx->d_hp_counts.p);
        ctx->stats.kernel_launches++;
    }
    if (nv > 0) {
        k_variant_decide<<<(nv + tb - 1) / tb, tb, 0, st>>>(nv, (const unsigned long long *)ctx->d_var_lastw.p, ctx->d_hp_counts.p,
                                                            ctx->d_ps.p, p->snp_confidence, ps_final.p, hap_final.p);
        ctx->stats.kernel_launches++;
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ps.p, ps_final.p, 4 * (size_t)nv, cudaMemcpyDeviceToDevice, st));
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_hap_ref.p, hap_final.p, (size_t)nv, cudaMemcpyDeviceToDevice, st));
    }
    LPS_CUDA(ctx, cudaGetLastError());
    return LPS_OK;          // asynchronous: the caller waits for the stream when it needs the result
}
