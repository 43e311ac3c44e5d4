// k_annotate.cu — per-variant notes computed on the device from the contig reference.
//
//   homopolymer  = homopolymerLength(pos)                 reference src/shared/Util.cpp:21-54
//   is_danger    = indel followed by a 2-mer repeated x5   reference src/phase/ParsingBam.cpp:378-417
//   filtered     = right member of a homopolymer SNP pair  reference src/phase/ParsingBam.cpp:866-888
//
// One thread per variant; the filterSNP chain (left member survives and meets the next variant) only
// runs inside clusters of variants <= 2 bp apart, so the thread of each cluster head walks its cluster.
#include "lps_ctx.cuh"

namespace {

__device__ __forceinline__ char ref_at(const char *__restrict__ ref, int64_t len, int64_t i) {
    return (i >= 0 && i < len) ? ref[i] : '\0';
}

__global__ void k_variant_notes(const char *__restrict__ ref, int64_t ref_len, int n, const int32_t *__restrict__ pos,
                                const uint16_t *__restrict__ ref_len_v, const uint16_t *__restrict__ alt_len_v,
                                uint8_t *__restrict__ hom, uint8_t *__restrict__ danger) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t p = pos[i];
    int h = 1;
    if (p + 1 < ref_len) {
        char e = ref[p];
        int64_t q = p - 1;
        while (q >= 0 && ref[q] == e) {
            q--; h++;
            if (h >= 10 || q < 0) break;
        }
        q = p + 1;
        if (q < ref_len) {
            while (ref[q] == e) {
                q++; h++;
                if (q >= ref_len) break;
                if (h >= 10) break;
            }
        }
    }
    hom[i] = (uint8_t)h;
    int d = 0;
    if (ref_len_v[i] > 1 || alt_len_v[i] > 1) {
        char r0 = ref_at(ref, ref_len, p + 1), r1 = ref_at(ref, ref_len, p + 2);
        int k = 0;
        int64_t rp = p;
        while (k < 5) {
            if (r0 != ref_at(ref, ref_len, rp + 1) || r1 != ref_at(ref, ref_len, rp + 2)) break;
            rp += 2; k++;
        }
        d = (k == 5);
    }
    danger[i] = (uint8_t)d;
}

__global__ void k_filter_snp_chain(int n, const int32_t *__restrict__ pos, const uint8_t *__restrict__ hom,
                                   uint8_t *__restrict__ filtered) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // cluster head: no left neighbour within 2 bp (it can never be erased, so it becomes `cur`)
    if (i > 0 && pos[i] - pos[i - 1] <= 2) return;
    int cur = i;
    for (int nxt = i + 1; nxt < n && pos[nxt] - pos[nxt - 1] <= 2; nxt++) {
        if (hom[cur] >= 3 && hom[nxt] >= 3 && pos[nxt] - pos[cur] <= 2) filtered[nxt] = 1;
        else cur = nxt;
    }
}

// the 8-byte record the walking kernel reads per variant (DevVariants::vrec)
__global__ void k_pack_vrec(int n, const int32_t *__restrict__ pos, const uint8_t *__restrict__ ref0, const uint8_t *__restrict__ alt0,
                            const uint16_t *__restrict__ ref_len, const uint16_t *__restrict__ alt_len, const uint8_t *__restrict__ hom,
                            const uint8_t *__restrict__ danger, const uint8_t *__restrict__ filtered, const uint8_t *__restrict__ hp1_is_alt,
                            uint2 *__restrict__ vrec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned fl = (ref_len[i] == 1 ? 1u : 0u) | (alt_len[i] == 1 ? 2u : 0u) | (danger[i] ? 4u : 0u) | (filtered[i] ? 8u : 0u);
    if (hp1_is_alt && hp1_is_alt[i]) fl |= 16u;
    vrec[i] = make_uint2((unsigned)pos[i], (unsigned)ref0[i] | ((unsigned)alt0[i] << 8) | (fl << 16) | ((unsigned)hom[i] << 24));
}

}  // namespace

int lps_launch_annotate(lps_ctx *ctx) {
    int n = ctx->var.n;
    LPS_CUDA(ctx, ctx->d_vrec.reserve((size_t)n + 1));
    ctx->var.vrec = ctx->d_vrec.p;
    if (n == 0) return LPS_OK;
    LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_vfiltered.p, 0, (size_t)n, ctx->stream));
    int tb = 256, gb = (n + tb - 1) / tb;
    k_variant_notes<<<gb, tb, 0, ctx->stream>>>(ctx->d_ref.p, ctx->ref_len, n, ctx->var.pos, ctx->var.ref_len,
                                               ctx->var.alt_len, ctx->d_vhom.p, ctx->d_vdanger.p);
    ctx->stats.kernel_launches++;
    if (ctx->is_ont) {
        k_filter_snp_chain<<<gb, tb, 0, ctx->stream>>>(n, ctx->var.pos, ctx->d_vhom.p, ctx->d_vfiltered.p);
        ctx->stats.kernel_launches++;
    }
    k_pack_vrec<<<gb, tb, 0, ctx->stream>>>(n, ctx->var.pos, ctx->var.ref0, ctx->var.alt0, ctx->var.ref_len, ctx->var.alt_len, ctx->d_vhom.p,
                                           ctx->d_vdanger.p, ctx->d_vfiltered.p, ctx->have_tag_variants ? ctx->d_vhp1_is_alt.p : nullptr,
                                           ctx->d_vrec.p);
    ctx->stats.kernel_launches++;
    LPS_CUDA(ctx, cudaGetLastError());
    return LPS_OK;
}
