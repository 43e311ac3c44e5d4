// lps_api.cu — the extern "C" boundary declared in include/lps.h.
#include <chrono>
#include <climits>
#include <cmath>
#include <cub/cub.cuh>
#include "lps_ctx.cuh"

namespace {

template <typename T> int h2d(lps_ctx *ctx, DevBuf<T> &dst, const T *src, size_t n, size_t pad = 0) {
    LPS_CUDA(ctx, dst.reserve(n + pad + 1));
    if (n) LPS_CUDA(ctx, cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += n * sizeof(T);
    return LPS_OK;
}
template <typename T> int d2h(lps_ctx *ctx, std::vector<T> &dst, const T *src, size_t n) {
    dst.resize(n);
    if (n) LPS_CUDA(ctx, cudaMemcpyAsync(dst.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.d2h_bytes += n * sizeof(T);
    return LPS_OK;
}
#define TRY(x) do { int rc__ = (x); if (rc__ != LPS_OK) return rc__; } while (0)

// BAM's uint32 CIGAR ops -> the 16-bit stream the kernels read (len << 4 | op; a length >= 4095 becomes 0xFFF and goes to the side
// table).  Eight ops per thread: two 128-bit loads, one 128-bit store.  Escapes are appended as (op index << 28 | length) keys and
// sorted afterwards (they are rare: N ops, long matches of high-accuracy reads), which gives the table in ascending op order.
__global__ void k_narrow_cigar32(size_t n, const uint32_t *__restrict__ in, uint16_t *__restrict__ out, unsigned long long *__restrict__ long_keys,
                                 unsigned int long_cap, unsigned int *__restrict__ n_long) {
    const size_t i8 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i8 >= n) return;
    uint32_t w[8];
    if (i8 + 8 <= n) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(in + i8)), b = __ldg(reinterpret_cast<const uint4 *>(in + i8 + 4));
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
        for (int j = 0; j < 8; j++) w[j] = i8 + j < n ? in[i8 + j] : 1u;
    }
    uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t len = w[j] >> 4;
        if (len >= 0xFFFu) {
            h[j] = 0xFFF0u | (w[j] & 15u);
            const unsigned int k = atomicAdd(n_long, 1u);
            if (k < long_cap) long_keys[k] = ((unsigned long long)(i8 + j) << 28) | (unsigned long long)len;
        } else h[j] = w[j] & 0xFFFFu;
    }
    *reinterpret_cast<uint4 *>(out + i8) = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
}
// the 8-bit wire format -> the 16-bit stream: one warp per block of 256 ops (8 per lane: one 64-bit load, one 128-bit store); the
// k-th 0xFF byte of the stream takes esc16[k], k = escapes before the block (host table) + escapes before the op inside the block
__global__ void k_expand_cigar8(size_t n_ops, const uint8_t *__restrict__ in8, const uint16_t *__restrict__ esc16, const uint32_t *__restrict__ esc_blk,
                                uint16_t *__restrict__ out) {
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const size_t i8 = warp * 256 + (size_t)lane * 8;
    if (warp * 256 >= n_ops) return;
    unsigned long long w = 0x0101010101010101ull * 0u;
    if (i8 + 8 <= n_ops) w = *reinterpret_cast<const unsigned long long *>(in8 + i8);
    else for (int j = 0; j < 8; j++) if (i8 + j < n_ops) w |= (unsigned long long)in8[i8 + j] << (8 * j);
    int c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) c += (i8 + j < n_ops && ((w >> (8 * j)) & 0xFFull) == 0xFFull) ? 1 : 0;
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    size_t e = (size_t)esc_blk[warp] + (size_t)(inc - c);
    uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t b = (uint32_t)((w >> (8 * j)) & 0xFFull);
        uint32_t v;
        if (b < 0x80u) v = ((b + 1u) << 4) | 0u;                       // M
        else if (b < 0xB8u) v = ((b - 0x80u + 1u) << 4) | 1u;          // I
        else if (b < 0xF0u) v = ((b - 0xB8u + 1u) << 4) | 2u;          // D
        else v = (i8 + j < n_ops && b == 0xFFu) ? (uint32_t)esc16[e++] : 1u;
        h[j] = v;
    }
    if (i8 < n_ops) *reinterpret_cast<uint4 *>(out + i8) = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
}

__global__ void k_unpack_long_keys(unsigned int n, const unsigned long long *__restrict__ keys, uint64_t *__restrict__ at, uint32_t *__restrict__ len) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { at[i] = keys[i] >> 28; len[i] = (uint32_t)(keys[i] & 0xFFFFFFFull); }
}

int fetch_host_calls(lps_ctx *ctx) {
    if (ctx->host_calls_valid) return LPS_OK;
    const size_t n = (size_t)ctx->batch.n_reads;
    TRY(d2h(ctx, ctx->h_call_off, ctx->d_call_off.p, n + 1));
    TRY(d2h(ctx, ctx->h_calls, ctx->d_calls.p, (size_t)ctx->n_calls));
    TRY(d2h(ctx, ctx->h_status, ctx->d_status.p, n));
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->host_calls_valid = true;
    return LPS_OK;
}

struct WallTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    float ms() const { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

float elapsed(lps_ctx *ctx, int a, int b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]);
    return ms;
}

// a uint32 CIGAR stream in DEVICE memory -> ctx->d_cigar16 + the side table of escaped lengths
int narrow_cigar(lps_ctx *ctx, const uint32_t *d_cigar32, uint64_t n_ops) {
    cudaStream_t st = ctx->stream;
    LPS_CUDA(ctx, ctx->d_cigar16.reserve((size_t)n_ops + 64));
    LPS_CUDA(ctx, ctx->d_n_long.reserve(1));
    unsigned int n_long = 0;
    size_t cap = std::max<size_t>(ctx->d_long_keys.cap, 4096);
    for (int attempt = 0; attempt < 2; attempt++) {
        LPS_CUDA(ctx, ctx->d_long_keys.reserve(cap));
        LPS_CUDA(ctx, cudaMemsetAsync(ctx->d_n_long.p, 0, 4, st));
        if (n_ops) {
            const size_t threads = (n_ops + 7) / 8;
            k_narrow_cigar32<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((size_t)n_ops, d_cigar32, ctx->d_cigar16.p,
                                                                                (unsigned long long *)ctx->d_long_keys.p,
                                                                                (unsigned int)std::min<size_t>(ctx->d_long_keys.cap, 0xFFFFFFFFu), ctx->d_n_long.p);
            ctx->stats.kernel_launches++;
        }
        LPS_CUDA(ctx, cudaMemcpyAsync(&n_long, ctx->d_n_long.p, 4, cudaMemcpyDeviceToHost, st));
        LPS_CUDA(ctx, cudaStreamSynchronize(st));
        if (n_long <= ctx->d_long_keys.cap) break;
        cap = (size_t)n_long + 1024;
    }
    LPS_CUDA(ctx, ctx->d_cigar_long_at.reserve((size_t)n_long + 1));
    LPS_CUDA(ctx, ctx->d_cigar_long_len.reserve((size_t)n_long + 1));
    if (n_long) {
        DevBuf<uint64_t> &sorted = ctx->d_aln_keys_sorted;       // scratch of the graph stage, free at submit time
        LPS_CUDA(ctx, sorted.reserve((size_t)n_long + 1));
        size_t tmp = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tmp, ctx->d_long_keys.p, sorted.p, (int)n_long, 0, 64, st);
        LPS_CUDA(ctx, ctx->d_cub_tmp.reserve(tmp + 256));
        cub::DeviceRadixSort::SortKeys(ctx->d_cub_tmp.p, tmp, ctx->d_long_keys.p, sorted.p, (int)n_long, 0, 64, st);
        k_unpack_long_keys<<<(n_long + 255) / 256, 256, 0, st>>>(n_long, (const unsigned long long *)sorted.p, ctx->d_cigar_long_at.p, ctx->d_cigar_long_len.p);
        ctx->stats.kernel_launches += 2;
    }
    LPS_CUDA(ctx, cudaGetLastError());
    DevBatch &d = ctx->batch;
    d.cigar16 = ctx->d_cigar16.p; d.long_at = ctx->d_cigar_long_at.p; d.long_len = ctx->d_cigar_long_len.p; d.n_long = n_long;
    return LPS_OK;
}

// O(n) sanity of the offsets an integrator hands in: the kernels index with them directly
int validate_batch(lps_ctx *ctx, const lps_read_batch *b) {
    for (int32_t i = 0; i < b->n_reads; i++) {
        const int64_t lq = b->l_qseq[i];
        if (lq < 0) return ctx->fail(LPS_E_ARG, "negative l_qseq");
        if (b->cigar_off[i] > b->cigar_len || (uint64_t)b->n_cigar[i] > b->cigar_len - b->cigar_off[i])
            return ctx->fail(LPS_E_ARG, "cigar_off + n_cigar runs past cigar_len");
        if (b->sq) {
            if (lq > 0 && ((b->seq_off[i] & 15u) || b->seq_off[i] > b->sq_bytes || lps_sq_row_bytes((int32_t)lq) > b->sq_bytes - b->seq_off[i]))
                return ctx->fail(LPS_E_ARG, "seq_off is not a 16-byte aligned row of sq[] that ends within sq_bytes");
            continue;
        }
        if (lq > 0 && (b->seq_off[i] > b->seq_bytes || (uint64_t)(lq + 1) / 2 > b->seq_bytes - b->seq_off[i]))
            return ctx->fail(LPS_E_ARG, "seq_off + (l_qseq + 1) / 2 runs past seq_bytes");
        if (lq > 0 && (b->qual_off[i] > b->qual_bytes || (uint64_t)lq > b->qual_bytes - b->qual_off[i]))
            return ctx->fail(LPS_E_ARG, "qual_off + l_qseq runs past qual_bytes");
    }
    return LPS_OK;
}

}  // namespace

extern "C" {

#ifndef LPS_SOURCE_HASH
#define LPS_SOURCE_HASH "unknown"
#endif
// the hash covers every source of this library (csrc/Makefile), so a stale binary is recognisable
const char *lps_version(void) { return "longphase-s_b200 0.2 (sm_100a) src:" LPS_SOURCE_HASH; }

int lps_set_blocking_sync(int device, int on) {
    if (cudaSetDevice(device) != cudaSuccess) return LPS_E_CUDA;
    const cudaError_t e = cudaSetDeviceFlags(on == 1 ? cudaDeviceScheduleBlockingSync : on == 2 ? cudaDeviceScheduleYield : cudaDeviceScheduleSpin);
    cudaGetLastError();
    return e == cudaSuccess ? LPS_OK : LPS_E_CUDA;
}

int lps_ctx_create(int device, lps_ctx **out) {
    if (!out) return LPS_E_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) return LPS_E_CUDA;   // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return LPS_E_CUDA;
    {
        // the SEQ / QUAL gathers touch one byte per 32-byte sector: ask L2 not to fetch the neighbouring sector along with it
        const char *env = getenv("LPS_L2_FETCH");
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, env ? (size_t)atoi(env) : 32);
        cudaGetLastError();
    }
    lps_ctx *ctx = new lps_ctx();
    ctx->device = device;
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) ctx->sm_count = sms;
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LPS_E_CUDA; }
    for (auto &ev : ctx->ev) cudaEventCreate(&ev);
    for (auto &ev : ctx->user_ev) cudaEventCreate(&ev);
    for (auto &ev : ctx->kev) cudaEventCreate(&ev);
    cudaEventCreateWithFlags(&ctx->ev_clips, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&ctx->stream_side, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LPS_E_CUDA; }
    {
        // PQ = (int)(-10*log10(min/(max+min))) (HaplotagStrategy.cpp:287) tabulated with the HOST libm, so that the
        // truncation to int agrees with the reference on the same machine; the kernel only looks it up
        std::vector<int8_t> lut(256 * 256, 0);
        for (int mn = 1; mn < 256; mn++)
            for (int mx = mn; mx < 256; mx++) {
                int pq = -10 * (std::log10((double)mn / double(mx + mn)));
                lut[(size_t)mn * 256 + mx] = (int8_t)pq;
            }
        if (ctx->d_pq_lut.reserve(lut.size()) != cudaSuccess ||
            cudaMemcpy(ctx->d_pq_lut.p, lut.data(), lut.size(), cudaMemcpyHostToDevice) != cudaSuccess) { delete ctx; return LPS_E_CUDA; }
    }
    if (lps_prepare_call_alleles(ctx) != LPS_OK) { delete ctx; return LPS_E_CUDA; }   // function attributes are per device
    *out = ctx;
    return LPS_OK;
}

void lps_ctx_destroy(lps_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    // every DevBuf / PinBuf member frees itself in its destructor (delete below), with this device current
    for (auto &ev : ctx->ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->user_ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->kev) cudaEventDestroy(ev);
    if (ctx->ev_clips) cudaEventDestroy(ctx->ev_clips);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->stream_side) { cudaStreamSynchronize(ctx->stream_side); cudaStreamDestroy(ctx->stream_side); }
    cudaStreamDestroy(ctx->stream);
    for (auto &cs : ctx->stream_k) if (cs) cudaStreamDestroy(cs);
    if (ctx->stream_up) cudaStreamDestroy(ctx->stream_up);
    if (ctx->stream_down) cudaStreamDestroy(ctx->stream_down);
    delete ctx;
}

const char *lps_last_error(const lps_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int lps_contig_set_reference(lps_ctx *ctx, const char *ref_ascii, int64_t len) {
    if (!ctx || (len > 0 && !ref_ascii) || len < 0) return LPS_E_ARG;
    cudaSetDevice(ctx->device);
    TRY(h2d(ctx, ctx->d_ref, ref_ascii, (size_t)len));
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the caller may free or reuse ref_ascii as soon as this returns
    ctx->ref_len = len;
    ctx->have_variants = false;
    return LPS_OK;
}

int lps_contig_set_variants(lps_ctx *ctx, const lps_variants *v, int is_ont) {
    if (!ctx || !v || v->n < 0) return LPS_E_ARG;
    if (v->n && (!v->pos || !v->ref0 || !v->alt0 || !v->ref_len || !v->alt_len)) return ctx->fail(LPS_E_ARG, "null variant array");
    for (int i = 1; i < v->n; i++)
        if (v->pos[i] <= v->pos[i - 1]) return ctx->fail(LPS_E_ARG, "variant positions must be strictly ascending");
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)v->n;
    TRY(h2d(ctx, ctx->d_vpos, v->pos, n));
    TRY(h2d(ctx, ctx->d_vref0, v->ref0, n));
    TRY(h2d(ctx, ctx->d_valt0, v->alt0, n));
    TRY(h2d(ctx, ctx->d_vref_len, v->ref_len, n));
    TRY(h2d(ctx, ctx->d_valt_len, v->alt_len, n));
    LPS_CUDA(ctx, ctx->d_vhom.reserve(n + 1));
    LPS_CUDA(ctx, ctx->d_vdanger.reserve(n + 1));
    LPS_CUDA(ctx, ctx->d_vfiltered.reserve(n + 1));
    ctx->h_vpos.assign(v->pos, v->pos + n);
    ctx->var.n = v->n;
    ctx->var.pos = ctx->d_vpos.p; ctx->var.ref0 = ctx->d_vref0.p; ctx->var.alt0 = ctx->d_valt0.p;
    ctx->var.ref_len = ctx->d_vref_len.p; ctx->var.alt_len = ctx->d_valt_len.p;
    ctx->var.hom = ctx->d_vhom.p; ctx->var.danger = ctx->d_vdanger.p; ctx->var.filtered = ctx->d_vfiltered.p;
    ctx->is_ont = is_ont;
    ctx->have_tag_variants = false;
    ctx->have_tumor_variants = false;
    ctx->som = DevSomatic();
    if (v->gt_kind) { TRY(h2d(ctx, ctx->d_nor_gt, v->gt_kind, n)); ctx->som.nor_gt = ctx->d_nor_gt.p; }
    if (v->hp1_is_alt && v->ps) {
        TRY(h2d(ctx, ctx->d_vhp1_is_alt, v->hp1_is_alt, n));
        TRY(h2d(ctx, ctx->d_vps, v->ps, n));
        ctx->have_tag_variants = true;
    }
    TRY(lps_launch_annotate(ctx));
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_variants = true;
    ctx->have_calls = false; ctx->have_graph = false;
    return LPS_OK;
}

int lps_contig_get_notes(lps_ctx *ctx, lps_variant_notes *out) {
    if (!ctx || !out) return LPS_E_ARG;
    if (!ctx->have_variants) return ctx->fail(LPS_E_STATE, "lps_contig_set_variants has not been called");
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)ctx->var.n;
    TRY(d2h(ctx, ctx->h_vhom, ctx->d_vhom.p, n));
    TRY(d2h(ctx, ctx->h_vdanger, ctx->d_vdanger.p, n));
    TRY(d2h(ctx, ctx->h_vfiltered, ctx->d_vfiltered.p, n));
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->n = ctx->var.n;
    out->homopolymer = ctx->h_vhom.data(); out->is_danger = ctx->h_vdanger.data(); out->filtered = ctx->h_vfiltered.data();
    return LPS_OK;
}

int lps_batch_submit(lps_ctx *ctx, const lps_read_batch *b) {
    if (!ctx || !b || b->n_reads < 0) return LPS_E_ARG;
    const size_t n = (size_t)b->n_reads;
    if (n && (!b->ref_start || !b->l_qseq || !b->n_cigar || !b->cigar_off || !b->seq_off || (!b->qual_off && !b->sq) || !b->flag || !b->mapq ||
              !b->name_rank))
        return ctx->fail(LPS_E_ARG, "null read array");
    TRY(validate_batch(ctx, b));
    cudaSetDevice(ctx->device);
    cudaEventRecord(ctx->ev[0], ctx->stream);
    TRY(h2d(ctx, ctx->d_ref_start, b->ref_start, n));
    TRY(h2d(ctx, ctx->d_l_qseq, b->l_qseq, n));
    TRY(h2d(ctx, ctx->d_n_cigar, b->n_cigar, n));
    TRY(h2d(ctx, ctx->d_cigar_off, b->cigar_off, n));
    TRY(h2d(ctx, ctx->d_seq_off, b->seq_off, n));
    if (!b->sq) TRY(h2d(ctx, ctx->d_qual_off, b->qual_off, n));
    TRY(h2d(ctx, ctx->d_flag, b->flag, n));
    TRY(h2d(ctx, ctx->d_mapq, b->mapq, n));
    TRY(h2d(ctx, ctx->d_name_rank, b->name_rank, n));
    if (b->cigar8) {
        if (b->n_cigar_long && (!b->cigar_long_len || !b->cigar_long_at)) return ctx->fail(LPS_E_ARG, "cigar8 without its long-op table");
        if (!b->cigar_esc_blk || (b->n_cigar_esc && !b->cigar_esc16)) return ctx->fail(LPS_E_ARG, "cigar8 without its escape tables");
        const size_t n_blk = (size_t)(b->cigar_len / 256) + 2;
        if (b->cigar_esc_blk[0] != 0 || b->cigar_esc_blk[(b->cigar_len + 255) / 256] != b->n_cigar_esc)
            return ctx->fail(LPS_E_ARG, "cigar_esc_blk does not match n_cigar_esc");
        for (uint64_t i = 0; i < b->n_cigar_long; i++)
            if (b->cigar_long_at[i] >= b->cigar_len || (b->cigar_long_len[i] >> 28) || (i && b->cigar_long_at[i] <= b->cigar_long_at[i - 1]))
                return ctx->fail(LPS_E_ARG, "cigar_long_at / cigar_long_len are not a valid side table");
        TRY(h2d(ctx, ctx->d_cigar8, b->cigar8, (size_t)b->cigar_len, 64));
        TRY(h2d(ctx, ctx->d_cigar_esc16, b->cigar_esc16, (size_t)b->n_cigar_esc));
        TRY(h2d(ctx, ctx->d_cigar_esc_blk, b->cigar_esc_blk, n_blk));
        TRY(h2d(ctx, ctx->d_cigar_long_len, b->cigar_long_len, (size_t)b->n_cigar_long));
        TRY(h2d(ctx, ctx->d_cigar_long_at, b->cigar_long_at, (size_t)b->n_cigar_long));
        LPS_CUDA(ctx, ctx->d_cigar16.reserve((size_t)b->cigar_len + 64 + 256));
        if (b->cigar_len) {
            const size_t warps = ((size_t)b->cigar_len + 255) / 256;
            k_expand_cigar8<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, ctx->stream>>>((size_t)b->cigar_len, ctx->d_cigar8.p, ctx->d_cigar_esc16.p,
                                                                                           ctx->d_cigar_esc_blk.p, ctx->d_cigar16.p);
            ctx->stats.kernel_launches++;
            LPS_CUDA(ctx, cudaGetLastError());
        }
        ctx->batch.cigar16 = ctx->d_cigar16.p; ctx->batch.long_at = ctx->d_cigar_long_at.p; ctx->batch.long_len = ctx->d_cigar_long_len.p;
        ctx->batch.n_long = (uint32_t)b->n_cigar_long;
    } else if (b->cigar16) {
        if (b->n_cigar_long && (!b->cigar_long_len || !b->cigar_long_at)) return ctx->fail(LPS_E_ARG, "cigar16 without its long-op table");
        for (uint64_t i = 0; i < b->n_cigar_long; i++)
            if (b->cigar_long_at[i] >= b->cigar_len || (b->cigar16[b->cigar_long_at[i]] >> 4) != 0xFFFu || (b->cigar_long_len[i] >> 28) ||
                (i && b->cigar_long_at[i] <= b->cigar_long_at[i - 1]))
                return ctx->fail(LPS_E_ARG, "cigar_long_at / cigar_long_len do not match cigar16");
        TRY(h2d(ctx, ctx->d_cigar16, b->cigar16, (size_t)b->cigar_len, 64));
        TRY(h2d(ctx, ctx->d_cigar_long_len, b->cigar_long_len, (size_t)b->n_cigar_long));
        TRY(h2d(ctx, ctx->d_cigar_long_at, b->cigar_long_at, (size_t)b->n_cigar_long));
        ctx->batch.cigar16 = ctx->d_cigar16.p; ctx->batch.long_at = ctx->d_cigar_long_at.p; ctx->batch.long_len = ctx->d_cigar_long_len.p;
        ctx->batch.n_long = (uint32_t)b->n_cigar_long;
    } else {
        if (b->cigar_len && !b->cigar) return ctx->fail(LPS_E_ARG, "null CIGAR stream");
        TRY(h2d(ctx, ctx->d_cigar, b->cigar, (size_t)b->cigar_len, 64));
        TRY(narrow_cigar(ctx, ctx->d_cigar.p, b->cigar_len));
    }
    // SEQ and QUAL are 85 % of a batch but the kernels touch ~2 bytes per allele call of them.  When the caller's
    // buffers are pinned (cudaHostAlloc / cudaHostRegister) they stay on the host and the resolve phase of the kernel
    // gathers the few sectors it needs straight over PCIe (UVA zero-copy); pageable buffers are copied as before.
    const uint8_t *seq_dev = nullptr, *qual_dev = nullptr;
    bool zero_copy = false;
    if (b->sq) {
        // interleaved SEQ + QUAL rows: one array, one 16-byte unit per allele call; pinned rows stay on the host (they must be
        // 16-byte aligned there: the kernel loads whole units), pageable rows are copied
        const char *env = getenv("LPS_ZERO_COPY");
        cudaPointerAttributes as;
        if (!(env && env[0] == '0') && b->sq_bytes && cudaPointerGetAttributes(&as, b->sq) == cudaSuccess && as.type == cudaMemoryTypeHost &&
            as.devicePointer && ((uintptr_t)as.devicePointer & 15u) == 0) {
            seq_dev = (const uint8_t *)as.devicePointer; zero_copy = true;
        }
        cudaGetLastError();
        if (!zero_copy) {
            TRY(h2d(ctx, ctx->d_seq4, b->sq, (size_t)b->sq_bytes, 16));
            seq_dev = ctx->d_seq4.p;
        }
    } else {
        const char *env = getenv("LPS_ZERO_COPY");
        cudaPointerAttributes as, aq;
        if (!(env && env[0] == '0') && b->seq_bytes && b->qual_bytes &&
            cudaPointerGetAttributes(&as, b->seq4) == cudaSuccess && cudaPointerGetAttributes(&aq, b->qual) == cudaSuccess &&
            as.type == cudaMemoryTypeHost && aq.type == cudaMemoryTypeHost && as.devicePointer && aq.devicePointer) {
            seq_dev = (const uint8_t *)as.devicePointer; qual_dev = (const uint8_t *)aq.devicePointer; zero_copy = true;
        }
        cudaGetLastError();   // a pageable pointer makes cudaPointerGetAttributes report an error on old drivers
        if (!zero_copy) {
            TRY(h2d(ctx, ctx->d_seq4, b->seq4, (size_t)b->seq_bytes, 16));
            TRY(h2d(ctx, ctx->d_qual, b->qual, (size_t)b->qual_bytes, 16));
            seq_dev = ctx->d_seq4.p; qual_dev = ctx->d_qual.p;
        }
    }
    ctx->zero_copy = zero_copy;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->h_name_rank.assign(b->name_rank, b->name_rank + n);
    ctx->h_flag.assign(b->flag, b->flag + n);
    lps_host_index_names(ctx);
    TRY(h2d(ctx, ctx->d_multi_members, ctx->h_multi_members.data(), ctx->h_multi_members.size()));
    TRY(h2d(ctx, ctx->d_multi_group_off, ctx->h_multi_group_off.data(), ctx->h_multi_group_off.size()));
    uint64_t s = 0;
    for (size_t i = 0; i < n; i++) s += (uint64_t)(b->l_qseq[i] > 0 ? b->l_qseq[i] : 0);
    ctx->sum_l_qseq = s;
    DevBatch &d = ctx->batch;
    d.n_reads = b->n_reads;
    d.ref_start = ctx->d_ref_start.p; d.l_qseq = ctx->d_l_qseq.p; d.n_cigar = ctx->d_n_cigar.p;
    d.cigar_off = ctx->d_cigar_off.p; d.seq_off = ctx->d_seq_off.p; d.qual_off = b->sq ? nullptr : ctx->d_qual_off.p;
    d.flag = ctx->d_flag.p; d.mapq = ctx->d_mapq.p; d.name_rank = ctx->d_name_rank.p;
    d.cigar_len = b->cigar_len; d.seq4 = seq_dev; d.seq_bytes = b->sq ? b->sq_bytes : b->seq_bytes;
    d.qual = qual_dev; d.qual_bytes = b->sq ? 0 : b->qual_bytes;
    d.sq = b->sq ? 1 : 0;
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.ms_h2d = elapsed(ctx, 0, 1);
    ctx->have_batch = true; ctx->have_calls = false; ctx->have_graph = false;
    return LPS_OK;
}

int lps_pack_cigar16(const uint32_t *cigar, uint64_t n, uint64_t base_index, uint16_t *out16, uint32_t *long_len, uint64_t *long_at,
                     uint64_t long_cap, uint64_t *n_long) {
    if ((n && (!cigar || !out16)) || !n_long) return LPS_E_ARG;
    uint64_t nl = *n_long;
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t w = cigar[i];
        if ((w >> 4) < 0xFFFu) { out16[i] = (uint16_t)w; continue; }
        if (nl >= long_cap || !long_len || !long_at) return LPS_E_ARG;
        out16[i] = (uint16_t)(0xFFF0u | (w & 15u));
        long_len[nl] = w >> 4; long_at[nl] = base_index + i; nl++;
    }
    *n_long = nl;
    return LPS_OK;
}

int lps_pack_cigar8(const uint32_t *cigar, uint64_t n, uint64_t base_index, uint8_t *out8, uint16_t *esc16, uint64_t esc_cap, uint64_t *n_esc,
                    uint32_t *esc_blk, uint32_t *long_len, uint64_t *long_at, uint64_t long_cap, uint64_t *n_long) {
    if ((n && (!cigar || !out8)) || !n_esc || !n_long || !esc_blk) return LPS_E_ARG;
    uint64_t ne = *n_esc, nl = *n_long;
    if (base_index == 0) esc_blk[0] = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t at = base_index + i;
        if ((at & 255u) == 0) esc_blk[at >> 8] = (uint32_t)ne;
        const uint32_t w = cigar[i], op = w & 15u, len = w >> 4;
        if (op == 0 && len >= 1 && len <= 128) out8[i] = (uint8_t)(len - 1);
        else if (op == 1 && len >= 1 && len <= 56) out8[i] = (uint8_t)(0x80u + len - 1);
        else if (op == 2 && len >= 1 && len <= 56) out8[i] = (uint8_t)(0xB8u + len - 1);
        else {
            if (ne >= esc_cap || !esc16) return LPS_E_ARG;
            out8[i] = 0xFFu;
            if (len < 0xFFFu) esc16[ne++] = (uint16_t)w;
            else {
                if (nl >= long_cap || !long_len || !long_at) return LPS_E_ARG;
                esc16[ne++] = (uint16_t)(0xFFF0u | op);
                long_len[nl] = len; long_at[nl] = at; nl++;
            }
        }
    }
    // the entries a reader of the stream so far needs: the block after the last op (ceil) carries the running total
    esc_blk[(base_index + n + 255) >> 8] = (uint32_t)ne;
    esc_blk[((base_index + n) >> 8) + 1] = (uint32_t)ne;
    *n_esc = ne; *n_long = nl;
    return LPS_OK;
}

uint64_t lps_sq_row_bytes(int32_t l_qseq) { return l_qseq > 0 ? (uint64_t)SQ_UNIT * (((uint64_t)l_qseq + SQ_BASES - 1) / SQ_BASES) : 0; }

int lps_pack_sq(const uint8_t *seq4, const uint8_t *qual, int32_t l_qseq, uint8_t *out) {
    if (l_qseq < 0 || (l_qseq > 0 && (!seq4 || !qual || !out))) return LPS_E_ARG;
    const uint32_t lq = (uint32_t)l_qseq, seq_bytes = (lq + 1) / 2;
    for (uint32_t at = 0, u = 0; at < lq; at += SQ_BASES, u++) {
        uint8_t *unit = out + (size_t)u * SQ_UNIT;
        const uint32_t nb = lq - at < SQ_BASES ? lq - at : SQ_BASES;        // bases of this unit
        memcpy(unit, qual + at, nb);
        memset(unit + nb, 0, SQ_UNIT - nb);
        const uint32_t s0 = at / 2, ns = seq_bytes - s0 < SQ_BASES / 2 ? seq_bytes - s0 : SQ_BASES / 2;   // at is even: unit starts on a byte of seq4
        memcpy(unit + SQ_BASES, seq4 + s0, ns);
    }
    return LPS_OK;
}

int lps_pack_sq_batch(int32_t n_reads, const int32_t *l_qseq, const uint64_t *seq_off, const uint64_t *qual_off, const uint8_t *seq4,
                      const uint8_t *qual, const uint64_t *sq_off, uint8_t *sq) {
    if (n_reads < 0 || (n_reads && (!l_qseq || !seq_off || !qual_off || !sq_off))) return LPS_E_ARG;
    for (int32_t r = 0; r < n_reads; r++) {
        if (l_qseq[r] <= 0) continue;
        if (sq_off[r] & 15u) return LPS_E_ARG;
        const int rc = lps_pack_sq(seq4 + seq_off[r], qual + qual_off[r], l_qseq[r], sq + sq_off[r]);
        if (rc != LPS_OK) return rc;
    }
    return LPS_OK;
}

int lps_sq_peek(const uint8_t *row, int32_t l_qseq, int32_t qi, uint8_t *seq_code, uint8_t *quality) {
    if (!row || qi < 0 || qi >= l_qseq) return LPS_E_ARG;
    const uint32_t u = sq_unit_of((uint32_t)qi);
    uint32_t w[4];
    memcpy(w, row + (size_t)u * SQ_UNIT, SQ_UNIT);                          // little-endian host, like the device
    unsigned code, q;
    sq_extract(w[0], w[1], w[2], w[3], (uint32_t)qi - u * SQ_BASES, code, q);
    if (seq_code) *seq_code = (uint8_t)code;
    if (quality) *quality = (uint8_t)q;
    return LPS_OK;
}

int lps_batch_submit_device(lps_ctx *ctx, const lps_read_batch *b) {
    if (!ctx || !b || b->n_reads < 0) return LPS_E_ARG;
    cudaSetDevice(ctx->device);
    DevBatch &d = ctx->batch;
    d.n_reads = b->n_reads;
    d.ref_start = b->ref_start; d.l_qseq = b->l_qseq; d.n_cigar = b->n_cigar; d.cigar_off = b->cigar_off;
    d.seq_off = b->seq_off; d.qual_off = b->qual_off; d.flag = b->flag; d.mapq = b->mapq; d.name_rank = b->name_rank;
    d.cigar_len = b->cigar_len; d.seq4 = b->seq4; d.seq_bytes = b->seq_bytes;
    d.qual = b->qual; d.qual_bytes = b->qual_bytes;
    d.sq = 0;
    if (b->sq) {
        // interleaved SEQ + QUAL rows used where they lie (phase calls only): the kernel loads whole 16-byte units
        if (((uintptr_t)b->sq & 15u) != 0) return ctx->fail(LPS_E_ARG, "a device-resident sq stream must be 16-byte aligned");
        d.seq4 = b->sq; d.seq_bytes = b->sq_bytes; d.qual = nullptr; d.qual_bytes = 0; d.qual_off = nullptr; d.sq = 1;
    }
    if (b->cigar16) {
        // the 16-bit stream is used where it lies: the bulk copies of k_call_alleles need a 16-byte aligned base
        if (((uintptr_t)b->cigar16 & 15u) != 0) return ctx->fail(LPS_E_ARG, "a device-resident cigar16 stream must be 16-byte aligned");
        if (b->n_cigar_long && (!b->cigar_long_len || !b->cigar_long_at)) return ctx->fail(LPS_E_ARG, "cigar16 without its long-op table");
        if (b->n_cigar_long >> 32) return ctx->fail(LPS_E_ARG, "too many escaped CIGAR ops");
        d.cigar16 = b->cigar16; d.long_at = b->cigar_long_at; d.long_len = b->cigar_long_len; d.n_long = (uint32_t)b->n_cigar_long;
    } else {
        if (b->cigar_len && !b->cigar) return ctx->fail(LPS_E_ARG, "null CIGAR stream");
        TRY(narrow_cigar(ctx, b->cigar, b->cigar_len));
    }
    TRY(d2h(ctx, ctx->h_name_rank, b->name_rank, (size_t)b->n_reads));
    TRY(d2h(ctx, ctx->h_flag, b->flag, (size_t)b->n_reads));
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    lps_host_index_names(ctx);
    TRY(h2d(ctx, ctx->d_multi_members, ctx->h_multi_members.data(), ctx->h_multi_members.size()));
    TRY(h2d(ctx, ctx->d_multi_group_off, ctx->h_multi_group_off.data(), ctx->h_multi_group_off.size()));
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->sum_l_qseq = b->sq ? b->sq_bytes / SQ_UNIT * SQ_BASES : b->qual_bytes;   // sizes the call pool (an upper bound; the pool grows on demand)
    ctx->have_batch = true; ctx->have_calls = false; ctx->have_graph = false;
    return LPS_OK;
}

int lps_phase_call_alleles(lps_ctx *ctx, const lps_phase_params *p, int want_host, lps_calls *out) {
    if (!ctx || !p) return LPS_E_ARG;
    if (!ctx->have_variants || !ctx->have_batch) return ctx->fail(LPS_E_STATE, "variants and a read batch must be set first");
    cudaSetDevice(ctx->device);
    WallTimer wt;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    TRY(lps_launch_call_alleles(ctx, p));
    cudaEventRecord(ctx->ev[3], ctx->stream);
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.ms_call_alleles = elapsed(ctx, 2, 3);
    ctx->have_graph = false;
    if (out) {
        memset(out, 0, sizeof(*out));
        out->n_reads = ctx->batch.n_reads;
        out->n_calls = ctx->n_calls;
        out->n_clips = (int32_t)ctx->h_clip_pos.size();
        out->clip_pos = ctx->h_clip_pos.data(); out->clip_front = ctx->h_clip_front.data(); out->clip_back = ctx->h_clip_back.data();
        if (want_host) {
            cudaEventRecord(ctx->ev[4], ctx->stream);
            TRY(fetch_host_calls(ctx));
            cudaEventRecord(ctx->ev[5], ctx->stream);
            LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->stats.ms_d2h = elapsed(ctx, 4, 5);
            out->call_off = ctx->h_call_off.data(); out->calls = ctx->h_calls.data(); out->read_status = ctx->h_status.data();
        }
    }
    ctx->stats.ms_wall_call_alleles = wt.ms();
    return LPS_OK;
}

int lps_tag_reads(lps_ctx *ctx, const lps_tag_params *p, int want_calls, lps_tag_result *out) {
    if (!ctx || !p) return LPS_E_ARG;
    if (!ctx->have_variants || !ctx->have_batch) return ctx->fail(LPS_E_STATE, "variants and a read batch must be set first");
    if (!ctx->have_tag_variants) return ctx->fail(LPS_E_STATE, "the variant table lacks hp1_is_alt / ps (phased VCF fields)");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)ctx->batch.n_reads;
    cudaEventRecord(ctx->ev[2], st);
    TRY(lps_launch_call_alleles(ctx, nullptr, p, want_calls));
    cudaEventRecord(ctx->ev[3], st);
    TRY(d2h(ctx, ctx->h_tag_cat, ctx->d_tag_cat.p, n));
    TRY(d2h(ctx, ctx->h_tag_hp, ctx->d_tag_hp.p, n));
    TRY(d2h(ctx, ctx->h_tag_ps, ctx->d_tag_ps.p, n));
    TRY(d2h(ctx, ctx->h_tag_pq, ctx->d_tag_pq.p, n));
    TRY(d2h(ctx, ctx->h_tag_h1, ctx->d_tag_h1.p, n));
    TRY(d2h(ctx, ctx->h_tag_h2, ctx->d_tag_h2.p, n));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->stats.ms_tag_reads = elapsed(ctx, 2, 3);
    ctx->have_calls = false; ctx->have_graph = false; ctx->host_calls_valid = false;
    if (want_calls) {
        TRY(d2h(ctx, ctx->h_call_off, ctx->d_call_off.p, n + 1));
        TRY(d2h(ctx, ctx->h_calls, ctx->d_calls.p, (size_t)ctx->n_calls));
        LPS_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (out) memset(out, 0, sizeof(*out));
    // ReadStatistics (HaplotagProcess.cpp:282-354) and the PQ values beyond the device table
    lps_tag_result acc;
    memset(&acc, 0, sizeof(acc));
    for (size_t r = 0; r < n; r++) {
        acc.total_alignment++;
        const int cat = ctx->h_tag_cat[r];
        if (cat == LPS_TAG_PROCESSED) {
            if (ctx->h_flag[r] & 0x800) acc.total_supplementary++;
            const int h1 = ctx->h_tag_h1[r], h2 = ctx->h_tag_h2[r];
            const double mx = h1 > h2 ? h1 : h2, mn = h1 > h2 ? h2 : h1;
            if (ctx->h_tag_pq[r] < 0) ctx->h_tag_pq[r] = -10 * (std::log10((double)mn / double(mx + mn)));
            if (mx / (mx + mn) < p->percentage_threshold) acc.total_high_similarity++;
            if (mx == 0) acc.total_without_variant++;
            const int hp = ctx->h_tag_hp[r];
            if (hp == 0) { acc.total_hp0++; acc.total_untag++; }
            else { if (hp == 1) acc.total_hp1++; else acc.total_hp2++; acc.total_tag++; }
        } else {
            acc.total_untag++;
            if (cat == LPS_TAG_LOW_MAPQ) acc.total_lower_quality++;
            else if (cat == LPS_TAG_UNMAPPED) acc.total_unmapped++;
            else if (cat == LPS_TAG_SECONDARY) acc.total_secondary++;
            else if (cat == LPS_TAG_SUPPLEMENTARY) acc.total_supplementary++;
            else if (cat == LPS_TAG_EMPTY_VARIANTS) acc.total_empty_variant++;
            else acc.total_other_case++;
        }
    }
    if (out) {
        *out = acc;
        out->n_reads = ctx->batch.n_reads;
        out->category = ctx->h_tag_cat.data(); out->hp = ctx->h_tag_hp.data(); out->ps = ctx->h_tag_ps.data();
        out->pq = ctx->h_tag_pq.data(); out->h1 = ctx->h_tag_h1.data(); out->h2 = ctx->h_tag_h2.data();
        if (want_calls) { out->n_calls = ctx->n_calls; out->call_off = ctx->h_call_off.data(); out->calls = ctx->h_calls.data(); }
    }
    return LPS_OK;
}

int lps_contig_set_tumor_variants(lps_ctx *ctx, const lps_tumor_variants *t) {
    if (!ctx || !t) return LPS_E_ARG;
    if (!ctx->have_variants || !ctx->have_tag_variants)
        return ctx->fail(LPS_E_STATE, "lps_contig_set_variants (with hp1_is_alt / ps) must be called first");
    if (t->n != ctx->var.n) return ctx->fail(LPS_E_ARG, "the tumor arrays must be parallel to the variant table");
    const size_t n = (size_t)t->n;
    if (n && (!t->tum_present || !t->ref0 || !t->alt0 || !t->ref_len || !t->alt_len || !t->gt_kind || !t->ps || !t->is_somatic || !t->derive_hp))
        return ctx->fail(LPS_E_ARG, "null tumor variant array");
    cudaSetDevice(ctx->device);
    std::vector<uint8_t> nor_gt;
    if (ctx->som.nor_gt) { TRY(d2h(ctx, nor_gt, ctx->som.nor_gt, n)); LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); }
    std::vector<int32_t> slot(n, -1), prev_nor(n, INT_MIN);
    ctx->h_tum_var.clear();
    int32_t last_nor = INT_MIN;
    for (size_t i = 0; i < n; i++) {
        const bool N = t->nor_present ? t->nor_present[i] != 0 : true, T = t->tum_present[i] != 0;
        if (!N && !T) return ctx->fail(LPS_E_ARG, "a position of the union map holds neither a NORMAL nor a TUMOR record");
        if (T) {
            if (t->gt_kind[i] == 1 && t->ps[i] == -1) return ctx->fail(LPS_E_ARG, "phased tumor record without a phase set (HaplotagStrategy.cpp:337)");
            slot[i] = (int32_t)ctx->h_tum_var.size();
            ctx->h_tum_var.push_back((int32_t)i);
        } else if (t->is_somatic[i]) return ctx->fail(LPS_E_ARG, "is_somatic set on a position without a TUMOR record");
        if (t->derive_hp[i] < 0 || t->derive_hp[i] > 2) return ctx->fail(LPS_E_ARG, "derive_hp must be 0, 1 or 2 (HaplotagLogging.cpp:41)");
        prev_nor[i] = last_nor;
        if (N && (nor_gt.empty() || nor_gt[i] == 1)) last_nor = ctx->h_vpos[i];
    }
    const size_t nt = ctx->h_tum_var.size();
    ctx->h_t_alt0.assign(t->alt0, t->alt0 + n); ctx->h_t_ref_len.assign(t->ref_len, t->ref_len + n); ctx->h_t_alt_len.assign(t->alt_len, t->alt_len + n);
    DevSomatic &s = ctx->som;
    if (t->nor_present) { TRY(h2d(ctx, ctx->d_nor_present, t->nor_present, n)); s.nor_present = ctx->d_nor_present.p; } else s.nor_present = nullptr;
    TRY(h2d(ctx, ctx->d_tum_present, t->tum_present, n)); s.tum_present = ctx->d_tum_present.p;
    TRY(h2d(ctx, ctx->d_t_ref0, t->ref0, n)); s.t_ref0 = ctx->d_t_ref0.p;
    TRY(h2d(ctx, ctx->d_t_alt0, t->alt0, n)); s.t_alt0 = ctx->d_t_alt0.p;
    TRY(h2d(ctx, ctx->d_t_ref_len, t->ref_len, n)); s.t_ref_len = ctx->d_t_ref_len.p;
    TRY(h2d(ctx, ctx->d_t_alt_len, t->alt_len, n)); s.t_alt_len = ctx->d_t_alt_len.p;
    TRY(h2d(ctx, ctx->d_t_gt, t->gt_kind, n)); s.t_gt = ctx->d_t_gt.p;
    TRY(h2d(ctx, ctx->d_is_somatic, t->is_somatic, n)); s.is_somatic = ctx->d_is_somatic.p;
    TRY(h2d(ctx, ctx->d_derive_hp, t->derive_hp, n)); s.derive_hp = ctx->d_derive_hp.p;
    TRY(h2d(ctx, ctx->d_slot_of_var, slot.data(), n)); s.slot_of_var = ctx->d_slot_of_var.p;
    TRY(h2d(ctx, ctx->d_prev_nor, prev_nor.data(), n)); s.prev_nor = ctx->d_prev_nor.p;
    TRY(h2d(ctx, ctx->d_tum_var, ctx->h_tum_var.data(), nt)); s.tum_var = ctx->d_tum_var.p;
    s.n_tum = (int32_t)nt;
    // one block of per-slot counters; cover_start / cover_end are adjacent (initialised together by the launcher)
    const size_t words[] = {LPS_PB_FIELDS, 9, 9, LPS_CASE_FIELDS, 2, 9, 9, 9, 9, 1, 1, 2 * LPS_WINDOW_BINS};
    size_t total = 0;
    for (size_t w : words) total += w * nt;
    LPS_CUDA(ctx, ctx->d_som_counters.reserve(total + 1));
    ctx->som_counter_words = total;
    int32_t *q = ctx->d_som_counters.p;
    int32_t **dst[] = {&s.pos_base, &s.read_hp_count, &s.somatic_read_hp_count, &s.case_count, &s.allele_count, &s.hp_before_count,
                       &s.hp_after_count, &s.h3_before_count, &s.h3_after_count, &s.cover_start, &s.cover_end, &s.window_hist};
    for (size_t k = 0; k < sizeof(words) / sizeof(words[0]); k++) { *dst[k] = q; q += words[k] * nt; }
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_tumor_variants = true;
    return LPS_OK;
}

namespace {

// per-alignment products of the somatic family -> host; PQ values beyond the device table use the host libm
int fetch_read_tags(lps_ctx *ctx, bool somatic_pq, lps_read_tags *out) {
    const size_t n = (size_t)ctx->batch.n_reads;
    TRY(d2h(ctx, ctx->h_tag_cat, ctx->d_tag_cat.p, n));
    TRY(d2h(ctx, ctx->h_tag_hp, ctx->d_tag_hp.p, n));
    TRY(d2h(ctx, ctx->h_tag_ps, ctx->d_tag_ps.p, n));
    TRY(d2h(ctx, ctx->h_tag_pq, ctx->d_tag_pq.p, n));
    TRY(d2h(ctx, ctx->h_tag_h1, ctx->d_tag_h1.p, n));
    TRY(d2h(ctx, ctx->h_tag_h2, ctx->d_tag_h2.p, n));
    TRY(d2h(ctx, ctx->h_tag_h3, ctx->d_tag_h3.p, n));
    TRY(d2h(ctx, ctx->h_tag_nps, ctx->d_tag_nps.p, n));
    TRY(d2h(ctx, ctx->h_tag_end, ctx->d_tag_end.p, n));
    TRY(d2h(ctx, ctx->h_tag_len, ctx->d_tag_len.p, n));
    TRY(d2h(ctx, ctx->h_tag_hpb, ctx->d_tag_hpb.p, n));
    TRY(d2h(ctx, ctx->h_tag_sim, ctx->d_tag_sim.p, n));
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t r = 0; r < n; r++)
        if (ctx->h_tag_pq[r] < 0) {
            const int h1 = ctx->h_tag_h1[r], h2 = ctx->h_tag_h2[r];
            const double mx = h1 > h2 ? h1 : h2, mn = h1 > h2 ? h2 : h1;
            ctx->h_tag_pq[r] = -10 * (std::log10((double)mn / double(mx + mn)));
        }
    (void)somatic_pq;
    out->n_reads = ctx->batch.n_reads;
    out->category = ctx->h_tag_cat.data(); out->read_hp = ctx->h_tag_hp.data(); out->ps = ctx->h_tag_ps.data(); out->pq = ctx->h_tag_pq.data();
    out->h1 = ctx->h_tag_h1.data(); out->h2 = ctx->h_tag_h2.data(); out->h3 = ctx->h_tag_h3.data(); out->n_ps = ctx->h_tag_nps.data();
    out->end_pos = ctx->h_tag_end.data(); out->read_len = ctx->h_tag_len.data();
    return LPS_OK;
}

int run_extract(lps_ctx *ctx, const lps_tag_params *p, int mode, lps_extract_result *out) {
    if (!ctx || !p) return LPS_E_ARG;
    if (!ctx->have_variants || !ctx->have_batch) return ctx->fail(LPS_E_STATE, "variants and a read batch must be set first");
    if (!ctx->have_tumor_variants) return ctx->fail(LPS_E_STATE, "lps_contig_set_tumor_variants has not been called");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const bool tumor = mode == LPS_MODE_EXTRACT_TUMOR;
    cudaEventRecord(ctx->ev[2], st);
    TRY(lps_launch_call_alleles(ctx, nullptr, p, tumor ? 1 : 0, mode));
    if (tumor) TRY(lps_launch_window_diff(ctx, p->have_reference));
    cudaEventRecord(ctx->ev[3], st);
    TRY(d2h(ctx, ctx->h_som_counters, ctx->d_som_counters.p, ctx->som_counter_words));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->stats.ms_tag_reads = elapsed(ctx, 2, 3);
    ctx->have_calls = false; ctx->have_graph = false; ctx->host_calls_valid = false;
    if (!out) return LPS_OK;
    memset(out, 0, sizeof(*out));
    TRY(fetch_read_tags(ctx, tumor, &out->reads));
    const DevSomatic &s = ctx->som;
    const int32_t *h = ctx->h_som_counters.data(), *d = ctx->d_som_counters.p;
    out->n_tum = s.n_tum; out->tum_var = ctx->h_tum_var.data();
    out->pos_base = h + (s.pos_base - d); out->read_hp_count = h + (s.read_hp_count - d);
    ctx->h_ratios_f.resize((size_t)s.n_tum * LPS_RF_FIELDS + 1); ctx->h_ratios_d.resize((size_t)s.n_tum * LPS_RD_FIELDS + 1);
    ctx->h_case_reads.resize((size_t)s.n_tum + 1);
    lps_host_post_process(s.n_tum, ctx->h_tum_var.data(), ctx->h_t_alt0.data(), ctx->h_t_ref_len.data(), ctx->h_t_alt_len.data(), out->pos_base,
                          out->read_hp_count, h + (s.case_count - d), tumor, ctx->h_ratios_f.data(), ctx->h_ratios_d.data(), ctx->h_case_reads.data());
    out->ratios_f = ctx->h_ratios_f.data(); out->ratios_d = ctx->h_ratios_d.data(); out->case_read_count = ctx->h_case_reads.data();
    if (tumor) {
        out->somatic_read_hp_count = h + (s.somatic_read_hp_count - d); out->case_count = h + (s.case_count - d);
        out->allele_count = h + (s.allele_count - d); out->window_hist = h + (s.window_hist - d);
        out->n_window_items = ctx->n_wd_items;
        const size_t n = (size_t)ctx->batch.n_reads;
        TRY(d2h(ctx, ctx->h_call_off, ctx->d_call_off.p, n + 1));
        TRY(d2h(ctx, ctx->h_calls, ctx->d_calls.p, (size_t)ctx->n_calls));
        LPS_CUDA(ctx, cudaStreamSynchronize(st));
        out->n_calls = ctx->n_calls; out->call_off = ctx->h_call_off.data(); out->calls = ctx->h_calls.data();
    }
    return LPS_OK;
}

}  // namespace

int lps_extract_normal(lps_ctx *ctx, const lps_tag_params *p, lps_extract_result *out) { return run_extract(ctx, p, LPS_MODE_EXTRACT_NORMAL, out); }
int lps_extract_tumor(lps_ctx *ctx, const lps_tag_params *p, lps_extract_result *out) { return run_extract(ctx, p, LPS_MODE_EXTRACT_TUMOR, out); }

int lps_somatic_tag_reads(lps_ctx *ctx, const lps_tag_params *p, int want_calls, lps_somatic_tag_result *out) {
    if (!ctx || !p) return LPS_E_ARG;
    if (!ctx->have_variants || !ctx->have_batch) return ctx->fail(LPS_E_STATE, "variants and a read batch must be set first");
    if (!ctx->have_tumor_variants) return ctx->fail(LPS_E_STATE, "lps_contig_set_tumor_variants has not been called");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)ctx->batch.n_reads;
    cudaEventRecord(ctx->ev[2], st);
    TRY(lps_launch_call_alleles(ctx, nullptr, p, want_calls, LPS_MODE_SOMATIC_TAG));
    cudaEventRecord(ctx->ev[3], st);
    TRY(d2h(ctx, ctx->h_som_counters, ctx->d_som_counters.p, ctx->som_counter_words));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->stats.ms_tag_reads = elapsed(ctx, 2, 3);
    ctx->have_calls = false; ctx->have_graph = false; ctx->host_calls_valid = false;
    if (!out) return LPS_OK;
    memset(out, 0, sizeof(*out));
    TRY(fetch_read_tags(ctx, true, &out->reads));
    out->hp_before = ctx->h_tag_hpb.data(); out->derive_similarity = ctx->h_tag_sim.data();
    const DevSomatic &s = ctx->som;
    const int32_t *h = ctx->h_som_counters.data(), *d = ctx->d_som_counters.p;
    out->n_tum = s.n_tum; out->tum_var = ctx->h_tum_var.data();
    out->hp_before_count = h + (s.hp_before_count - d); out->hp_after_count = h + (s.hp_after_count - d);
    out->h3_before_count = h + (s.h3_before_count - d); out->h3_after_count = h + (s.h3_after_count - d);
    out->cover_start = h + (s.cover_start - d); out->cover_end = h + (s.cover_end - d);
    if (want_calls) {
        TRY(d2h(ctx, ctx->h_call_off, ctx->d_call_off.p, n + 1));
        TRY(d2h(ctx, ctx->h_calls, ctx->d_calls.p, (size_t)ctx->n_calls));
        LPS_CUDA(ctx, cudaStreamSynchronize(st));
        out->n_calls = ctx->n_calls; out->call_off = ctx->h_call_off.data(); out->calls = ctx->h_calls.data();
    }
    // ReadStatistics (HaplotagProcess.cpp:282-354; SomaticHaplotagProcess.cpp:352, 398-402; HaplotagStrategy.cpp:452-602)
    for (size_t r = 0; r < n; r++) {
        out->total_alignment++;
        const int cat = ctx->h_tag_cat[r];
        if (cat != LPS_TAG_PROCESSED) {
            out->total_untag++;
            if (cat == LPS_TAG_LOW_MAPQ) out->total_lower_quality++;
            else if (cat == LPS_TAG_UNMAPPED) out->total_unmapped++;
            else if (cat == LPS_TAG_SECONDARY) out->total_secondary++;
            else if (cat == LPS_TAG_SUPPLEMENTARY) out->total_supplementary++;
            else if (cat == LPS_TAG_EMPTY_VARIANTS) out->total_empty_variant++;
            else out->total_other_case++;
            continue;
        }
        if (ctx->h_flag[r] & 0x800) out->total_supplementary++;
        const int h1 = ctx->h_tag_h1[r], h2 = ctx->h_tag_h2[r], h3 = ctx->h_tag_h3[r], hp = ctx->h_tag_hp[r];
        const double mx = h1 > h2 ? h1 : h2, mn = h1 > h2 ? h2 : h1;
        const double nsim = mx == 0 ? 0.0 : mx / (mx + mn);
        if (h3 != 0) { if (!(1.0 >= p->percentage_threshold)) out->total_high_similarity++; }
        else if (mx != 0 && !(nsim >= p->percentage_threshold)) out->total_high_similarity++;
        if (ctx->h_tag_nps[r] > 1) out->total_cross_two_block++;
        if (mx == 0 && h3 == 0) out->total_without_variant++;
        if (h1 == 0 && h2 == 0 && h3 != 0 && hp == 3) out->total_read_only_h3++;
        out->total_hp[hp]++;
        if (hp != 0) out->total_tag++; else out->total_untag++;
    }
    return LPS_OK;
}

// CNV mismatch filter of addEdge (PhasingGraph.cpp:783-791) from the clip map: the state machine over the clip positions
// (run twice, as in PhasingProcess.cpp:147-148) is host code; without an interval (the normal case) nothing is erased
static int host_cnv_filter(lps_ctx *ctx) {
    ctx->h_cnv_start.clear(); ctx->h_cnv_end.clear();
    lps_host_cnv_intervals(ctx->h_clip_pos, ctx->h_clip_front, ctx->h_clip_back, ctx->h_cnv_start, ctx->h_cnv_end);
    lps_host_cnv_intervals(ctx->h_clip_pos, ctx->h_clip_front, ctx->h_clip_back, ctx->h_cnv_start, ctx->h_cnv_end);
    ctx->have_erased = false;
    if (!ctx->h_cnv_start.empty()) {
        // the filter looks at which alignments survive the overlap filter
        const size_t n = (size_t)ctx->batch.n_reads;
        TRY(d2h(ctx, ctx->h_read_dead, ctx->d_read_dead.p, n));
        TRY(fetch_host_calls(ctx));
        std::vector<uint8_t> erased;
        TRY(lps_host_cnv_filter(ctx, erased));
        bool any = false;
        for (uint8_t e : erased) if (e) { any = true; break; }
        if (any) {
            TRY(h2d(ctx, ctx->d_call_erased, erased.data(), erased.size()));
            ctx->have_erased = true;
        }
    }
    return LPS_OK;
}

int lps_phase_build_edges(lps_ctx *ctx, const lps_phase_params *p, int want_host, lps_edges *out) {
    if (!ctx || !p) return LPS_E_ARG;
    if (!ctx->have_calls) return ctx->fail(LPS_E_STATE, "lps_phase_call_alleles must run first");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    WallTimer wt;
    // ---- filters at the head of addEdge (PhasingGraph.cpp:707-791): overlap filter on the device, CNV filter from the clip map ----
    WallTimer wf;
    TRY(lps_finish_clips(ctx));
    TRY(lps_launch_overlap_filter(ctx, p));
    TRY(host_cnv_filter(ctx));
    ctx->stats.ms_host_filters = wf.ms();
    // ---- device: merge by name, fan out, ordered fold ----
    cudaEventRecord(ctx->ev[2], st);
    TRY(lps_launch_build_edges(ctx, p, true));
    cudaEventRecord(ctx->ev[3], st);
    TRY(lps_fetch_graph_counts(ctx));
    ctx->stats.ms_build_edges = elapsed(ctx, 2, 3);
    const size_t nn = (size_t)ctx->n_nodes;
    TRY(d2h(ctx, ctx->h_node_var, ctx->d_node_var.p, nn));
    TRY(d2h(ctx, ctx->h_node_type, ctx->d_node_type.p, nn));
    ctx->h_weights.clear();
    if (want_host) TRY(d2h(ctx, ctx->h_weights, ctx->d_weights.p, nn * (size_t)ctx->window * 4));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    if (out) {
        memset(out, 0, sizeof(*out));
        out->n_nodes = ctx->n_nodes; out->window = ctx->window;
        out->node_var = ctx->h_node_var.data(); out->node_type = ctx->h_node_type.data();
        out->weights = want_host ? ctx->h_weights.data() : nullptr;
        out->n_contrib = ctx->n_contrib; out->n_contrib_far = ctx->n_contrib_far;
    }
    ctx->stats.ms_wall_build_edges = wt.ms();
    return LPS_OK;
}

// edgeConnectResult on the HOST over the device's vote bytes: kept for windows beyond the device sweep's 63 successors and as an
// A/B switch (LPS_HOST_SWEEP=1); lps_sweep_votes exposes the same code without any device
static int host_sweep(lps_ctx *ctx, const lps_phase_params *p) {
    cudaStream_t st = ctx->stream;
    TRY(lps_fetch_graph_counts(ctx));
    const size_t nn = (size_t)ctx->n_nodes, nv = (size_t)ctx->var.n;
    const size_t RS = (size_t)lps_vote_row_stride(ctx->window);
    LPS_CUDA(ctx, ctx->p_vote_info.reserve(nn * RS + 128));
    LPS_CUDA(ctx, ctx->p_last_link.reserve(nn + 16));
    TRY(d2h(ctx, ctx->h_node_var, ctx->d_node_var.p, nn));
    TRY(d2h(ctx, ctx->h_node_type, ctx->d_node_type.p, nn));
    if (nn) {
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_vote_info.p, ctx->d_vote_info.p, nn * RS, cudaMemcpyDeviceToHost, st));
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_last_link.p, ctx->d_last_link.p, nn, cudaMemcpyDeviceToHost, st));
    }
    ctx->stats.d2h_bytes += nn * RS + nn;
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    WallTimer ws;
    std::vector<int32_t> node_pos(nn), node_ps(nn);
    std::vector<int8_t> node_hap(nn);
    for (size_t k = 0; k < nn; k++) node_pos[k] = ctx->h_vpos[(size_t)ctx->h_node_var[k]];
    if (const char *dump = getenv("LPS_DUMP_SWEEP")) {          // inputs of the sweep, for tools/sweep_bench (tuning the host code offline)
        if (FILE *f = fopen(dump, "wb")) {
            const int32_t hdr[4] = {(int32_t)nn, ctx->window, p->distance, (int32_t)RS};
            fwrite(hdr, 4, 4, f); fwrite(node_pos.data(), 4, nn, f); fwrite(ctx->h_node_type.data(), 1, nn, f);
            fwrite(ctx->p_vote_info.p, 1, nn * RS, f); fwrite(ctx->p_last_link.p, 1, nn, f);
            fclose(f);
        }
    }
    ctx->stats.sweep_simd = lps_host_sweep(p, ctx->n_nodes, ctx->window, node_pos.data(), ctx->h_node_type.data(), ctx->p_vote_info.p,
                                           ctx->p_last_link.p, node_ps.data(), node_hap.data());
    ctx->h_ps_sweep.assign(nv, 0);
    ctx->h_hap_sweep.assign(nv, -1);
    for (size_t k = 0; k < nn; k++) {
        ctx->h_ps_sweep[(size_t)ctx->h_node_var[k]] = node_ps[k];
        ctx->h_hap_sweep[(size_t)ctx->h_node_var[k]] = node_hap[k];
    }
    ctx->stats.ms_host_sweep = ws.ms();
    TRY(h2d(ctx, ctx->d_ps, ctx->h_ps_sweep.data(), nv));
    TRY(h2d(ctx, ctx->d_hap_ref, ctx->h_hap_sweep.data(), nv));
    return LPS_OK;
}

static bool use_host_sweep(const lps_ctx *ctx) {
    const char *env = getenv("LPS_HOST_SWEEP");
    return ctx->window > 63 || (env && env[0] == '1');
}

int lps_phase_solve(lps_ctx *ctx, const lps_phase_params *p, lps_phase_result *out) {
    if (!ctx || !p) return LPS_E_ARG;
    if (!ctx->have_graph) return ctx->fail(LPS_E_STATE, "lps_phase_build_edges must run first");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const size_t nv = (size_t)ctx->var.n, n = (size_t)ctx->batch.n_reads;
    WallTimer wt;
    ctx->stats.ms_host_sweep = 0.f;
    // ---- sweep (edgeConnectResult) over the one-byte votes computed by the fold epilogue ----
    if (use_host_sweep(ctx)) TRY(host_sweep(ctx, p));
    else {
        TRY(lps_launch_sweep(ctx, p, ctx->var.n));
        TRY(d2h(ctx, ctx->h_ps_sweep, ctx->d_ps.p, nv));
        TRY(d2h(ctx, ctx->h_hap_sweep, ctx->d_hap_ref.p, nv));
        LPS_CUDA(ctx, cudaStreamSynchronize(st));
    }
    // ---- device: read correction ----
    cudaEventRecord(ctx->ev[2], st);
    TRY(lps_launch_read_correction(ctx, p));
    cudaEventRecord(ctx->ev[3], st);
    TRY(d2h(ctx, ctx->h_ps, ctx->d_ps.p, nv));
    TRY(d2h(ctx, ctx->h_hap_ref, ctx->d_hap_ref.p, nv));
    TRY(d2h(ctx, ctx->h_read_hp, ctx->d_read_hp.p, n));
    TRY(d2h(ctx, ctx->h_hp_counts, ctx->d_hp_counts.p, nv * 4));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->stats.ms_read_correction = elapsed(ctx, 2, 3);
    if (out) {
        memset(out, 0, sizeof(*out));
        out->n_variants = ctx->var.n; out->ps = ctx->h_ps.data(); out->hap_ref = ctx->h_hap_ref.data();
        out->n_reads = ctx->batch.n_reads; out->read_hp = ctx->h_read_hp.data(); out->hp_counts = ctx->h_hp_counts.data();
        out->ps_sweep = ctx->h_ps_sweep.data(); out->hap_ref_sweep = ctx->h_hap_sweep.data();
    }
    ctx->stats.ms_wall_solve = wt.ms();
    return LPS_OK;
}

int lps_sweep_votes(const lps_phase_params *p, int32_t n_nodes, int32_t window, const int32_t *node_pos, const uint8_t *node_type,
                    const uint8_t *votes, int32_t *node_ps, int8_t *node_hap_ref) {
    if (!p || n_nodes < 0 || window < 1 || window > 127 || (n_nodes && (!node_pos || !node_type || !votes || !node_ps || !node_hap_ref)))
        return LPS_E_ARG;
    // plain [n_nodes][window] rows -> the block-shifted rows k_fold_edges writes (host_phase.cpp), plus last_link
    const size_t N = (size_t)n_nodes, W = (size_t)window, RS = (size_t)lps_vote_row_stride(window);
    std::vector<uint8_t> buf(N * RS + 64, 0);
    uint8_t *rows = buf.data() + ((16 - ((uintptr_t)buf.data() & 15)) & 15);
    std::vector<int8_t> last(N, -1);
    for (size_t k = 0; k < N; k++) {
        const size_t shift = (k + 1) & 15;
        for (size_t d = 0; d < W && k + 1 + d < N; d++) {
            const uint8_t info = votes[k * W + d];
            rows[k * RS + shift + d] = info;
            if (info & 3u) last[k] = (int8_t)d;
        }
    }
    return lps_host_sweep(p, n_nodes, window, node_pos, node_type, rows, last.data(), node_ps, node_hap_ref);
}

// The body of the contig loop (PhasingProcess.cpp:113-173) as ONE asynchronous pipeline: after the allele-calling kernel has
// reported its counters (the one wait in the middle: they size everything downstream) every stage is only enqueued - overlap
// filter, graph construction and fold, sweep, read correction, the copies of the results into pinned memory - and the host waits
// once, at the end.  While the device works the host turns the clip map into CNV intervals; two rare conditions need host code in
// the middle of the pipeline (a CNV interval: calls have to be erased before the graph is built; a merged read of > 16 calls
// with tied positions: libstdc++'s std::sort order has to be replayed) - then the result of the fast pipeline is discarded and
// the contig is redone stage by stage (lps_phase_build_edges / lps_phase_solve), which handles both.
int lps_phase_contig(lps_ctx *ctx, const lps_phase_params *p, lps_phase_result *out) {
    if (!ctx || !p) return LPS_E_ARG;
    if (!ctx->have_variants || !ctx->have_batch) return ctx->fail(LPS_E_STATE, "variants and a read batch must be set first");
    const char *staged = getenv("LPS_STAGED");
    if (use_host_sweep(ctx) || p->connect_adjacent > 63 || (staged && staged[0] == '1')) {
        TRY(lps_phase_call_alleles(ctx, p, 0, nullptr));
        TRY(lps_phase_build_edges(ctx, p, 0, nullptr));
        return lps_phase_solve(ctx, p, out);
    }
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const size_t nv = (size_t)ctx->var.n, n = (size_t)ctx->batch.n_reads;
    WallTimer wt;
    cudaEventRecord(ctx->ev[2], st);
    TRY(lps_launch_call_alleles(ctx, p, nullptr, 0, -1, true));
    cudaEventRecord(ctx->ev[3], st);
    ctx->stats.ms_wall_call_alleles = wt.ms();
    WallTimer wb;
    ctx->have_erased = false;
    TRY(lps_launch_overlap_filter(ctx, p));
    cudaEventRecord(ctx->ev[4], st);
    TRY(lps_launch_build_edges(ctx, p, false));
    cudaEventRecord(ctx->ev[5], st);
    TRY(lps_launch_sweep(ctx, p, ctx->var.n));
    LPS_CUDA(ctx, ctx->p_ps_sweep.reserve(nv + 1)); LPS_CUDA(ctx, ctx->p_hap_sweep.reserve(nv + 1));
    LPS_CUDA(ctx, ctx->p_ps.reserve(nv + 1)); LPS_CUDA(ctx, ctx->p_hap.reserve(nv + 1));
    LPS_CUDA(ctx, ctx->p_read_hp.reserve(n + 1)); LPS_CUDA(ctx, ctx->p_hp_counts.reserve(4 * nv + 4));
    LPS_CUDA(ctx, ctx->p_status.reserve(8));
    if (nv) {
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_ps_sweep.p, ctx->d_ps.p, 4 * nv, cudaMemcpyDeviceToHost, st));
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_hap_sweep.p, ctx->d_hap_ref.p, nv, cudaMemcpyDeviceToHost, st));
    }
    cudaEventRecord(ctx->ev[6], st);
    TRY(lps_launch_read_correction(ctx, p));
    cudaEventRecord(ctx->ev[7], st);
    if (nv) {
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_ps.p, ctx->d_ps.p, 4 * nv, cudaMemcpyDeviceToHost, st));
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_hap.p, ctx->d_hap_ref.p, nv, cudaMemcpyDeviceToHost, st));
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_hp_counts.p, ctx->d_hp_counts.p, 16 * nv, cudaMemcpyDeviceToHost, st));
    }
    if (n) LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_read_hp.p, ctx->d_read_hp.p, n, cudaMemcpyDeviceToHost, st));
    LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_status.p, ctx->d_n_tie.p, 4, cudaMemcpyDeviceToHost, st));
    LPS_CUDA(ctx, cudaMemcpyAsync(ctx->p_status.p + 1, ctx->d_sweep_ok.p, 4, cudaMemcpyDeviceToHost, st));
    ctx->stats.d2h_bytes += 26 * nv + n + 8;
    // ---- host, while the device works: clip map -> CNV intervals ----
    WallTimer wf;
    TRY(lps_finish_clips(ctx));
    ctx->h_cnv_start.clear(); ctx->h_cnv_end.clear();
    lps_host_cnv_intervals(ctx->h_clip_pos, ctx->h_clip_front, ctx->h_clip_back, ctx->h_cnv_start, ctx->h_cnv_end);
    ctx->stats.ms_host_filters = wf.ms();
    ctx->stats.ms_host_sweep = 0.f;
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->stats.ms_wall_build_edges = wb.ms();
    ctx->stats.ms_wall_solve = 0.f;
    ctx->stats.ms_call_alleles = elapsed(ctx, 2, 3);
    ctx->stats.ms_build_edges = elapsed(ctx, 4, 5);
    ctx->stats.ms_sweep = elapsed(ctx, 5, 6);
    ctx->stats.ms_read_correction = elapsed(ctx, 6, 7);
    cudaEventElapsedTime(&ctx->stats.ms_kernel_fold_edges, ctx->kev[2], ctx->kev[3]);
    ctx->stats.sweep_fallbacks += ctx->p_status.p[1] ? 0 : 1;
    if (getenv("LPS_DEBUG_SLOW") && (!ctx->h_cnv_start.empty() || ctx->p_status.p[0] != 0))
        fprintf(stderr, "lps_phase_contig: slow path (cnv intervals %zu, tie groups for the host %d)\n", ctx->h_cnv_start.size(), ctx->p_status.p[0]);
    if (!ctx->h_cnv_start.empty() || ctx->p_status.p[0] != 0) {
        // rare: host code is needed between the kernels; redo the contig stage by stage (the calls are still on the device)
        ctx->stats.slow_path_contigs++;
        TRY(lps_phase_build_edges(ctx, p, 0, nullptr));
        return lps_phase_solve(ctx, p, out);
    }
    if (out) {
        memset(out, 0, sizeof(*out));
        out->n_variants = ctx->var.n; out->ps = ctx->p_ps.p; out->hap_ref = ctx->p_hap.p;
        out->n_reads = ctx->batch.n_reads; out->read_hp = ctx->p_read_hp.p; out->hp_counts = ctx->p_hp_counts.p;
        out->ps_sweep = ctx->p_ps_sweep.p; out->hap_ref_sweep = ctx->p_hap_sweep.p;
    }
    return LPS_OK;
}

int lps_event_record(lps_ctx *ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 4) return LPS_E_ARG;
    cudaSetDevice(ctx->device);
    LPS_CUDA(ctx, cudaEventRecord(ctx->user_ev[slot], ctx->stream));
    return LPS_OK;
}

int lps_event_elapsed_ms(lps_ctx *ctx, int slot_a, int slot_b, float *ms) {
    if (!ctx || !ms || slot_a < 0 || slot_a >= 4 || slot_b < 0 || slot_b >= 4) return LPS_E_ARG;
    cudaSetDevice(ctx->device);
    LPS_CUDA(ctx, cudaEventSynchronize(ctx->user_ev[slot_b]));
    LPS_CUDA(ctx, cudaEventElapsedTime(ms, ctx->user_ev[slot_a], ctx->user_ev[slot_b]));
    return LPS_OK;
}

int lps_get_stats(lps_ctx *ctx, lps_stats *out) {
    if (!ctx || !out) return LPS_E_ARG;
    *out = ctx->stats;
    return LPS_OK;
}

}  // extern "C"
