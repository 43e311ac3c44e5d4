// k_bgzf_deflate.cu — BGZF block deflation on the device (second half: the writer side of the tagging passes).
// Replaces bgzf_write -> bgzf_flush -> deflate_block -> bgzf_compress (htslib/bgzf.c:553-612, 697, 1440-1500; zlib deflate with
// windowBits -15) for a batch of blocks: the tagged-BAM writer is two thirds of the reference's `haplotag` wall time.
//
// What a BAM stream is made of decides the encoder.  Bases (4 bits) and base qualities (8 bits) are ~95 % of a long-read record and
// hold almost no repeated strings; zlib's gain on them is the entropy coding of the quality alphabet, not LZ77.  Measured on the
// uncompressed stream of this repository's ONT-like BAMs, per 65 280-byte block: zlib level 6 -> 0.617 of the input, level 1 ->
// 0.656, Huffman coding alone with a per-block (dynamic) code -> 0.631, LZ77 with the fixed code -> 0.822.  So the kernel writes ONE
// dynamic-Huffman deflate block of literals per BGZF member (RFC 1951 3.2.7), or a stored block when that is not smaller: every
// inflater accepts it, the records are identical to the reference's, the file is ~2 % larger than htslib's default level.
//
// No string matching means no chain through a hash table: a member is two streaming passes over its <= 65 280 input bytes
// (histogram + CRC32, then code emission) around a 257-symbol length-limited Huffman construction.  One THREAD per member: a
// 1 GB stream is 16 k independent members, enough threads to cover their own latencies; the per-thread tables (~6 KB) live in local
// memory (L1).  Members are written into fixed 65 312-byte slots, an exclusive scan of their sizes and a copy kernel make the
// stream contiguous.  The block encoder is one __host__ __device__ function: tests/test_bgzf_deflate.py runs it on the host against
// zlib's inflate (lps_bgzf_deflate_block_host) and the kernel's output is compared with it byte for byte on the GPU box.
#include <cub/cub.cuh>
#include "lps_ctx.cuh"

namespace {

constexpr uint32_t DEFL_MAX_IN = 0xff00u;        // BGZF_BLOCK_SIZE (htslib/bgzf.c:67): input bytes per member
constexpr uint32_t DEFL_SLOT = 65312u;           // 18 (header) + 5 (stored-block header) + 65 280 + 8 (trailer) = 65 311, rounded up to 16
constexpr int LIT_SYMS = 257;                    // literals 0..255 and end-of-block (256); no length symbols are ever used
constexpr int CL_SYMS = 19;                      // code-length alphabet (RFC 1951 3.2.7)

// ---- CRC-32 (ISO 3309, reflected 0xEDB88320), one table of 256 words ----
__host__ __device__ inline uint32_t crc_entry(uint32_t n) {
    uint32_t c = n;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    return c;
}

// ---- lengths of a length-limited prefix code ----
// freq[0..N): symbol counts.  len[0..N) receives the code length of every symbol (0 = unused).  At least two symbols get a code
// (the format needs a complete code; zlib forces the same, trees.c build_tree).  Plain Huffman by the two-queue merge over the
// symbols sorted by (count, symbol); when the tree is deeper than LIMIT the number of codes per length is repaired until the Kraft
// sum is exactly 1 again (one code leaves the deepest level, one shallower code moves one level down to make room for two), and
// the lengths are handed out again in count order, the rarest symbols taking the longest codes.
template <int N, int LIMIT>
__host__ __device__ void build_lengths(const uint32_t *freq, uint8_t *len) {
    uint16_t order[N + 1];
    uint32_t w[2 * N + 2];
    uint16_t parent[2 * N + 2];
    uint8_t depth[2 * N + 2];
    int used = 0;
    for (int s = 0; s < N; s++) {
        len[s] = 0;
        if (freq[s]) order[used++] = (uint16_t)s;
    }
    if (used == 0) { order[used++] = 0; }
    if (used == 1) { order[used++] = (uint16_t)(order[0] == 0 ? 1 : 0); }       // a second symbol of count 0 completes the code
    // insertion sort by (count, symbol); the symbols come in ascending order, so equal counts keep it
    for (int i = 1; i < used; i++) {
        const uint16_t s = order[i];
        const uint32_t f = freq[s];
        int j = i - 1;
        while (j >= 0 && (freq[order[j]] > f || (freq[order[j]] == f && order[j] > s))) { order[j + 1] = order[j]; j--; }
        order[j + 1] = s;
    }
    for (int i = 0; i < used; i++) w[i] = freq[order[i]];
    int li = 0, ii = used, ni = used;                          // next leaf, next unmerged internal node, next internal node to create
    for (int k = 0; k + 1 < used; k++) {
        int a, b;
        if (li < used && (ii >= ni || w[li] <= w[ii])) a = li++; else a = ii++;
        if (li < used && (ii >= ni || w[li] <= w[ii])) b = li++; else b = ii++;
        w[ni] = w[a] + w[b];
        parent[a] = (uint16_t)ni; parent[b] = (uint16_t)ni;
        ni++;
    }
    const int root = ni - 1;
    depth[root] = 0;
    for (int v = root - 1; v >= 0; v--) depth[v] = (uint8_t)(depth[parent[v]] + 1);
    int num[64];
    for (int i = 0; i < 64; i++) num[i] = 0;
    for (int i = 0; i < used; i++) num[depth[i] < 63 ? depth[i] : 63]++;
    bool deep = false;
    for (int i = LIMIT + 1; i < 64; i++) if (num[i]) deep = true;
    if (deep) {
        for (int i = LIMIT + 1; i < 64; i++) { num[LIMIT] += num[i]; num[i] = 0; }
        uint32_t total = 0;
        for (int i = LIMIT; i >= 1; i--) total += (uint32_t)num[i] << (LIMIT - i);
        while (total != (1u << LIMIT)) {
            num[LIMIT]--;
            for (int i = LIMIT - 1; i >= 1; i--)
                if (num[i]) { num[i]--; num[i + 1] += 2; break; }
            total--;
        }
    }
    int at = 0;
    for (int l = LIMIT; l >= 1; l--)
        for (int c = 0; c < num[l]; c++) len[order[at++]] = (uint8_t)l;
}

// canonical codes (RFC 1951 3.2.2) of the lengths, bit-reversed for the LSB-first packer: out[s] = reversed code | length << 16
template <int N>
__host__ __device__ void assign_codes(const uint8_t *len, uint32_t *out) {
    uint32_t count[16], next[16];
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int s = 0; s < N; s++) count[len[s]]++;
    count[0] = 0;
    uint32_t code = 0;
    next[0] = 0;
    for (int b = 1; b < 16; b++) { code = (code + count[b - 1]) << 1; next[b] = code; }
    for (int s = 0; s < N; s++) {
        const uint32_t l = len[s];
        if (!l) { out[s] = 0; continue; }
        uint32_t c = next[l]++, r = 0;
        for (uint32_t k = 0; k < l; k++) { r = (r << 1) | (c & 1u); c >>= 1; }
        out[s] = r | (l << 16);
    }
}

struct BitOut {
    uint8_t *p;
    uint32_t at;
    uint64_t acc;
    int n;
    __host__ __device__ void put(uint32_t v, int bits) {
        acc |= (uint64_t)v << n;
        n += bits;
        if (n >= 32) {
            p[at] = (uint8_t)acc; p[at + 1] = (uint8_t)(acc >> 8); p[at + 2] = (uint8_t)(acc >> 16); p[at + 3] = (uint8_t)(acc >> 24);
            at += 4; acc >>= 32; n -= 32;
        }
    }
    __host__ __device__ void finish() {                        // pads the last byte with zero bits
        while (n > 0) { p[at++] = (uint8_t)acc; acc >>= 8; n -= 8; }
        n = 0;
    }
};

__host__ __device__ inline void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
__host__ __device__ inline void put32(uint8_t *p, uint32_t v) { put16(p, v); put16(p + 2, v >> 16); }

// One BGZF member (RFC 1952 member with htslib's BC extra field, bgzf.c:85-100) for n <= 65 280 input bytes; `out` must hold
// DEFL_SLOT bytes.  Returns the member's size.  crc_table: 256 words.
__host__ __device__ uint32_t deflate_member(const uint8_t *__restrict__ in, uint32_t n, uint8_t *__restrict__ out, const uint32_t *__restrict__ crc_table) {
    // ---- pass 1: histogram and CRC-32 ----
    uint32_t freq[LIT_SYMS];
    for (int s = 0; s < LIT_SYMS; s++) freq[s] = 0;
    uint32_t crc = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t b = in[i];
        freq[b]++;
        crc = (crc >> 8) ^ crc_table[(crc ^ b) & 0xFFu];
    }
    crc = ~crc;
    freq[256] = 1;
    // ---- the two codes ----
    uint8_t lit_len[LIT_SYMS];
    build_lengths<LIT_SYMS, 15>(freq, lit_len);
    uint32_t cl_freq[CL_SYMS];
    for (int s = 0; s < CL_SYMS; s++) cl_freq[s] = 0;
    for (int s = 0; s < LIT_SYMS; s++) cl_freq[lit_len[s]]++;
    cl_freq[1] += 2;                                           // two distance codes of one bit, as zlib sends for a block without matches
    uint8_t cl_len[CL_SYMS];
    build_lengths<CL_SYMS, 7>(cl_freq, cl_len);
    // ---- size of the dynamic block in bits ----
    uint64_t bits = 3 + 5 + 5 + 4 + 3 * CL_SYMS;
    for (int s = 0; s < CL_SYMS; s++) bits += (uint64_t)cl_freq[s] * cl_len[s];
    for (int s = 0; s < LIT_SYMS; s++) bits += (uint64_t)freq[s] * lit_len[s];
    const uint32_t dyn_bytes = (uint32_t)((bits + 7) >> 3);
    uint8_t *body = out + 18;
    uint32_t body_bytes;
    if (dyn_bytes < n + 5u) {
        uint32_t lit_code[LIT_SYMS], cl_code[CL_SYMS];
        assign_codes<LIT_SYMS>(lit_len, lit_code);
        assign_codes<CL_SYMS>(cl_len, cl_code);
        BitOut bo{body, 0u, 0ull, 0};
        bo.put(1u, 1);                                         // BFINAL
        bo.put(2u, 2);                                         // BTYPE = 10: dynamic Huffman codes
        bo.put((uint32_t)(LIT_SYMS - 257), 5);                 // HLIT
        bo.put(1u, 5);                                         // HDIST: 2 distance codes
        bo.put((uint32_t)(CL_SYMS - 4), 4);                    // HCLEN: all 19 lengths of the code-length code follow
        const uint8_t perm[CL_SYMS] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (int i = 0; i < CL_SYMS; i++) bo.put(cl_len[perm[i]], 3);
        for (int s = 0; s < LIT_SYMS; s++) { const uint32_t c = cl_code[lit_len[s]]; bo.put(c & 0xFFFFu, (int)(c >> 16)); }
        for (int d = 0; d < 2; d++) { const uint32_t c = cl_code[1]; bo.put(c & 0xFFFFu, (int)(c >> 16)); }
        // ---- pass 2: the literals ----
        for (uint32_t i = 0; i < n; i++) { const uint32_t c = lit_code[in[i]]; bo.put(c & 0xFFFFu, (int)(c >> 16)); }
        { const uint32_t c = lit_code[256]; bo.put(c & 0xFFFFu, (int)(c >> 16)); }
        bo.finish();
        body_bytes = bo.at;
    } else {
        body[0] = 0x01;                                        // BFINAL, BTYPE = 00, padding: a stored block (incompressible bytes)
        put16(body + 1, n);
        put16(body + 3, ~n & 0xFFFFu);
        for (uint32_t i = 0; i < n; i++) body[5 + i] = in[i];
        body_bytes = 5 + n;
    }
    const uint32_t total = 18 + body_bytes + 8;
    const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    for (int i = 0; i < 16; i++) out[i] = hdr[i];
    put16(out + 16, total - 1);                                // BSIZE
    put32(body + body_bytes, crc);
    put32(body + body_bytes + 4, n);                           // ISIZE
    return total;
}

__global__ void __launch_bounds__(64) k_bgzf_deflate(uint32_t n_blocks, const uint8_t *__restrict__ in, uint64_t in_len, uint32_t block_bytes,
                                                     uint8_t *__restrict__ slots, uint64_t *__restrict__ sizes) {
    __shared__ uint32_t crc_table[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) crc_table[i] = crc_entry((uint32_t)i);
    __syncthreads();
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_blocks) return;
    if (k == n_blocks) { sizes[k] = 0; return; }               // the scan's last element: the total
    const uint64_t at = (uint64_t)k * block_bytes;
    const uint32_t n = (uint32_t)(in_len - at < block_bytes ? in_len - at : block_bytes);
    sizes[k] = deflate_member(in + at, n, slots + (size_t)k * DEFL_SLOT, crc_table);
}

// slots -> contiguous stream; one CTA per member
__global__ void __launch_bounds__(256) k_bgzf_compact(const uint8_t *__restrict__ slots, const uint64_t *__restrict__ sizes, const uint64_t *__restrict__ offsets,
                                                      uint8_t *__restrict__ out) {
    const uint32_t k = blockIdx.x;
    const uint32_t n = (uint32_t)sizes[k];
    const uint8_t *src = slots + (size_t)k * DEFL_SLOT;
    uint8_t *dst = out + offsets[k];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

const uint32_t *host_crc_table() {
    static uint32_t t[256];
    static bool ready = false;
    if (!ready) { for (uint32_t i = 0; i < 256; i++) t[i] = crc_entry(i); ready = true; }
    return t;
}

}  // namespace

extern "C" {

uint64_t lps_bgzf_deflate_bound(uint64_t in_len, uint32_t block_bytes) {
    if (block_bytes == 0 || block_bytes > DEFL_MAX_IN) return 0;
    const uint64_t n_blocks = (in_len + block_bytes - 1) / block_bytes;
    return in_len + n_blocks * (18 + 5 + 8);
}

int lps_bgzf_deflate_block_host(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t out_cap, uint32_t *out_len) {
    if ((n && !in) || !out || !out_len || n > DEFL_MAX_IN || out_cap < DEFL_SLOT) return LPS_E_ARG;
    static const uint32_t *table = host_crc_table();
    *out_len = deflate_member(in, n, out, table);
    return LPS_OK;
}

int lps_bgzf_deflate(lps_ctx *ctx, const uint8_t *in, uint64_t in_len, uint32_t block_bytes, uint8_t *out, uint64_t out_cap, uint64_t *out_len) {
    if (!ctx || !out_len) return LPS_E_ARG;
    *out_len = 0;
    if (block_bytes == 0 || block_bytes > DEFL_MAX_IN) return ctx->fail(LPS_E_ARG, "a BGZF member holds 1 .. 65280 input bytes");
    if ((in_len && !in) || (in_len && !out)) return ctx->fail(LPS_E_ARG, "null buffer");
    if (out_cap < lps_bgzf_deflate_bound(in_len, block_bytes)) return ctx->fail(LPS_E_ARG, "out_cap is below lps_bgzf_deflate_bound");
    if (in_len == 0) return LPS_OK;
    const uint64_t n_blocks64 = (in_len + block_bytes - 1) / block_bytes;
    if (n_blocks64 > (1ull << 24)) return ctx->fail(LPS_E_ARG, "too many BGZF members for one call (1 TB of input)");
    const uint32_t n_blocks = (uint32_t)n_blocks64;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    LPS_CUDA(ctx, ctx->d_bgzf_in.reserve((size_t)in_len + 16));
    LPS_CUDA(ctx, ctx->d_defl_slots.reserve((size_t)n_blocks * DEFL_SLOT));
    LPS_CUDA(ctx, ctx->d_defl_sizes.reserve((size_t)n_blocks + 1));
    LPS_CUDA(ctx, ctx->d_defl_off.reserve((size_t)n_blocks + 1));
    LPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_bgzf_in.p, in, (size_t)in_len, cudaMemcpyHostToDevice, st));
    cudaEventRecord(ctx->kev[4], st);
    k_bgzf_deflate<<<(n_blocks + 1 + 63) / 64, 64, 0, st>>>(n_blocks, ctx->d_bgzf_in.p, in_len, block_bytes, ctx->d_defl_slots.p, ctx->d_defl_sizes.p);
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->d_defl_sizes.p, ctx->d_defl_off.p, (int)n_blocks + 1, st);
    LPS_CUDA(ctx, ctx->d_cub_tmp.reserve(tmp));
    cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, tmp, ctx->d_defl_sizes.p, ctx->d_defl_off.p, (int)n_blocks + 1, st);
    uint64_t total = 0;
    LPS_CUDA(ctx, cudaMemcpyAsync(&total, ctx->d_defl_off.p + n_blocks, sizeof(total), cudaMemcpyDeviceToHost, st));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    LPS_CUDA(ctx, cudaGetLastError());
    if (total > out_cap) return ctx->fail(LPS_E_CUDA, "deflated stream exceeds its bound");       // cannot happen: every member is <= its input + 31
    LPS_CUDA(ctx, ctx->d_defl_out.reserve((size_t)total + 16));
    k_bgzf_compact<<<n_blocks, 256, 0, st>>>(ctx->d_defl_slots.p, ctx->d_defl_sizes.p, ctx->d_defl_off.p, ctx->d_defl_out.p);
    cudaEventRecord(ctx->kev[5], st);
    LPS_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_defl_out.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    LPS_CUDA(ctx, cudaStreamSynchronize(st));
    LPS_CUDA(ctx, cudaGetLastError());
    cudaEventElapsedTime(&ctx->stats.ms_kernel_bgzf, ctx->kev[4], ctx->kev[5]);
    ctx->stats.kernel_launches += 3;
    ctx->stats.h2d_bytes += in_len;
    ctx->stats.d2h_bytes += total + sizeof(total);
    *out_len = total;
    return LPS_OK;
}

}  // extern "C"
