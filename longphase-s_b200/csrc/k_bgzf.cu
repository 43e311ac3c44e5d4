// k_bgzf.cu — BGZF block inflation on the device (77 % of `phase` wall time in the reference is
// bgzf_read_block -> inflate_block -> zlib inflate, htslib/bgzf.c:988-1200, 792-812, 744-785).
//
// One warp per BGZF block (<= 64 KiB of output, RFC 1951 raw deflate inside an RFC 1952 member).  Huffman decoding is a chain —
// the position of a code is known only when the previous one has been decoded — so the 32 lanes decode the SAME bit stream in
// lockstep, redundantly: every lane holds the same bit buffer, reads the same table entries (shared-memory broadcasts), takes the
// same branches.  Nothing is ever exchanged between lanes and there is no divergence; the lanes differ only where bytes move:
//   * a literal is stored by lane 0;
//   * a match (length 3..258) is copied by all lanes at once from the warp's output window;
//   * the window is flushed to global memory 32 consecutive bytes per store instruction;
//   * the compressed input is staged 256 bytes at a time by coalesced loads issued one refill ahead.
// The window is a 2 KiB ring in shared memory; matches that reach further back (up to 32 KiB) read the already flushed bytes
// from global memory (L2).  The kernel is bound by the latency of the decode chain, so what matters most is how many warps an SM
// holds, and shared memory per warp decides that.  1 GB of BAM-like bytes on a B200, ring size -> warps per SM -> time:
// 16 KiB -> 11 -> 68.3 ms, 8 KiB -> 17 -> 44.4 ms, 4 KiB -> 26 -> 36.5 ms, 2 KiB -> 32 (the CTA limit) -> 32.3 ms; two or four
// warps per CTA with a 1 KiB ring (up to 42 warps per SM) were slower again (38 ms), and so was a 9-bit literal table.
// Tables: a 10-bit (literal/length) and an 8-bit (distance) direct lookup with the canonical count/symbol arrays behind them for
// longer codes, rebuilt per deflate block.
#include "lps_ctx.cuh"

#ifndef LPS_BGZF_LIT_FAST
#define LPS_BGZF_LIT_FAST 10        // bits of the direct literal/length lookup
#endif
#ifndef LPS_BGZF_CTA_WARPS
#define LPS_BGZF_CTA_WARPS 1        // BGZF blocks (= warps) per CTA; more than 32 resident warps per SM need more than one
#endif
#ifndef LPS_BGZF_RING
#define LPS_BGZF_RING 2048          // bytes of output window per warp in shared memory (power of two >= 1024)
#endif

namespace {

constexpr int RING = LPS_BGZF_RING, RMASK = RING - 1, FLUSH = RING / 4;
constexpr int CTA_WARPS = LPS_BGZF_CTA_WARPS;
constexpr int IN_WORDS = 64;                       // staged input, 32-bit words
constexpr int LIT_FAST = LPS_BGZF_LIT_FAST, DIST_FAST = 8, CL_FAST = 7;
constexpr int NEAR_MAX = RING - 512;               // matches up to this distance are served by the ring

enum { BGZF_OK = 0, BGZF_BAD_BLOCK_TYPE = 1, BGZF_BAD_STORED = 2, BGZF_BAD_CODE = 3, BGZF_BAD_DISTANCE = 4, BGZF_OVERRUN = 5,
       BGZF_BAD_LENGTHS = 6, BGZF_SIZE_MISMATCH = 7, BGZF_TRUNCATED = 8 };

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                         4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct WarpSmem {
    uint8_t ring[RING];
    uint32_t in[IN_WORDS];
    uint16_t lit_fast[1 << LIT_FAST];              // (symbol << 4) | code length, 0 = not a short code
    uint16_t dist_fast[1 << DIST_FAST];
    uint16_t cl_fast[1 << CL_FAST];
    uint16_t lit_sym[288], dist_sym[32], cl_sym[20];   // symbols ordered by (code length, symbol): canonical decoding
    uint16_t lit_count[16], dist_count[16], cl_count[16];
    uint8_t lens[320];                             // code lengths: literal/length alphabet, then distance alphabet
};

// the bit reader: identical in every lane
struct Bits {
    uint64_t buf = 0;
    int cnt = 0;
    uint32_t next = 0;        // word that follows the ones already in buf
    uint32_t widx = 0;        // index of `next` in the stream of aligned words
    uint32_t loaded_end = 0;  // words [loaded_end - IN_WORDS, loaded_end) are staged in shared memory
    uint32_t total = 0;       // words of the stream
    uint32_t pre0 = 0, pre1 = 0;   // this lane's two words of the NEXT staging round, already on their way from global memory
    const uint32_t *src = nullptr;
};

__device__ __forceinline__ void stage_input(Bits &b, WarpSmem &s, int lane) {
    // publish the words fetched during the previous round, then start fetching the round after
    __syncwarp();
    s.in[lane] = b.pre0;
    s.in[lane + 32] = b.pre1;
    b.loaded_end += IN_WORDS;
    const uint32_t w0 = b.loaded_end + lane, w1 = w0 + 32;
    b.pre0 = w0 < b.total ? __ldg(b.src + w0) : 0u;
    b.pre1 = w1 < b.total ? __ldg(b.src + w1) : 0u;
    __syncwarp();
}

__device__ __forceinline__ void fetch_next(Bits &b, WarpSmem &s, int lane) {
    b.widx++;
    if (b.widx == b.loaded_end) stage_input(b, s, lane);
    b.next = s.in[b.widx & (IN_WORDS - 1)];
}

// at least 33 bits in the buffer afterwards
__device__ __forceinline__ void ensure(Bits &b, WarpSmem &s, int lane) {
    if (b.cnt <= 32) {
        b.buf |= (uint64_t)b.next << b.cnt;
        b.cnt += 32;
        fetch_next(b, s, lane);
    }
}
__device__ __forceinline__ uint32_t take(Bits &b, int n) {
    const uint32_t v = (uint32_t)b.buf & ((1u << n) - 1u);
    b.buf >>= n;
    b.cnt -= n;
    return v;
}

// Canonical Huffman tables from code lengths (RFC 1951 3.2.2).  Serial parts run in lane 0, table filling in all lanes.
// Returns false for an over-subscribed set of lengths.
__device__ bool build_tables(const uint8_t *lens, int n, uint16_t *count, uint16_t *sym, uint16_t *fast, int fast_bits, int lane) {
    __syncwarp();
    for (int i = lane; i < (1 << fast_bits); i += 32) fast[i] = 0;
    if (lane < 16) count[lane] = 0;
    __syncwarp();
    if (lane == 0)
        for (int i = 0; i < n; i++) count[lens[i]]++;
    __syncwarp();
    int offs[16], code_of[16];
    int left = 1, code = 0, o = 0;
    bool ok = true;
    offs[0] = 0; code_of[0] = 0;
    for (int l = 1; l < 16; l++) {
        const int c = count[l];
        left = (left << 1) - c;
        if (left < 0) ok = false;
        offs[l] = o; o += c;
        code = (code + (l > 1 ? (int)count[l - 1] : 0)) << 1;
        code_of[l] = code;
    }
    if (!ok) return false;
    // every lane walks the symbols (the running code of each length is part of the walk); the replicated entries of a short code
    // are written by the lanes in parallel
    for (int i = 0; i < n; i++) {
        const int l = lens[i];
        if (!l) continue;
        const int c = code_of[l]++;
        if (lane == 0) sym[offs[l]] = (uint16_t)i;
        offs[l]++;
        if (l <= fast_bits) {
            const uint32_t rev = __brev((uint32_t)c) >> (32 - l);
            const uint16_t e = (uint16_t)((i << 4) | l);
            for (int k = lane; k < (1 << (fast_bits - l)); k += 32) fast[rev | ((uint32_t)k << l)] = e;
        }
    }
    __syncwarp();
    return true;
}

// one symbol; -1 = invalid code.  Needs 15 valid bits in the buffer.
__device__ __forceinline__ int decode_symbol(Bits &b, const uint16_t *fast, int fast_bits, const uint16_t *count, const uint16_t *sym) {
    const uint32_t e = fast[(uint32_t)b.buf & ((1u << fast_bits) - 1u)];
    if (e) {
        b.buf >>= (e & 15u);
        b.cnt -= (int)(e & 15u);
        return (int)(e >> 4);
    }
    // longer than the direct table: canonical decoding one bit at a time (the first bit read is the most significant of the code)
    int code = 0, first = 0, index = 0;
    uint64_t bits = b.buf;
    for (int l = 1; l < 16; l++) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[l];
        if (code - c < first) {
            b.buf >>= l;
            b.cnt -= l;
            return sym[index + (code - first)];
        }
        index += c; first += c;
        first <<= 1; code <<= 1;
    }
    return -1;
}

__device__ __forceinline__ void flush(WarpSmem &s, uint8_t *__restrict__ out, uint32_t from, uint32_t to, int lane) {
    __syncwarp();
    for (uint32_t i = from + lane; i < to; i += 32) out[i] = s.ring[i & RMASK];
    __syncwarp();
}

// SPECULATE: literal runs decoded through a 32-offset parallel lookup (false: one table lookup per symbol; kept for A/B runs and
// as a second implementation in the parity tests, environment variable LPS_BGZF_SPECULATE=0).  Measured on 1 GB of BAM-like
// bytes (0.78 symbols per byte, 3.2 bytes per match): 44.3 ms against 49.5 ms.  Deferring the ring store of far matches until the
// next near match was tried too and lost (50.7 ms): only 20 % of the matches are far and draining costs every near match.
template <bool SPECULATE>
__global__ void __launch_bounds__(32 * CTA_WARPS) k_bgzf_inflate(uint32_t n_blocks, const uint8_t *__restrict__ data, const lps_bgzf_block *__restrict__ blocks,
                                                     uint8_t *__restrict__ out_all, uint8_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSmem &s = reinterpret_cast<WarpSmem *>(smem_raw)[threadIdx.x >> 5];     // warps of a CTA share nothing
    const int lane = threadIdx.x & 31;
    const uint32_t blk = blockIdx.x * CTA_WARPS + (threadIdx.x >> 5);
    if (blk >= n_blocks) return;
    const lps_bgzf_block B = blocks[blk];
    uint8_t *out = out_all + B.out_off;
    const uint32_t out_len = B.out_len;

    // ---- input: aligned 32-bit words; the stream starts `skip` bytes into the first one ----
    Bits b;
    {
        const uintptr_t addr = (uintptr_t)(data + B.comp_off);
        const uint32_t skip = (uint32_t)(addr & 3u);
        b.src = (const uint32_t *)(addr - skip);
        b.total = (skip + B.comp_len + 3u) / 4u;
        b.pre0 = (uint32_t)lane < b.total ? __ldg(b.src + lane) : 0u;
        b.pre1 = (uint32_t)lane + 32u < b.total ? __ldg(b.src + lane + 32) : 0u;
        stage_input(b, s, lane);                   // words 0..63 staged, 64..127 in flight
        b.next = s.in[0];
        b.widx = 0;
        ensure(b, s, lane);
        take(b, 8 * (int)skip);
    }

    uint32_t p = 0, flushed = 0;
    int err = BGZF_OK;
    bool last = false;
    while (!last && err == BGZF_OK) {
        ensure(b, s, lane);
        last = take(b, 1);
        const uint32_t type = take(b, 2);
        if (type == 0) {
            // ---- stored ----
            take(b, b.cnt & 7);
            ensure(b, s, lane);
            const uint32_t len = take(b, 16);
            ensure(b, s, lane);
            const uint32_t nlen = take(b, 16);
            if ((len ^ nlen) != 0xFFFFu) { err = BGZF_BAD_STORED; break; }
            if (p + len > out_len) { err = BGZF_OVERRUN; break; }
            for (uint32_t i = 0; i < len; i++) {
                ensure(b, s, lane);
                const uint32_t byte = take(b, 8);
                if (lane == 0) s.ring[p & RMASK] = (uint8_t)byte;
                p++;
                if (p - flushed >= FLUSH) { flush(s, out, flushed, p, lane); flushed = p; }
            }
            continue;
        }
        if (type == 3) { err = BGZF_BAD_BLOCK_TYPE; break; }
        int hlit, hdist;
        if (type == 1) {
            // ---- fixed codes (RFC 1951 3.2.6) ----
            hlit = 288; hdist = 30;
            __syncwarp();
            for (int i = lane; i < 288; i += 32) s.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
            if (lane < 30) s.lens[288 + lane] = 5;
            __syncwarp();
        } else {
            // ---- dynamic codes (3.2.7) ----
            ensure(b, s, lane);
            hlit = (int)take(b, 5) + 257;
            hdist = (int)take(b, 5) + 1;
            const int hclen = (int)take(b, 4) + 4;
            if (hlit > 286 || hdist > 30) { err = BGZF_BAD_LENGTHS; break; }
            __syncwarp();
            if (lane < 19) s.lens[lane] = 0;
            __syncwarp();
            for (int i = 0; i < hclen; i++) {
                ensure(b, s, lane);
                const uint32_t l = take(b, 3);
                if (lane == 0) s.lens[c_cl_order[i]] = (uint8_t)l;
            }
            if (!build_tables(s.lens, 19, s.cl_count, s.cl_sym, s.cl_fast, CL_FAST, lane)) { err = BGZF_BAD_LENGTHS; break; }
            // the code lengths of both alphabets, run-length coded; staged in registers of the walk and written by lane 0
            int i = 0, prev = 0;
            const int total = hlit + hdist;
            while (i < total && err == BGZF_OK) {
                ensure(b, s, lane);
                const int sym = decode_symbol(b, s.cl_fast, CL_FAST, s.cl_count, s.cl_sym);
                if (sym < 0) { err = BGZF_BAD_CODE; break; }
                int rep = 1, val = sym;
                if (sym == 16) { if (i == 0) { err = BGZF_BAD_LENGTHS; break; } rep = 3 + (int)take(b, 2); val = prev; }
                else if (sym == 17) { rep = 3 + (int)take(b, 3); val = 0; }
                else if (sym == 18) { rep = 11 + (int)take(b, 7); val = 0; }
                if (i + rep > total) { err = BGZF_BAD_LENGTHS; break; }
                // lens[0..hlit) literal/length, lens[288..288+hdist) distance
                for (int k = lane; k < rep; k += 32) {
                    const int j = i + k;
                    s.lens[j < hlit ? j : 288 + (j - hlit)] = (uint8_t)val;
                }
                i += rep; prev = val;
            }
            if (err != BGZF_OK) break;
            __syncwarp();
            if (s.lens[256] == 0) { err = BGZF_BAD_LENGTHS; break; }
        }
        if (!build_tables(s.lens, hlit, s.lit_count, s.lit_sym, s.lit_fast, LIT_FAST, lane)) { err = BGZF_BAD_LENGTHS; break; }
        if (!build_tables(s.lens + 288, hdist, s.dist_count, s.dist_sym, s.dist_fast, DIST_FAST, lane)) { err = BGZF_BAD_LENGTHS; break; }

        // ---- the symbols of this deflate block ----
        for (;;) {
            ensure(b, s, lane);
            if (SPECULATE) {
                // Runs of literals.  Lane k looks up the code that WOULD start k bits into the buffer (one shared-memory gather for
                // all 32 offsets); the chain "this code is L bits long, so the next one starts L bits later" then hops from lane to
                // lane by register shuffles instead of going through the table once per symbol.  It stops at the first code that is
                // not a short literal, which the ordinary path below decodes.
                const uint32_t my_e = s.lit_fast[(uint32_t)(b.buf >> lane) & ((1u << LIT_FAST) - 1u)];
                const int lim = min(31, b.cnt - LIT_FAST);        // offsets whose 10 lookup bits are all valid
                int pos = 0, n = 0;
                uint32_t mine = 0;
                bool more = false;
                for (;;) {
                    const uint32_t e = __shfl_sync(0xFFFFFFFFu, my_e, pos);
                    if (e == 0u || e >= (256u << 4)) break;
                    if (lane == n) mine = e >> 4;
                    n++;
                    pos += (int)(e & 15u);
                    if (pos > lim) { more = true; break; }
                }
                if (n) {
                    if (p + (uint32_t)n > out_len) { err = BGZF_OVERRUN; break; }
                    if (lane < n) s.ring[(p + lane) & RMASK] = (uint8_t)mine;
                    p += (uint32_t)n;
                    b.buf >>= pos;
                    b.cnt -= pos;
                    if (p - flushed >= FLUSH) { flush(s, out, flushed, p, lane); flushed = p; }
                    if (more) continue;
                    ensure(b, s, lane);
                }
            }
            const int sym = decode_symbol(b, s.lit_fast, LIT_FAST, s.lit_count, s.lit_sym);
            if (sym < 256) {
                if (sym < 0) { err = BGZF_BAD_CODE; break; }
                if (p >= out_len) { err = BGZF_OVERRUN; break; }
                if (lane == 0) s.ring[p & RMASK] = (uint8_t)sym;
                p++;
            } else {
                if (sym == 256) break;
                const int li = sym - 257;
                if (li >= 29) { err = BGZF_BAD_CODE; break; }
                const uint32_t len = c_len_base[li] + take(b, c_len_extra[li]);
                ensure(b, s, lane);
                const int ds = decode_symbol(b, s.dist_fast, DIST_FAST, s.dist_count, s.dist_sym);
                if (ds < 0 || ds >= 30) { err = BGZF_BAD_CODE; break; }
                const uint32_t dist = c_dist_base[ds] + take(b, c_dist_extra[ds]);
                if (dist > p) { err = BGZF_BAD_DISTANCE; break; }
                if (p + len > out_len) { err = BGZF_OVERRUN; break; }
                const uint32_t from = p - dist;
                __syncwarp();
                if (dist <= (uint32_t)NEAR_MAX) {
                    if (dist >= len) {
                        for (uint32_t i = lane; i < len; i += 32) s.ring[(p + i) & RMASK] = s.ring[(from + i) & RMASK];
                    } else {
                        // the source overlaps the destination: the bytes repeat with period `dist`
                        for (uint32_t i = lane; i < len; i += 32) s.ring[(p + i) & RMASK] = s.ring[(from + i % dist) & RMASK];
                    }
                } else {
                    // further back than the ring keeps: those bytes were flushed long ago (flushed >= p - FLUSH)
                    for (uint32_t i = lane; i < len; i += 32) s.ring[(p + i) & RMASK] = __ldcg(out + from + i);
                }
                __syncwarp();
                p += len;
            }
            if (p - flushed >= FLUSH) { flush(s, out, flushed, p, lane); flushed = p; }
        }
    }
    if (err == BGZF_OK) {
        flush(s, out, flushed, p, lane);
        if (p != out_len) err = BGZF_SIZE_MISMATCH;
    }
    if (lane == 0) status[blk] = (uint8_t)err;
}

const char *bgzf_error_name(int e) {
    switch (e) {
        case BGZF_BAD_BLOCK_TYPE: return "reserved deflate block type";
        case BGZF_BAD_STORED: return "stored block length check failed";
        case BGZF_BAD_CODE: return "invalid Huffman code";
        case BGZF_BAD_DISTANCE: return "match distance reaches before the start of the block";
        case BGZF_OVERRUN: return "more output than ISIZE announces";
        case BGZF_BAD_LENGTHS: return "invalid code lengths";
        case BGZF_SIZE_MISMATCH: return "output shorter than ISIZE announces";
        default: return "unknown error";
    }
}

// CRC-32 (IEEE 802.3, the one zlib's crc32() computes), slicing-by-8, host side
struct CrcTables {
    uint32_t t[8][256];
    CrcTables() {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; i++)
            for (int k = 1; k < 8; k++) t[k][i] = (t[k - 1][i] >> 8) ^ t[0][t[k - 1][i] & 0xFFu];
    }
};
uint32_t crc32_host(const uint8_t *p, size_t n) {
    static const CrcTables T;
    uint32_t c = 0xFFFFFFFFu;
    while (n >= 8) {
        uint32_t a, b;
        memcpy(&a, p, 4); memcpy(&b, p + 4, 4);
        a ^= c;
        c = T.t[7][a & 0xFFu] ^ T.t[6][(a >> 8) & 0xFFu] ^ T.t[5][(a >> 16) & 0xFFu] ^ T.t[4][a >> 24] ^ T.t[3][b & 0xFFu] ^
            T.t[2][(b >> 8) & 0xFFu] ^ T.t[1][(b >> 16) & 0xFFu] ^ T.t[0][b >> 24];
        p += 8; n -= 8;
    }
    while (n--) c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xFFu];
    return ~c;
}

// enqueues the kernel for blocks [0, n) of d_blocks on `st`; their status bytes go to d_status[0 .. n)
void enqueue_inflate(lps_ctx *ctx, cudaStream_t st, const uint8_t *d_data, const lps_bgzf_block *d_blocks, uint64_t n, uint8_t *d_out,
                     uint8_t *d_status) {
    if (!n) return;
    const char *env = getenv("LPS_BGZF_SPECULATE");
    const bool speculate = !(env && env[0] == '0');
    (speculate ? k_bgzf_inflate<true> : k_bgzf_inflate<false>)<<<(unsigned)((n + CTA_WARPS - 1) / CTA_WARPS), 32 * CTA_WARPS, CTA_WARPS * sizeof(WarpSmem), st>>>(
        (uint32_t)n, d_data, d_blocks, d_out, d_status);
    ctx->stats.kernel_launches++;
}

int check_status(lps_ctx *ctx, uint64_t n_blocks) {
    for (uint64_t k = 0; k < n_blocks; k++)
        if (ctx->h_bgzf_status[k] != BGZF_OK)
            return ctx->fail(LPS_E_DATA, "BGZF block " + std::to_string(k) + ": " + bgzf_error_name(ctx->h_bgzf_status[k]));
    return LPS_OK;
}

int launch_inflate(lps_ctx *ctx, const uint8_t *d_data, const lps_bgzf_block *d_blocks, uint64_t n_blocks, uint8_t *d_out) {
    LPS_CUDA(ctx, ctx->d_bgzf_status.reserve((size_t)n_blocks + 1));
    cudaEventRecord(ctx->kev[4], ctx->stream);
    enqueue_inflate(ctx, ctx->stream, d_data, d_blocks, n_blocks, d_out, ctx->d_bgzf_status.p);
    cudaEventRecord(ctx->kev[5], ctx->stream);
    LPS_CUDA(ctx, cudaGetLastError());
    ctx->h_bgzf_status.resize((size_t)n_blocks);
    if (n_blocks)
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->h_bgzf_status.data(), ctx->d_bgzf_status.p, (size_t)n_blocks, cudaMemcpyDeviceToHost, ctx->stream));
    LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stats.ms_kernel_bgzf, ctx->kev[4], ctx->kev[5]);
    return check_status(ctx, n_blocks);
}

int check_blocks(lps_ctx *ctx, const lps_bgzf_block *blocks, uint64_t n_blocks, uint64_t n_bytes, uint64_t out_cap) {
    if (n_blocks > 0x7FFFFFFFull) return ctx->fail(LPS_E_ARG, "too many BGZF blocks for one call");
    for (uint64_t k = 0; k < n_blocks; k++) {
        const lps_bgzf_block &b = blocks[k];
        if (b.comp_off + b.comp_len > n_bytes || b.out_off + b.out_len > out_cap || b.out_len > 65536u)
            return ctx->fail(LPS_E_ARG, "BGZF block " + std::to_string(k) + " lies outside the buffers");
    }
    return LPS_OK;
}

}  // namespace

// check_header (htslib/bgzf.c:876-883) + the member walk of bgzf_read_block (:1135-1176)
int lps_bgzf_scan(const uint8_t *data, uint64_t n_bytes, lps_bgzf_block *blocks, uint64_t cap, uint64_t *n_blocks, uint64_t *out_bytes) {
    if ((n_bytes && !data) || !n_blocks || !out_bytes) return LPS_E_ARG;
    uint64_t off = 0, n = 0, total = 0;
    while (off < n_bytes) {
        if (n_bytes - off < 18 + 8) return LPS_E_DATA;
        const uint8_t *h = data + off;
        // ID1 ID2 CM FLG.FEXTRA, XLEN == 6, subfield 'B' 'C' of length 2: exactly what htslib accepts
        if (!(h[0] == 31 && h[1] == 139 && h[2] == 8 && (h[3] & 4) != 0 && (h[10] | (h[11] << 8)) == 6 && h[12] == 'B' && h[13] == 'C' &&
              (h[14] | (h[15] << 8)) == 2))
            return LPS_E_DATA;
        const uint64_t block_length = (uint64_t)(h[16] | (h[17] << 8)) + 1;
        if (block_length < 18 + 8 || off + block_length > n_bytes) return LPS_E_DATA;
        const uint8_t *t = h + block_length - 8;
        const uint32_t isize = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
        if (isize > 65536u) return LPS_E_DATA;
        if (blocks && n < cap) {
            blocks[n].comp_off = off + 18; blocks[n].comp_len = (uint32_t)(block_length - 18 - 8);
            blocks[n].out_len = isize; blocks[n].out_off = total;
            blocks[n].crc32 = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
            blocks[n].reserved_ = 0;
        }
        n++; total += isize; off += block_length;
    }
    *n_blocks = n; *out_bytes = total;
    return (blocks && n > cap) ? LPS_E_ARG : LPS_OK;
}

int lps_bgzf_inflate(lps_ctx *ctx, const uint8_t *data, uint64_t n_bytes, const lps_bgzf_block *blocks, uint64_t n_blocks, uint8_t *out,
                     uint64_t out_cap, int check_crc) {
    if (!ctx) return LPS_E_ARG;
    if ((n_bytes && !data) || (n_blocks && !blocks) || (out_cap && !out)) return ctx->fail(LPS_E_ARG, "null buffer");
    int rc = check_blocks(ctx, blocks, n_blocks, n_bytes, out_cap);
    if (rc != LPS_OK) return rc;
    cudaSetDevice(ctx->device);
    uint64_t out_bytes = 0;
    for (uint64_t k = 0; k < n_blocks; k++) out_bytes = std::max<uint64_t>(out_bytes, blocks[k].out_off + blocks[k].out_len);
    LPS_CUDA(ctx, ctx->d_bgzf_in.reserve((size_t)n_bytes + 16));
    LPS_CUDA(ctx, ctx->d_bgzf_out.reserve((size_t)out_bytes + 16));
    LPS_CUDA(ctx, ctx->d_bgzf_blocks.reserve((size_t)n_blocks + 1));
    LPS_CUDA(ctx, ctx->d_bgzf_status.reserve((size_t)n_blocks + 1));
    ctx->h_bgzf_status.resize((size_t)n_blocks);
    if (n_blocks)
        LPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_bgzf_blocks.p, blocks, (size_t)n_blocks * sizeof(lps_bgzf_block), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += n_bytes + n_blocks * sizeof(lps_bgzf_block);
    ctx->stats.d2h_bytes += out_bytes;
    // The table lps_bgzf_scan makes is ascending in both offsets: the blocks are cut into chunks of ~64 MB of output, and the
    // upload of chunk c+1, the kernel of chunk c and the download of chunk c-1 run at the same time on three streams (PCIe is full
    // duplex).  Any other table takes the plain path: one upload, one launch, one download.
    bool ascending = true;
    for (uint64_t k = 1; k < n_blocks && ascending; k++)
        ascending = blocks[k].comp_off >= blocks[k - 1].comp_off + blocks[k - 1].comp_len && blocks[k].out_off >= blocks[k - 1].out_off + blocks[k - 1].out_len;
    uint64_t chunk_out = 64ull << 20;
    if (const char *env = getenv("LPS_BGZF_CHUNK")) chunk_out = strtoull(env, nullptr, 10);       // bytes of output per chunk (tests)
    if (!ascending || out_bytes <= chunk_out) {
        if (n_bytes) LPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_bgzf_in.p, data, (size_t)n_bytes, cudaMemcpyHostToDevice, ctx->stream));
        rc = launch_inflate(ctx, ctx->d_bgzf_in.p, ctx->d_bgzf_blocks.p, n_blocks, ctx->d_bgzf_out.p);
        if (rc != LPS_OK) return rc;
        if (out_bytes) LPS_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_bgzf_out.p, (size_t)out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        LPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        if (!ctx->stream_up) LPS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream_up, cudaStreamNonBlocking));
        if (!ctx->stream_down) LPS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream_down, cudaStreamNonBlocking));
        // one block keeps a warp busy for milliseconds whatever the size of the launch, so a chunk alone (~1000 blocks) would leave
        // most of the 4736 warp slots empty: the kernels of consecutive chunks go to four streams and run side by side
        for (auto &cs : ctx->stream_k)
            if (!cs) LPS_CUDA(ctx, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaEvent_t table_ready;
        cudaEventCreateWithFlags(&table_ready, cudaEventDisableTiming);
        cudaEventRecord(table_ready, ctx->stream);
        for (auto &cs : ctx->stream_k) cudaStreamWaitEvent(cs, table_ready, 0);
        size_t chunk_no = 0;
        std::vector<cudaEvent_t> up, done;
        cudaEventRecord(ctx->kev[4], ctx->stream);
        for (uint64_t b0 = 0; b0 < n_blocks;) {
            uint64_t b1 = b0 + 1;
            while (b1 < n_blocks && blocks[b1].out_off + blocks[b1].out_len - blocks[b0].out_off <= chunk_out) b1++;
            const uint64_t c_lo = blocks[b0].comp_off, c_hi = blocks[b1 - 1].comp_off + blocks[b1 - 1].comp_len;
            const uint64_t o_lo = blocks[b0].out_off, o_hi = blocks[b1 - 1].out_off + blocks[b1 - 1].out_len;
            cudaEvent_t e_up, e_done;
            cudaEventCreateWithFlags(&e_up, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&e_done, cudaEventDisableTiming);
            up.push_back(e_up); done.push_back(e_done);
            if (c_hi > c_lo) cudaMemcpyAsync(ctx->d_bgzf_in.p + c_lo, data + c_lo, (size_t)(c_hi - c_lo), cudaMemcpyHostToDevice, ctx->stream_up);
            cudaEventRecord(e_up, ctx->stream_up);
            cudaStream_t ks = ctx->stream_k[chunk_no++ % 4];
            cudaStreamWaitEvent(ks, e_up, 0);
            enqueue_inflate(ctx, ks, ctx->d_bgzf_in.p, ctx->d_bgzf_blocks.p + b0, b1 - b0, ctx->d_bgzf_out.p, ctx->d_bgzf_status.p + b0);
            cudaEventRecord(e_done, ks);
            cudaStreamWaitEvent(ctx->stream, e_done, 0);       // the context's stream joins every chunk: it carries the end event and the status copy
            cudaStreamWaitEvent(ctx->stream_down, e_done, 0);
            if (o_hi > o_lo) cudaMemcpyAsync(out + o_lo, ctx->d_bgzf_out.p + o_lo, (size_t)(o_hi - o_lo), cudaMemcpyDeviceToHost, ctx->stream_down);
            b0 = b1;
        }
        cudaEventRecord(ctx->kev[5], ctx->stream);
        cudaMemcpyAsync(ctx->h_bgzf_status.data(), ctx->d_bgzf_status.p, (size_t)n_blocks, cudaMemcpyDeviceToHost, ctx->stream);
        const cudaError_t e1 = cudaStreamSynchronize(ctx->stream), e2 = cudaStreamSynchronize(ctx->stream_down), e3 = cudaStreamSynchronize(ctx->stream_up);
        for (cudaEvent_t e : up) cudaEventDestroy(e);
        for (cudaEvent_t e : done) cudaEventDestroy(e);
        cudaEventDestroy(table_ready);
        LPS_CUDA(ctx, e1); LPS_CUDA(ctx, e2); LPS_CUDA(ctx, e3);
        LPS_CUDA(ctx, cudaGetLastError());
        cudaEventElapsedTime(&ctx->stats.ms_kernel_bgzf, ctx->kev[4], ctx->kev[5]);   // first launch to last completion, waits for uploads included
        rc = check_status(ctx, n_blocks);
        if (rc != LPS_OK) return rc;
    }
    if (check_crc)
        for (uint64_t k = 0; k < n_blocks; k++)
            if (crc32_host(out + blocks[k].out_off, blocks[k].out_len) != blocks[k].crc32)
                return ctx->fail(LPS_E_DATA, "BGZF block " + std::to_string(k) + ": CRC32 checksum mismatch");   // htslib/bgzf.c:777-781
    return LPS_OK;
}

int lps_bgzf_inflate_device(lps_ctx *ctx, const uint8_t *d_data, const lps_bgzf_block *d_blocks, uint64_t n_blocks, uint8_t *d_out) {
    if (!ctx) return LPS_E_ARG;
    if (n_blocks && (!d_data || !d_blocks || !d_out)) return ctx->fail(LPS_E_ARG, "null buffer");
    if (n_blocks > 0x7FFFFFFFull) return ctx->fail(LPS_E_ARG, "too many BGZF blocks for one call");
    cudaSetDevice(ctx->device);
    return launch_inflate(ctx, d_data, d_blocks, n_blocks, d_out);
}
