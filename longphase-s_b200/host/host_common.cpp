// host_common.cpp — error plumbing and small helpers of the C++ host (see host_common.h)
#include "host_common.h"

#include <zlib.h>

namespace lpsh {

static thread_local std::string g_error;
static std::string g_error_any;

int fail(const std::string &message) {
    g_error = message;
#pragma omp critical(lpsh_error)
    g_error_any = message;
    return -1;
}

bool read_gz(const std::string &path, std::string &text) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) return false;
    std::vector<char> buf(1 << 20);
    for (;;) {
        const int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n < 0) { int e; fprintf(stderr, "Error: %s.\n", gzerror(f, &e)); exit(EXIT_FAILURE); }
        if (n == 0) break;
        text.append(buf.data(), (size_t)n);
    }
    gzclose(f);
    return true;
}

int device_count() {
    int n = 0;
    for (; n < 64; n++) {
        lps_ctx *c = nullptr;
        if (lps_ctx_create(n, &c) != 0) break;
        lps_ctx_destroy(c);
    }
    return n;
}

}  // namespace lpsh

extern "C" const char *lpsh_last_error(void) {
    return lpsh::g_error.empty() ? lpsh::g_error_any.c_str() : lpsh::g_error.c_str();
}

// ---- synthetic batches -> BAM + BAI (bench / test tooling: the generator's SoA contigs written with htslib) ---------------
struct lpsh_bamw {
    samFile *out = nullptr;
    sam_hdr_t *hdr = nullptr;
    htsThreadPool pool = {NULL, 0};
    std::string path;
};

extern "C" {

lpsh_bamw *lpsh_bamw_open(const char *path, int n_contigs, const char **names, const int64_t *lens, int threads) {
    lpsh_bamw *w = new lpsh_bamw();
    w->path = path;
    w->out = sam_open(path, "wb");
    w->hdr = sam_hdr_init();
    if (!w->out || !w->hdr) { delete w; return nullptr; }
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (int i = 0; i < n_contigs; i++) text += std::string("@SQ\tSN:") + names[i] + "\tLN:" + std::to_string(lens[i]) + "\n";
    sam_hdr_add_lines(w->hdr, text.c_str(), text.size());
    if (threads > 1 && (w->pool.pool = hts_tpool_init(threads))) hts_set_opt(w->out, HTS_OPT_THREAD_POOL, &w->pool);
    if (sam_hdr_write(w->out, w->hdr) < 0) { delete w; return nullptr; }
    return w;
}

int lpsh_bamw_append(lpsh_bamw *w, int tid, const lps_read_batch *b, const char *names, int name_stride) {
    if (!w || !b) return -1;
    bam1_t *rec = bam_init1();
    std::string seq;
    int rc = 0;
    for (int32_t r = 0; r < b->n_reads && rc == 0; r++) {
        const int lq = b->l_qseq[r];
        seq.resize((size_t)lq);
        const uint8_t *s4 = b->seq4 + b->seq_off[r];
        for (int k = 0; k < lq; k++) seq[(size_t)k] = seq_nt16_str[(s4[k >> 1] >> ((~k & 1) << 2)) & 15];
        const char *name = names + (size_t)r * (size_t)name_stride;
        if (bam_set1(rec, strlen(name), name, b->flag[r], tid, b->ref_start[r], b->mapq[r], b->n_cigar[r], b->cigar + b->cigar_off[r], -1, -1, 0,
                     (size_t)lq, seq.data(), (const char *)(b->qual + b->qual_off[r]), 0) < 0 ||
            sam_write1(w->out, w->hdr, rec) < 0)
            rc = -1;
    }
    bam_destroy1(rec);
    return rc;
}

int lpsh_bamw_close(lpsh_bamw *w) {
    if (!w) return -1;
    int rc = sam_close(w->out) < 0 ? -1 : 0;
    sam_hdr_destroy(w->hdr);
    if (w->pool.pool) hts_tpool_destroy(w->pool.pool);
    if (rc == 0 && sam_index_build(w->path.c_str(), 0) < 0) rc = -1;
    delete w;
    return rc;
}

}  // extern "C"
