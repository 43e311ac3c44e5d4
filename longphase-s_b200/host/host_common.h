// host_common.h — pieces shared by the sub-commands of the C++ host: the SoA packer that turns htslib records into the
// lps_read_batch / lps_variants of include/lps.h, error plumbing, small file helpers.
#ifndef LPS_HOST_COMMON_H
#define LPS_HOST_COMMON_H

#include <htslib/faidx.h>
#include <htslib/kroundup.h>
#include <htslib/sam.h>
#include <htslib/thread_pool.h>
#include <htslib/vcf.h>

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include <fstream>
#include <functional>
#include <iterator>
#include <sstream>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "lps_host.h"

namespace lpsh {

// the output files carry the version of the tool whose format they follow (##longphaseVersion, @PG VN)
static const char *const REFERENCE_VERSION = "1.0.0";

// hts_open mode of the tagged BAM: "wb" as the reference, or "wb<L>" when LPS_BAM_LEVEL=0..9 is set (the BGZF writer's zlib deflate is
// what the tagging passes wait for; a lower level trades file size for wall time, the records are the same)
std::string bam_write_mode();
double now_ms();                                           // monotonic clock, for the [timing] lines on stderr
int fail(const std::string &message);                     // remembers the message for lpsh_last_error, returns -1
bool read_gz(const std::string &path, std::string &text); // whole file through zlib (plain files pass through)
int device_count();                                       // CUDA devices liblps_b200.so can open a context on

// One contig in the layout of include/lps.h.  Alignments are appended in file order; finish() derives the offsets'
// companions that need the whole batch (name ranks).
struct PackedContig {
    // variant table
    std::vector<int32_t> v_pos;
    std::vector<uint8_t> v_ref0, v_alt0, v_hp1_is_alt, v_gt_kind;
    std::vector<uint16_t> v_ref_len, v_alt_len;
    std::vector<int32_t> v_ps;
    bool tagged_variants = false;   // hp1_is_alt / ps / gt_kind are meaningful (tag family)
    // reads
    std::vector<int32_t> ref_start, l_qseq, name_rank;
    std::vector<uint32_t> n_cigar, cigar;
    std::vector<uint64_t> cigar_off, seq_off, qual_off, name_off;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> mapq, seq4, qual;
    std::string names;
    std::string ref;
    const std::string *ref_shared = nullptr;   // when set, the reference string lives elsewhere (one copy per contig, not per chunk)

    void add_variant(int pos, const std::string &ref_text, const std::string &alt_text) {
        v_pos.push_back(pos);
        v_ref0.push_back(ref_text.empty() ? 0 : (uint8_t)ref_text[0]);
        v_alt0.push_back(alt_text.empty() ? 0 : (uint8_t)alt_text[0]);
        v_ref_len.push_back((uint16_t)std::min<size_t>(ref_text.size(), 65535));
        v_alt_len.push_back((uint16_t)std::min<size_t>(alt_text.size(), 65535));
    }
    void add_alignment(const bam1_t *b) {
        const bam1_core_t &c = b->core;
        ref_start.push_back((int32_t)c.pos);
        l_qseq.push_back(c.l_qseq);
        n_cigar.push_back(c.n_cigar);
        flag.push_back(c.flag);
        mapq.push_back(c.qual);
        cigar_off.push_back(cigar.size());
        const uint32_t *cg = bam_get_cigar(b);
        cigar.insert(cigar.end(), cg, cg + c.n_cigar);
        seq_off.push_back(seq4.size());
        const uint8_t *s = bam_get_seq(b);
        seq4.insert(seq4.end(), s, s + (c.l_qseq + 1) / 2);
        qual_off.push_back(qual.size());
        const uint8_t *q = bam_get_qual(b);
        qual.insert(qual.end(), q, q + c.l_qseq);
        name_off.push_back(names.size());
        names.append(bam_get_qname(b));
        names.push_back('\0');
    }
    // the same from the bytes of a BAM record as they lie in the file (after the block_size word; BAM is little endian, SAM spec 4.2).
    // false for a record whose real CIGAR sits in the CG tag (more than 65535 ops): the caller falls back to htslib for it
    bool add_raw_record(const uint8_t *p, uint32_t block_size) {
        auto le32 = [](const uint8_t *q) { return (uint32_t)q[0] | (uint32_t)q[1] << 8 | (uint32_t)q[2] << 16 | (uint32_t)q[3] << 24; };
        auto le16 = [](const uint8_t *q) { return (uint32_t)q[0] | (uint32_t)q[1] << 8; };
        const uint32_t l_name = p[8], n_cig = le16(p + 12), fl = le16(p + 14), l_seq = le32(p + 16);
        if (32ull + l_name + 4ull * n_cig + (l_seq + 1) / 2 + l_seq > block_size) return false;
        if (l_name == 0 || p[32 + l_name - 1] != '\0') return false;      // name not NUL-terminated inside l_read_name: htslib repairs it, we decline
        const uint8_t *name = p + 32, *cg = name + l_name, *sq = cg + 4 * n_cig, *ql = sq + (l_seq + 1) / 2;
        if (n_cig == 2 && (le32(cg) & 15u) == 4 /* S */ && (le32(cg) >> 4) == l_seq && (le32(cg + 4) & 15u) == 3 /* N */) return false;
        ref_start.push_back((int32_t)le32(p + 4));
        l_qseq.push_back((int32_t)l_seq);
        n_cigar.push_back(n_cig);
        flag.push_back((uint16_t)fl);
        mapq.push_back(p[9]);
        cigar_off.push_back(cigar.size());
        for (uint32_t k = 0; k < n_cig; k++) cigar.push_back(le32(cg + 4 * k));
        seq_off.push_back(seq4.size());
        seq4.insert(seq4.end(), sq, sq + (l_seq + 1) / 2);
        qual_off.push_back(qual.size());
        qual.insert(qual.end(), ql, ql + l_seq);
        name_off.push_back(names.size());
        names.append((const char *)name, strnlen((const char *)name, l_name - 1));   // never reads beyond the name field
        names.push_back('\0');
        return true;
    }
    // capacity hints from the previous chunk of the pass (chunks are alike): no regrowth copies while a chunk fills
    struct Sizes { size_t reads = 0, cigar = 0, seq4 = 0, qual = 0, names = 0; };
    Sizes sizes() const { Sizes z; z.reads = ref_start.size(); z.cigar = cigar.size(); z.seq4 = seq4.size(); z.qual = qual.size(); z.names = names.size(); return z; }
    void reserve_sizes(const Sizes &z) {
        auto grow = [](size_t n) { return n + n / 8 + 16; };
        if (!z.reads) return;
        const size_t n = grow(z.reads);
        ref_start.reserve(n); l_qseq.reserve(n); n_cigar.reserve(n); flag.reserve(n); mapq.reserve(n);
        cigar_off.reserve(n); seq_off.reserve(n); qual_off.reserve(n); name_off.reserve(n);
        cigar.reserve(grow(z.cigar)); seq4.reserve(grow(z.seq4)); qual.reserve(grow(z.qual)); names.reserve(grow(z.names));
    }
    // appends the alignments of `parts` (in that order) with one thread per part: sizes first, then every part copies itself into its
    // slice of the grown arrays, offsets rebased
    void append_parts(const std::vector<PackedContig> &parts) {
        const size_t np = parts.size();
        std::vector<size_t> r0(np + 1), c0(np + 1), s0(np + 1), q0(np + 1), n0(np + 1);
        r0[0] = ref_start.size(); c0[0] = cigar.size(); s0[0] = seq4.size(); q0[0] = qual.size(); n0[0] = names.size();
        for (size_t k = 0; k < np; k++) {
            r0[k + 1] = r0[k] + parts[k].ref_start.size(); c0[k + 1] = c0[k] + parts[k].cigar.size(); s0[k + 1] = s0[k] + parts[k].seq4.size();
            q0[k + 1] = q0[k] + parts[k].qual.size(); n0[k + 1] = n0[k] + parts[k].names.size();
        }
        ref_start.resize(r0[np]); l_qseq.resize(r0[np]); n_cigar.resize(r0[np]); flag.resize(r0[np]); mapq.resize(r0[np]);
        cigar_off.resize(r0[np]); seq_off.resize(r0[np]); qual_off.resize(r0[np]); name_off.resize(r0[np]);
        cigar.resize(c0[np]); seq4.resize(s0[np]); qual.resize(q0[np]); names.resize(n0[np]);
        std::vector<std::thread> workers;
        for (size_t k = 0; k < np; k++)
            workers.emplace_back([&, k] {
                const PackedContig &p = parts[k];
                const size_t n = p.ref_start.size();
                std::copy(p.ref_start.begin(), p.ref_start.end(), ref_start.begin() + (ptrdiff_t)r0[k]);
                std::copy(p.l_qseq.begin(), p.l_qseq.end(), l_qseq.begin() + (ptrdiff_t)r0[k]);
                std::copy(p.n_cigar.begin(), p.n_cigar.end(), n_cigar.begin() + (ptrdiff_t)r0[k]);
                std::copy(p.flag.begin(), p.flag.end(), flag.begin() + (ptrdiff_t)r0[k]);
                std::copy(p.mapq.begin(), p.mapq.end(), mapq.begin() + (ptrdiff_t)r0[k]);
                for (size_t r = 0; r < n; r++) {
                    cigar_off[r0[k] + r] = p.cigar_off[r] + c0[k]; seq_off[r0[k] + r] = p.seq_off[r] + s0[k];
                    qual_off[r0[k] + r] = p.qual_off[r] + q0[k]; name_off[r0[k] + r] = p.name_off[r] + n0[k];
                }
                std::copy(p.cigar.begin(), p.cigar.end(), cigar.begin() + (ptrdiff_t)c0[k]);
                std::copy(p.seq4.begin(), p.seq4.end(), seq4.begin() + (ptrdiff_t)s0[k]);
                std::copy(p.qual.begin(), p.qual.end(), qual.begin() + (ptrdiff_t)q0[k]);
                std::copy(p.names.begin(), p.names.end(), names.begin() + (ptrdiff_t)n0[k]);
            });
        for (std::thread &w : workers) w.join();
    }
    void truncate_reads(size_t n) {   // forget the alignments appended after the first n
        if (n >= ref_start.size()) return;
        cigar.resize((size_t)cigar_off[n]); seq4.resize((size_t)seq_off[n]); qual.resize((size_t)qual_off[n]); names.resize((size_t)name_off[n]);
        ref_start.resize(n); l_qseq.resize(n); n_cigar.resize(n); flag.resize(n); mapq.resize(n);
        cigar_off.resize(n); seq_off.resize(n); qual_off.resize(n); name_off.resize(n);
    }
    int32_t n_reads() const { return (int32_t)ref_start.size(); }
    // rank of every read name in std::string order; equal names share a rank (the reference folds edge weights in
    // std::map<std::string, ...> order, PhasingGraph.cpp:697,848)
    void finish() {
        const size_t n = ref_start.size();
        std::vector<uint32_t> order(n);
        for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
        const char *base = names.data();
        auto name = [&](uint32_t r) { return base + name_off[r]; };
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            const int c = strcmp(name(a), name(b));   // strcmp compares as unsigned char, like char_traits<char>::compare
            return c != 0 ? c < 0 : a < b;
        });
        name_rank.assign(n, 0);
        int32_t rank = -1;
        for (size_t k = 0; k < n; k++) {
            if (k == 0 || strcmp(name(order[k - 1]), name(order[k])) != 0) rank++;
            name_rank[order[k]] = rank;
        }
    }
    void view(lpsh_packed *out) const {
        memset(out, 0, sizeof(*out));
        lps_variants &v = out->variants;
        v.n = (int32_t)v_pos.size();
        v.pos = v_pos.data(); v.ref0 = v_ref0.data(); v.alt0 = v_alt0.data();
        v.ref_len = v_ref_len.data(); v.alt_len = v_alt_len.data();
        if (tagged_variants) { v.hp1_is_alt = v_hp1_is_alt.data(); v.ps = v_ps.data(); v.gt_kind = v_gt_kind.data(); }
        lps_read_batch &b = out->batch;
        b.n_reads = n_reads();
        b.ref_start = ref_start.data(); b.l_qseq = l_qseq.data(); b.n_cigar = n_cigar.data();
        b.cigar_off = cigar_off.data(); b.seq_off = seq_off.data(); b.qual_off = qual_off.data();
        b.flag = flag.data(); b.mapq = mapq.data(); b.name_rank = name_rank.data();
        b.cigar = cigar.data(); b.cigar_len = cigar.size();
        b.seq4 = seq4.data(); b.seq_bytes = seq4.size();
        b.qual = qual.data(); b.qual_bytes = qual.size();
        const std::string &r = ref_shared ? *ref_shared : ref;
        out->ref = r.data(); out->ref_len = (int64_t)r.size();
        out->names = names.data(); out->name_off = name_off.data();
    }
};

// ---- small helpers shared by the sub-commands ----------------------------------------------------------------------------
template <class T>
inline void take(const char *text, T &dst) {   // the reference reads every option value with operator>> of an istringstream
    std::istringstream in(text ? text : "");
    in >> dst;
}
// FileValidator::validateRequiredFile (src/shared/ArgumentManager.cpp:147-160)
inline bool required_file(const char *program, const std::string &path, const char *what) {
    if (path.empty()) { std::cerr << "[ERROR] " << program << ": missing " << what << ".\n"; return false; }
    if (!std::ifstream(path.c_str()).is_open()) { std::cerr << "[ERROR] " << program << ": " << what << ": " << path << " not exist.\n\n"; return false; }
    return true;
}
// the reference addresses FORMAT keys and sample values by counting ':' (ParsingBam.cpp:494-560, HaplotagVcfParser.cpp:238-262)
inline int subfield_of(const std::string &format, const char *key) {   // index of the sub-field whose text contains `key`, 0 if absent
    const size_t at = format.find(key);
    if (at == std::string::npos) return 0;
    return (int)std::count(format.begin(), format.begin() + at, ':');
}
inline size_t subfield_start(const std::string &sample, int k) {      // first character of the k-th ':' separated value (size() if fewer)
    size_t i = 0;
    for (int seen = 0; i < sample.size() && seen < k; i++) seen += sample[i] == ':';
    return i;
}
inline char peek(const std::string &s, size_t i) { return i < s.size() ? s[i] : '\0'; }
inline void drop_aux(bam1_t *b, const char *tag) {                     // GermlineHaplotagChrProcessor::initFlag (HaplotagProcess.cpp:428-436)
    uint8_t *p = bam_aux_get(b, tag);
    if (p) bam_aux_del(b, p);
}

// records of one chunk of a tagging pass, kept until the verdicts come back, next to their SoA form
struct Chunk {
    PackedContig pack;
    std::vector<bam1_t *> records;
    void clear() {
        for (bam1_t *b : records) bam_destroy1(b);
        records.clear();
        pack = PackedContig();
    }
};

// ---- reader / handler pipeline of a tagging pass -------------------------------------------------------------------------
// One thread reads and packs chunk after chunk (htslib record parsing + SoA copies) while the calling thread judges the previous
// chunk on the device, tags its records and writes them; at most `depth` finished chunks wait in between.  Chunks are handled
// strictly in the order they were read, so the output file is the byte stream of the serial loop.
//   read(contig, chunk)   -> 1 chunk filled, 0 contig exhausted, < 0 error      (reader thread only)
//   handle(contig, chunk) -> 0 or < 0 error                                     (calling thread only)
template <class ReadFn, class HandleFn>
int run_chunk_pipeline(int n_contigs, ReadFn read, HandleFn handle, size_t depth = 2) {
    struct Item { int contig; Chunk *chunk; };   // chunk == nullptr: end of stream (contig < 0: after a read error)
    std::mutex m;
    std::condition_variable cv;
    std::vector<Item> q;
    bool stop = false;
    auto push = [&](Item it) {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return q.size() < depth || stop; });
        if (stop && it.chunk) { it.chunk->clear(); delete it.chunk; return; }
        q.push_back(it);
        cv.notify_all();
    };
    std::thread reader([&] {
        for (int i = 0; i < n_contigs; i++) {
            for (;;) {
                { std::lock_guard<std::mutex> lk(m); if (stop) return; }
                Chunk *c = new Chunk();
                const int got = read(i, *c);
                if (got <= 0) { c->clear(); delete c; if (got < 0) { push(Item{-1, nullptr}); return; } break; }
                push(Item{i, c});
            }
        }
        push(Item{0, nullptr});
    });
    int rc = 0;
    for (;;) {
        Item it;
        {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return !q.empty(); });
            it = q.front();
            q.erase(q.begin());
            cv.notify_all();
        }
        if (!it.chunk) { if (it.contig < 0) rc = -1; break; }
        if (rc == 0) rc = handle(it.contig, *it.chunk);
        it.chunk->clear();
        delete it.chunk;
        if (rc != 0) {   // tell the reader to stop and drop what it still delivers
            { std::lock_guard<std::mutex> lk(m); stop = true; for (Item &x : q) if (x.chunk) { x.chunk->clear(); delete x.chunk; } q.clear(); }
            cv.notify_all();
            break;
        }
    }
    reader.join();
    return rc;
}

// BAM region reader that inflates on the device (SURVEY 8f rank 1 wired into the host; LPS_GPU_INFLATE=1): the compressed bytes of the
// region's index chunks are read in one piece, lps_bgzf_scan walks the members, the inflater (lps_bgzf_inflate on device 0, or the
// hook of lpsh_set_inflater) turns them into the uncompressed BAM stream, and the records of `itr`'s region are packed straight
// from those bytes with hts_itr_next's own acceptance test (htslib/hts.c: tid, beg < end of region, end > beg of region).
// 1 = done, 0 = not applicable (the caller uses htslib's reader), < 0 error.
int pack_region_inflated(const std::string &bam_path, const hts_itr_t *itr, PackedContig &pc);
// the same for the tagging passes, which need bam1_t records to tag and write: the inflated stream of a region and its record walk
struct InflatedRegion {
    std::vector<uint8_t> raw;
    uint64_t bytes = 0, at = 0, stop = 0;
    int tid = -1;
    hts_pos_t beg = 0, end = 0;
    bool done = false;
    const uint8_t *next(uint32_t *block_size, bool *error);
    static bool to_bam1(const uint8_t *record, uint32_t block_size, bam1_t *b);
};
int inflate_region(const std::string &bam_path, const hts_itr_t *itr, InflatedRegion &r);
inline bool gpu_inflate_requested() { const char *e = getenv("LPS_GPU_INFLATE"); return e && e[0] == '1'; }

// The alignments overlapping tid:[beg, end) of one BAM read by `readers` threads, each with its own file handle, on equal slices of the
// range; a record belongs to the slice its start lies in (the first slice also takes those that start before beg), so the concatenation
// is the single iterator's sequence (the file is coordinate sorted).  Every reader inflates and parses inline: for a run with fewer
// contigs than threads this is where the spare threads go.  1 = done, 0 = not split (the caller reads the region in one piece), < 0 error.
int pack_region_split(const std::string &bam_path, const std::string &fasta, const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end,
                      int readers, PackedContig &pc);

// Ordered parallel readers (LPS_TAG_READERS=K): the slices of a pass are read by K threads in any order and handled by the calling
// thread strictly in slice order, so the output is again the byte stream of the serial loop; a reader may run at most `ahead` slices
// in front of the handler (bounded memory).
//   read(slice, reader, chunk)  -> 0 or < 0 error      (reader threads; `reader` = 0..K-1 for per-thread file handles)
//   handle(slice, chunk)        -> 0 or < 0 error      (calling thread only, slices 0, 1, 2, ...)
template <class ReadFn, class HandleFn>
int run_ordered_slices(size_t n_slices, int readers, size_t ahead, ReadFn read, HandleFn handle) {
    std::mutex m;
    std::condition_variable cv;
    std::vector<Chunk *> ready(n_slices, nullptr);
    size_t next = 0, handled = 0;
    bool stop = false, failed = false;
    std::vector<std::thread> pool;
    for (int r = 0; r < readers; r++)
        pool.emplace_back([&, r] {
            for (;;) {
                size_t s;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return stop || next >= n_slices || next < handled + ahead; });
                    if (stop || next >= n_slices) return;
                    s = next++;
                }
                Chunk *c = new Chunk();
                const int rc = read(s, r, *c);
                std::lock_guard<std::mutex> lk(m);
                if (rc < 0) { c->clear(); delete c; failed = true; stop = true; }
                else ready[s] = c;
                cv.notify_all();
            }
        });
    int rc = 0;
    for (size_t s = 0; s < n_slices && rc == 0; s++) {
        Chunk *c = nullptr;
        {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return ready[s] != nullptr || failed; });
            if (failed) { rc = -1; break; }
            c = ready[s];
            ready[s] = nullptr;
        }
        rc = handle(s, *c);
        c->clear();
        delete c;
        { std::lock_guard<std::mutex> lk(m); handled = s + 1; }
        cv.notify_all();
    }
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    cv.notify_all();
    for (std::thread &t : pool) t.join();
    for (Chunk *c : ready) if (c) { c->clear(); delete c; }
    return rc;
}

// processSingleChrom's dispatch for a contig without variants (src/haplotag/HaplotagParsingBam.cpp:457-476): MAPQ and flags only, nothing
// reaches the tagger, so no device work exists for these records
inline uint8_t category_without_variants(int mapq, int flag, const lps_tag_params &tp) {
    return mapq < tp.mapping_quality && tp.mapq_filter ? LPS_TAG_LOW_MAPQ : (flag & 0x4) ? LPS_TAG_UNMAPPED : (flag & 0x100) ? LPS_TAG_SECONDARY
           : ((flag & 0x800) && !tp.tag_supplementary) ? LPS_TAG_SUPPLEMENTARY : LPS_TAG_EMPTY_VARIANTS;
}

// The files of a tagging pass (BamFileRAII, src/haplotag/HaplotagParsingBam.cpp:20-82: input with index, output with the @PG line
// `longphase-s` and the same header, one BGZF thread pool for both) and the region in flight with its chunked record reader
// (htslib's iterator, or the batched-inflate stream when LPS_GPU_INFLATE=1).  Shared by `haplotag` and `somatic_haplotag`.
// The tagged-BAM writer with the deflate on the device (the writer side; LPS_GPU_DEFLATE): records are
// serialised the way bam_write1 lays them out (htslib/sam.c:798-864; SAM spec 4.2, long CIGARs as the CG:B,I convention) into a host
// buffer; a full buffer is handed to a flusher thread, which has the library deflate it into BGZF members (lps_bgzf_deflate, one
// dynamic-Huffman block per 65 280 bytes) and appends them to the file while the calling thread goes on tagging.  The file ends
// with htslib's EOF marker.  The uncompressed stream is byte for byte what htslib would have written; the compressed bytes are not.
struct DeviceBamWriter {
    FILE *fp = nullptr;
    std::vector<uint8_t> header, raw, busy_raw, comp;   // the BAM header goes into members of its own, as bam_hdr_write's bgzf_flush leaves it
    bool header_written = false;
    size_t flush_bytes = (size_t)4096 * 0xff00;      // ~267 MB of records per device call: 4096 members, one thread each
    std::thread flusher;
    bool flusher_running = false;
    int flusher_rc = 0;
    lps_ctx *ctx = nullptr;                           // the flusher's own context on device 0 (created on first use)
    double ms_deflate = 0, ms_file = 0;
    uint64_t bytes_in = 0, bytes_out = 0;
    int open(const std::string &path, bam_hdr_t *hdr);
    int write(const bam1_t *b);
    int close();                                      // flushes, writes the EOF marker, closes the file; < 0 on any error so far
  private:
    int hand_over();                                  // waits for the flusher, then gives it `raw`
    int deflate_and_append(const std::vector<uint8_t> &in);
};
// The device writes plain "wb" BAM output (no --cram, no LPS_BAM_LEVEL) when LPS_GPU_DEFLATE=1, never when it is 0, and otherwise
// when the device is the judge of the pass as well (`device_pass`: the product binaries; a pass driven by a caller's own judge -
// the CPU tests - keeps htslib's writer unless asked).
bool gpu_deflate_requested(const std::string &out_mode, bool device_pass);

struct TagBamIO {
    DeviceBamWriter *dev_out = nullptr;                // set instead of `out` when the deflate runs on the device
    bool device_pass = false;                          // the pass is judged on the device (lpsh_tag_run / lpsh_som_run)
    samFile *in = nullptr, *out = nullptr;
    bam_hdr_t *hdr = nullptr;
    hts_idx_t *idx = nullptr;
    htsThreadPool pool = {NULL, 0};
    std::string bam_path;
    int cur = -1;                      // contig whose region is being read
    hts_itr_t *itr = nullptr;
    bool itr_done = false, use_inflated = false;
    InflatedRegion inflated;
    PackedContig::Sizes last_chunk;    // capacity hints for the next chunk
    int open(const std::string &bam, const std::string &fasta, const std::string &out_path, const std::string &out_mode, int threads,
             const std::string &command);
    int start_region(int contig, const std::string &region);
    // appends up to max_records alignments of the region to ck.records / ck.pack: 1 = some, 0 = region exhausted, < 0 error
    int fill(Chunk &ck, size_t max_records);
    void end_region();
    int close();
    // sam_write1 on the output, or the device writer
    int write(bam1_t *b) { return dev_out ? dev_out->write(b) : (sam_write1(out, hdr, b) < 0 ? -1 : 0); }
    bool has_output() const { return out != nullptr || dev_out != nullptr; }
    bool is_open() const { return in != nullptr || out != nullptr || dev_out != nullptr; }
};

// Starts the CUDA driver / context on device 0 in the background (seconds on a box without persistence mode), so that it overlaps
// the VCF / FASTA / first BAM reads; join before the first real lps_ctx_create.
std::thread warm_up_device();

// ---- the text VCF loader of the tag family (VcfParser::parserProcess, src/haplotag/HaplotagVcfParser.cpp:206-545) ---------
// one record of one sample (VarData, HaplotagType.h:110-143)
struct SampleRecord {
    std::string ref, alt;
    int ps = -1;              // VarData::NONE_PHASED_SET
    int gt_kind = 0;          // GenomeType: 1 PHASED_HETERO, 2 UNPHASED_HETERO, 3 UNPHASED_HOMO
    bool hp1_is_alt = false;  // GT 1|0
};
struct SampleVcf {
    std::vector<std::string> chr_names;   // VCF_Info::chrVec (##contig order)
    std::map<std::string, int> chr_length;
    std::map<std::string, std::map<int, SampleRecord>> records;
};
// tumor = false: only phased heterozygous records are kept (NORMAL sample); tumor = true: also 0/1 and 1/1, indels above 100 bp dropped
void load_sample_vcf(const std::string &path, bool tumor, SampleVcf &out);

}  // namespace lpsh
#endif
