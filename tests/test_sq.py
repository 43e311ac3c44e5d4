"""The interleaved SEQ + QUAL rows (include/lps.h, lps_read_batch.sq): the host packers and lps_sq_peek - which runs the index
arithmetic of the kernel's gather compiled for the host - against a numpy restatement of BAM's two arrays; no GPU."""
import importlib

import numpy as np

from tests import cases

ffi = importlib.import_module("longphase_s_b200._ffi")


def rows_numpy(seq4, qual, lq):
    """The row of one read, written down independently: 16-byte units of ten qualities, five nibble bytes and a zero."""
    n_units = (lq + 9) // 10
    out = np.zeros(n_units * 16, np.uint8)
    for u in range(n_units):
        nb = min(10, lq - 10 * u)
        out[16 * u:16 * u + nb] = qual[10 * u:10 * u + nb]
        s0 = 5 * u
        ns = min(5, (lq + 1) // 2 - s0)
        out[16 * u + 10:16 * u + 10 + ns] = seq4[s0:s0 + ns]
    return out


def test_row_bytes():
    lib = ffi.load_library()
    assert [lib.lps_sq_row_bytes(n) for n in (-3, 0, 1, 9, 10, 11, 20, 21, 100000)] == [0, 0, 16, 16, 16, 32, 32, 48, 160000]


def test_pack_and_peek_every_index():
    lib = ffi.load_library()
    rng = np.random.default_rng(7)
    for lq in (1, 2, 9, 10, 11, 19, 20, 21, 29, 30, 31, 255, 1000, 1001, 4099):
        codes = rng.integers(0, 16, lq).astype(np.uint8)
        qual = np.ascontiguousarray(rng.integers(0, 94, lq).astype(np.uint8))
        padded = np.concatenate([codes, np.zeros(lq & 1, np.uint8)])
        seq4 = np.ascontiguousarray((padded[0::2] << 4) | padded[1::2])          # BAM: even index in the high nibble
        row = np.full(int(lib.lps_sq_row_bytes(lq)) + 16, 0xAA, np.uint8)          # a guard unit behind the row
        assert lib.lps_pack_sq(ffi.ptr(seq4, ffi.u8p), ffi.ptr(qual, ffi.u8p), lq, ffi.ptr(row, ffi.u8p)) == 0
        assert (row[-16:] == 0xAA).all(), "lps_pack_sq wrote past the row"
        assert np.array_equal(row[:-16], rows_numpy(seq4, qual, lq)), f"row of a {lq}-base read differs"
        code, q = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
        for qi in range(lq):
            assert lib.lps_sq_peek(ffi.ptr(row, ffi.u8p), lq, qi, ffi.ptr(code, ffi.u8p), ffi.ptr(q, ffi.u8p)) == 0
            assert (int(code[0]), int(q[0])) == (int(codes[qi]), int(qual[qi])), (lq, qi)
        assert lib.lps_sq_peek(ffi.ptr(row, ffi.u8p), lq, lq, ffi.ptr(code, ffi.u8p), ffi.ptr(q, ffi.u8p)) == -1
        assert lib.lps_sq_peek(ffi.ptr(row, ffi.u8p), lq, -1, ffi.ptr(code, ffi.u8p), ffi.ptr(q, ffi.u8p)) == -1


def test_batch_packer_on_a_contig():
    """Contig.pack_sq (threads over lps_pack_sq_batch) against the two arrays of the generator, read by read."""
    lib = ffi.load_library()
    contig, _ = cases.get("dense_indel_noseq")      # reads without SEQ (l_qseq = 0) included
    sq, sq_off = contig.pack_sq(threads=3)
    sq1, sq_off1 = contig.pack_sq(threads=1)
    assert np.array_equal(sq, sq1) and np.array_equal(sq_off, sq_off1)
    assert (sq_off % 16 == 0).all()
    rng = np.random.default_rng(3)
    code, q = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    for r in rng.choice(contig.n_reads, min(contig.n_reads, 40), replace=False):
        lq, so, qo = int(contig.l_qseq[r]), int(contig.seq_off[r]), int(contig.qual_off[r])
        if lq <= 0:
            continue
        row = sq[int(sq_off[r]):int(sq_off[r]) + int(lib.lps_sq_row_bytes(lq))]
        assert np.array_equal(row, rows_numpy(contig.seq4[so:so + (lq + 1) // 2], contig.qual[qo:qo + lq], lq))
        for qi in rng.integers(0, lq, 50):
            qi = int(qi)
            assert lib.lps_sq_peek(ffi.ptr(np.ascontiguousarray(row), ffi.u8p), lq, qi, ffi.ptr(code, ffi.u8p), ffi.ptr(q, ffi.u8p)) == 0
            want = (int(contig.seq4[so + (qi >> 1)]) >> ((~qi & 1) << 2)) & 15
            assert (int(code[0]), int(q[0])) == (want, int(contig.qual[qo + qi]))
    # rows tile the stream without gaps
    rows = np.array([int(lib.lps_sq_row_bytes(int(x))) for x in contig.l_qseq])
    assert np.array_equal(sq_off[1:], np.cumsum(rows)[:-1].astype(np.uint64))


def test_wire_format_is_pinned():
    """The rows of a fixed synthetic contig hash to a committed value: a change of the layout (unit size, byte order of the
    qualities, nibble packing) has to be made on purpose, in the kernel, the packer and here."""
    import hashlib
    contig, _ = cases.get("snp_only")
    sq, sq_off = contig.pack_sq(threads=1)
    assert len(sq) == 24242016
    assert hashlib.sha256(sq.tobytes() + sq_off.tobytes()).hexdigest() == "9d51025d11666b78a55d4c246e7a51c07a95e595a62e73cd119a40d72f589cc7"
