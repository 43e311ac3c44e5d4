"""lps_pack_cigar16 (host side of the compact CIGAR wire format, include/lps.h) against a numpy restatement; no GPU."""
import ctypes as C
import importlib

import numpy as np

from tests import cases

ffi = importlib.import_module("longphase_s_b200._ffi")


def pack(cigar, base=0, cap=None, n_long0=0):
    lib = ffi.load_library()
    n = len(cigar)
    want = int(((cigar >> 4) >= 0xFFF).sum())
    cap = n_long0 + want if cap is None else cap
    c16 = np.zeros(max(n, 1), np.uint16)
    ll, la = np.zeros(max(cap, 1), np.uint32), np.zeros(max(cap, 1), np.uint64)
    nl = C.c_uint64(n_long0)
    rc = lib.lps_pack_cigar16(ffi.ptr(cigar, ffi.u32p), n, base, ffi.ptr(c16, ffi.u16p), ffi.ptr(ll, ffi.u32p), ffi.ptr(la, ffi.u64p),
                              cap, C.byref(nl))
    return rc, c16[:n], ll, la, nl.value


def test_pack_matches_numpy():
    rng = np.random.default_rng(1)
    n = 100_000
    length = rng.integers(1, 60, n).astype(np.uint32)
    big = rng.random(n) < 0.001
    length[big] = rng.integers(4090, 1 << 28, int(big.sum()))
    length[:4] = [4094, 4095, 4096, (1 << 28) - 1]
    cigar = np.ascontiguousarray((length << 4) | rng.integers(0, 10, n).astype(np.uint32))
    rc, c16, ll, la, nl = pack(cigar, base=7)
    assert rc == 0
    is_long = length >= 4095
    assert nl == int(is_long.sum())
    assert np.array_equal(c16[~is_long], cigar[~is_long].astype(np.uint16))
    assert np.array_equal(c16[is_long], (0xFFF0 | (cigar[is_long] & 15)).astype(np.uint16))
    assert np.array_equal(ll[:nl], length[is_long]) and np.array_equal(la[:nl], np.flatnonzero(is_long).astype(np.uint64) + 7)
    # widening + patching (what the device does on arrival) gives the stream back
    back = c16.astype(np.uint32)
    back[(la[:nl] - 7).astype(np.int64)] = (ll[:nl] << 4) | (back[(la[:nl] - 7).astype(np.int64)] & 15)
    assert np.array_equal(back, cigar)


def test_pack_appends_and_reports_a_full_table():
    cigar = np.array([(5000 << 4) | 2, (10 << 4) | 0, (70000 << 4) | 4], np.uint32)
    rc, c16, ll, la, nl = pack(cigar, base=100, n_long0=3, cap=5)
    assert rc == 0 and nl == 5 and list(ll[3:5]) == [5000, 70000] and list(la[3:5]) == [100, 102]
    rc, *_ = pack(cigar, cap=1)
    assert rc < 0
    rc, c16, *_ , nl = pack(np.zeros(0, np.uint32))
    assert rc == 0 and nl == 0


def test_synthetic_contig_round_trip():
    contig, _ = cases.get("snp_indel")
    c16, ll, la = contig.pack_cigar16()
    assert len(ll) == 0 and np.array_equal(c16.astype(np.uint32), contig.cigar)


# ---- the 8-bit wire format (lps_pack_cigar8) -------------------------------------------------------------------------------------
def expand8(c8, esc16, n):
    """numpy restatement of k_expand_cigar8: the 16-bit stream from the 8-bit one."""
    b = c8.astype(np.uint32)
    out = np.zeros(n, np.uint32)
    m, i, d, e = b < 0x80, (b >= 0x80) & (b < 0xB8), (b >= 0xB8) & (b < 0xF0), b == 0xFF
    out[m] = ((b[m] + 1) << 4) | 0
    out[i] = ((b[i] - 0x80 + 1) << 4) | 1
    out[d] = ((b[d] - 0xB8 + 1) << 4) | 2
    out[e] = esc16[:int(e.sum())]
    assert (m | i | d | e).all()
    return out.astype(np.uint16)


def test_pack8_round_trip_and_block_table():
    rng = np.random.default_rng(3)
    n = 70_000
    op = rng.choice([0, 0, 0, 0, 1, 2, 2, 3, 4, 5, 7, 8], n).astype(np.uint32)
    length = rng.integers(1, 40, n).astype(np.uint32)
    length[rng.random(n) < 0.05] = rng.integers(50, 300, int((rng.random(n) < 0.05).sum()) or 1)[0]
    big = rng.random(n) < 0.001
    length[big] = rng.integers(4090, 1 << 27, int(big.sum()))
    length[:6] = [128, 129, 56, 57, 4095, 1]
    op[:6] = [0, 0, 1, 2, 0, 0]
    cigar = np.ascontiguousarray((length << 4) | op)

    class C_:
        pass
    c = C_()
    c.cigar = cigar
    synth = importlib.import_module("longphase_s_b200.synth")
    c8, esc16, esc_blk, long_len, long_at = synth.Contig.pack_cigar8(c)
    # the 16-bit stream the device would rebuild equals what lps_pack_cigar16 makes of the same ops
    rc, c16, ll, la, nl = pack(cigar)
    assert rc == 0
    assert np.array_equal(expand8(c8, esc16, n), c16)
    assert np.array_equal(long_len, ll[:nl]) and np.array_equal(long_at, la[:nl])
    is_esc = (c8 == 0xFF)
    want_blk = np.concatenate([[0], np.cumsum(is_esc)])[np.minimum(np.arange(n // 256 + 2) * 256, n)]
    assert np.array_equal(esc_blk[:(n + 255) // 256 + 1], want_blk[:(n + 255) // 256 + 1])
    assert c8[0] == 127 and c8[1] == 0xFF and c8[2] == 0x80 + 55 and c8[3] == 0xFF and c8[4] == 0xFF and c8[5] == 0


def test_pack8_appends_across_calls():
    lib = ffi.load_library()
    rng = np.random.default_rng(4)
    n = 3000
    cigar = np.ascontiguousarray(((rng.integers(1, 200, n).astype(np.uint32)) << 4) | rng.choice([0, 1, 2, 4], n).astype(np.uint32))
    c8 = np.zeros(n, np.uint8); esc16 = np.zeros(n, np.uint16); esc_blk = np.zeros(n // 256 + 2, np.uint32)
    ll = np.zeros(4, np.uint32); la = np.zeros(4, np.uint64)
    ne, nl = C.c_uint64(0), C.c_uint64(0)
    at = 0
    for size in (1, 255, 256, 700, n - 1212):
        part = np.ascontiguousarray(cigar[at:at + size])
        rc = lib.lps_pack_cigar8(ffi.ptr(part, ffi.u32p), size, at, ffi.ptr(c8[at:], ffi.u8p), ffi.ptr(esc16, ffi.u16p), n, C.byref(ne),
                                 ffi.ptr(esc_blk, ffi.u32p), ffi.ptr(ll, ffi.u32p), ffi.ptr(la, ffi.u64p), 4, C.byref(nl))
        assert rc == 0
        at += size
    assert at == n

    class C_:
        pass
    c = C_()
    c.cigar = cigar
    synth = importlib.import_module("longphase_s_b200.synth")
    one = synth.Contig.pack_cigar8(c)
    assert np.array_equal(one[0], c8) and np.array_equal(one[1], esc16[:ne.value]) and np.array_equal(one[2], esc_blk)
