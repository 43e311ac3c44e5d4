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
