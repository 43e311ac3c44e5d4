"""BGZF deflation (include/lps.h, lps_bgzf_deflate): the member encoder the kernel runs, compiled for the host, against zlib's inflate
and gzip's own CRC / ISIZE checks (no GPU); -m gpu: the kernel's stream must be byte for byte what the host-compiled encoder writes,
and this repository's device inflater must read it back."""
import ctypes as C
import gzip
import importlib
import struct
import zlib

import numpy as np
import pytest

from tests import bgzf_cases

host = importlib.import_module("longphase_s_b200.host")
ffi = importlib.import_module("longphase_s_b200._ffi")
SLOT = 65312


def member_host(payload):
    lib = ffi.load_library()
    a = np.frombuffer(payload, np.uint8) if len(payload) else np.zeros(1, np.uint8)
    out = np.full(SLOT + 64, 0xA5, np.uint8)
    n = C.c_uint32(0)
    assert lib.lps_bgzf_deflate_block_host(ffi.ptr(np.ascontiguousarray(a), ffi.u8p), len(payload), ffi.ptr(out, ffi.u8p), SLOT, C.byref(n)) == 0
    assert (out[SLOT:] == 0xA5).all(), "the encoder wrote past its slot"
    return out[:n.value].tobytes()


def check_member(m, payload):
    assert m[:16] == bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0])
    assert struct.unpack("<H", m[16:18])[0] == len(m) - 1                         # BSIZE
    assert struct.unpack("<II", m[-8:]) == (zlib.crc32(payload), len(payload))    # CRC32, ISIZE
    d = zlib.decompressobj(-15)
    raw = d.decompress(m[18:-8]) + d.flush()
    assert d.eof and d.unused_data == b"", "the deflate stream does not end where the member's trailer starts"
    assert raw == payload
    assert gzip.decompress(m) == payload                                          # a gzip reader checks CRC32 and ISIZE itself
    assert len(m) <= len(payload) + 31


def payloads():
    rng = np.random.default_rng(5)
    fib = [1, 1]
    while sum(fib) < 60000:
        fib.append(fib[-1] + fib[-2])
    skew = np.concatenate([np.full(f, k, np.uint8) for k, f in enumerate(fib)])    # Fibonacci counts: a Huffman tree deeper than 15
    rng.shuffle(skew)
    geo = np.minimum(rng.geometric(0.5, 65280) - 1, 255).astype(np.uint8)           # counts halving per symbol: depth ~ 16
    return {
        "empty": b"",
        "one_byte": b"A",
        "one_symbol": b"\0" * 65280,                                                 # one literal + end of block: two codes of one bit
        "two_symbols": bytes([7, 9] * 1000),
        "all_256_once": bytes(range(256)),
        "random_full": rng.integers(0, 256, 65280).astype(np.uint8).tobytes(),       # incompressible: a stored block
        "random_small": rng.integers(0, 256, 40).astype(np.uint8).tobytes(),
        "bam_like_full": bgzf_cases.bam_like(rng, 65280),
        "bam_like_tail": bgzf_cases.bam_like(rng, 12345),
        "text": (b"the quick brown fox jumps over the lazy dog; " * 2000)[:65280],
        "fibonacci": skew.tobytes()[:65280],
        "geometric": geo.tobytes(),
        "qualities": np.clip(rng.normal(20, 8, 65280), 0, 60).astype(np.uint8).tobytes(),
    }


def test_member_encoder_round_trips_through_zlib():
    for name, p in payloads().items():
        m = member_host(p)
        check_member(m, p)
    assert len(member_host(payloads()["random_full"])) == 18 + 5 + 65280 + 8        # stored
    assert len(member_host(payloads()["one_symbol"])) < 8400                          # one bit per byte


def test_code_lengths_respect_the_limits_of_the_format():
    """The dynamic header is parsed back here (RFC 1951 3.2.7): code lengths <= 15 (literals) and <= 7 (code-length code), both codes
    complete - for the inputs that force deep Huffman trees too."""
    for name in ("fibonacci", "geometric", "bam_like_full", "one_symbol", "two_symbols"):
        m = member_host(payloads()[name])
        bits = np.unpackbits(np.frombuffer(m[18:-8], np.uint8), bitorder="little")
        at = 0

        def take(n):
            nonlocal at
            v = int(sum(int(b) << k for k, b in enumerate(bits[at:at + n])))
            at += n
            return v
        assert take(1) == 1 and take(2) == 2, name                                    # BFINAL, dynamic
        hlit, hdist, hclen = take(5) + 257, take(5) + 1, take(4) + 4
        assert (hlit, hdist, hclen) == (257, 2, 19)
        perm = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
        cl = [0] * 19
        for k in range(hclen):
            cl[perm[k]] = take(3)
        assert max(cl) <= 7 and sum(2.0 ** -l for l in cl if l) == 1.0, (name, cl)
        assert cl[16] == cl[17] == cl[18] == 0                                        # no run-length symbols are used
        # canonical decode of the code-length code
        codes = {}
        code = 0
        for length in range(1, 8):
            for s in range(19):
                if cl[s] == length:
                    codes[(length, code)] = s
                    code += 1
            code <<= 1
        lens = []
        while len(lens) < hlit + hdist:
            c, n = 0, 0
            while True:
                c = (c << 1) | take(1)
                n += 1
                if (n, c) in codes:
                    lens.append(codes[(n, c)])
                    break
                assert n < 8
        lit, dist = lens[:hlit], lens[hlit:]
        assert max(lit) <= 15 and dist == [1, 1], name
        assert sum(2.0 ** -l for l in lit if l) == 1.0, (name, "literal code is not complete")
        assert lit[256] > 0


def test_size_against_zlib_on_bam_like_bytes():
    rng = np.random.default_rng(11)
    data = bgzf_cases.bam_like(rng, 20 * 65280)
    ours = sum(len(member_host(data[i:i + 65280])) for i in range(0, len(data), 65280))
    z6 = sum(len(bgzf_cases.member(data[i:i + 65280])) for i in range(0, len(data), 65280))
    zh = sum(len(bgzf_cases.member(data[i:i + 65280], strategy=zlib.Z_HUFFMAN_ONLY)) for i in range(0, len(data), 65280))
    assert ours <= 1.01 * zh, (ours, zh)          # zlib's own Huffman-only mode run-length codes its header; nothing else differs
    assert ours <= 1.10 * z6, (ours, z6)


def test_bound_and_argument_checks():
    lib = ffi.load_library()
    assert lib.lps_bgzf_deflate_bound(0, 65280) == 0
    assert lib.lps_bgzf_deflate_bound(65280, 65280) == 65280 + 31
    assert lib.lps_bgzf_deflate_bound(65281, 65280) == 65281 + 62
    assert lib.lps_bgzf_deflate_bound(10, 0) == 0 and lib.lps_bgzf_deflate_bound(10, 65281) == 0
    out = np.zeros(SLOT, np.uint8)
    n = C.c_uint32(0)
    assert lib.lps_bgzf_deflate_block_host(ffi.ptr(out, ffi.u8p), 65281, ffi.ptr(out, ffi.u8p), SLOT, C.byref(n)) == -1
    assert lib.lps_bgzf_deflate_block_host(ffi.ptr(out, ffi.u8p), 10, ffi.ptr(out, ffi.u8p), SLOT - 1, C.byref(n)) == -1


@pytest.mark.gpu
def test_gpu_deflate_matches_the_host_compiled_encoder_and_inflates():
    rng = np.random.default_rng(21)
    parts = [bgzf_cases.bam_like(rng, 40 * 65280 + 777), rng.integers(0, 256, 3 * 65280).astype(np.uint8).tobytes(), bytes(2 * 65280 + 5),
             payloads()["fibonacci"], payloads()["geometric"], b"tail"]
    data = b"".join(parts)
    ctx = host.Context(0)
    try:
        for block_bytes in (65280, 4096):
            comp = ctx.bgzf_deflate(np.frombuffer(data, np.uint8), block_bytes)
            want = b"".join(member_host(data[i:i + block_bytes]) for i in range(0, len(data), block_bytes))
            assert comp.tobytes() == want, f"kernel and host-compiled encoder disagree (block_bytes {block_bytes})"
            assert gzip.decompress(comp.tobytes()) == data                         # concatenated members, every CRC checked
            assert ctx.bgzf_inflate(comp, check_crc=True).tobytes() == data        # and back through k_bgzf_inflate
        assert len(ctx.bgzf_deflate(np.zeros(0, np.uint8), 65280)) == 0
    finally:
        ctx.close()


def test_member_bytes_are_pinned():
    """The encoder is deterministic and its choices (tie order of the Huffman merge, header without run-length symbols) are part of
    what the kernel must reproduce: the member of a fixed input hashes to a committed value."""
    import hashlib
    m = member_host(payloads()["bam_like_tail"])
    assert len(m) == 8482 and hashlib.sha256(m).hexdigest() == "522ba0bf9c02ddcfdfe93f0b0bf6c52b3834f1752c863f16ce6dd39694a082f1"
