"""BGZF: the host-side member walk (no GPU) and, -m gpu, the device inflation against zlib and against the golden BAM that the
reference's own htslib wrote."""
import ctypes as C
import gzip
import importlib

import numpy as np
import pytest

from tests import bgzf_cases

host = importlib.import_module("longphase_s_b200.host")
ffi = importlib.import_module("longphase_s_b200._ffi")


def test_scan_walks_the_members_of_the_golden_bam():
    data = np.frombuffer(open(bgzf_cases.GOLDEN_BAM, "rb").read(), np.uint8)
    blocks, total = host.bgzf_scan(data)
    want = gzip.decompress(data.tobytes())
    assert total == len(want) and want[:4] == b"BAM\x01"
    assert blocks["out_len"].sum() == total and blocks["out_len"][-1] == 0          # htslib's EOF member
    assert np.array_equal(blocks["out_off"], np.concatenate([[0], np.cumsum(blocks["out_len"])[:-1]]))
    # every table entry really frames one raw deflate stream with the announced size and CRC
    import zlib
    for b in blocks:
        raw = zlib.decompress(data[int(b["comp_off"]):int(b["comp_off"]) + int(b["comp_len"])].tobytes(), -15)
        assert len(raw) == b["out_len"] and zlib.crc32(raw) == b["crc32"]


def test_scan_rejects_malformed_members():
    good = bgzf_cases.member(b"hello world") + bgzf_cases.EOF_MEMBER
    for bad in (good[:-3],                                       # truncated
                good[:12] + b"XX" + good[14:],                   # not the BC subfield
                b"\x1f\x8b\x08\x00" + good[4:],                  # FEXTRA clear
                good[:10] + b"\x08\x00" + good[12:]):            # XLEN != 6 (htslib refuses it too, bgzf.c:880)
        with pytest.raises(host.LpsError):
            host.bgzf_scan(np.frombuffer(bad, np.uint8))
    blocks, total = host.bgzf_scan(np.zeros(0, np.uint8))
    assert len(blocks) == 0 and total == 0


@pytest.mark.gpu
@pytest.mark.parametrize("speculate", ["1", "0"])
def test_inflate_matches_zlib_on_every_block_type(speculate, monkeypatch):
    monkeypatch.setenv("LPS_BGZF_SPECULATE", speculate)      # both decoders of k_bgzf_inflate
    ctx = host.Context(0)
    try:
        for name, (data, want) in bgzf_cases.streams().items():
            got = ctx.bgzf_inflate(data, check_crc=True)
            assert got.tobytes() == want, name
    finally:
        ctx.close()


@pytest.mark.gpu
def test_inflate_chunk_pipeline(monkeypatch):
    """Large inputs are cut into chunks whose upload, kernel and download overlap on three streams; here with tiny chunks."""
    monkeypatch.setenv("LPS_BGZF_CHUNK", "150000")
    ctx = host.Context(0)
    try:
        for name in ("bam_like_l6", "zeros_l9", "random_stored", "far_matches_l6", "tiny_members"):
            data, want = bgzf_cases.streams()[name]
            assert ctx.bgzf_inflate(data, check_crc=True).tobytes() == want, name
    finally:
        ctx.close()


@pytest.mark.gpu
def test_inflate_golden_bam_from_htslib():
    data = open(bgzf_cases.GOLDEN_BAM, "rb").read()
    ctx = host.Context(0)
    try:
        assert ctx.bgzf_inflate(data).tobytes() == gzip.decompress(data)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_inflate_reports_corrupt_blocks():
    payload = bgzf_cases.bam_like(np.random.default_rng(2), 60_000)
    good = bytearray(bgzf_cases.member(payload) + bgzf_cases.EOF_MEMBER)
    ctx = host.Context(0)
    try:
        bad_crc = bytearray(good)
        bad_crc[len(bgzf_cases.member(payload)) - 8] ^= 0xFF
        with pytest.raises(host.LpsError, match="CRC32"):
            ctx.bgzf_inflate(bytes(bad_crc), check_crc=True)
        assert ctx.bgzf_inflate(bytes(bad_crc), check_crc=False).tobytes() == payload
        rng = np.random.default_rng(4)
        failures = 0
        for _ in range(40):                                      # flipped bits inside the deflate stream: an error or a CRC mismatch, never a hang
            bad = bytearray(good)
            bad[18 + int(rng.integers(0, len(good) - 60))] ^= 1 << int(rng.integers(0, 8))
            try:
                ctx.bgzf_inflate(bytes(bad), check_crc=True)
            except host.LpsError:
                failures += 1
        assert failures == 40
    finally:
        ctx.close()
