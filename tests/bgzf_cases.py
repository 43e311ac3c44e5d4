"""BGZF byte streams for the inflation tests: members written here with zlib (every deflate block type, every level) and the
golden BAM written by the reference's own htslib (tests/golden/htslib_small.bam)."""
import os
import struct
import zlib

import numpy as np

GOLDEN_BAM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "htslib_small.bam")


def member(payload, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem_level=8):
    """One BGZF member (htslib/bgzf.c:600-612 layout) around `payload` (<= 65536 bytes)."""
    co = zlib.compressobj(level, zlib.DEFLATED, -15, mem_level, strategy)
    comp = co.compress(payload) + co.flush()
    bsize = 18 + len(comp) + 8 - 1
    assert bsize < 65536
    head = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, ord("B"), ord("C"), 2, bsize)
    return head + comp + struct.pack("<II", zlib.crc32(payload), len(payload))


EOF_MEMBER = member(b"")          # htslib's EOF marker is an empty member


def bam_like(rng, n):
    """Bytes with the texture of BAM records: repetitive small integers (CIGAR), 4-bit sequence, noisy qualities, similar names."""
    parts = []
    while sum(map(len, parts)) < n:
        parts.append(b"read_%08d/ccs\0" % rng.integers(0, 10**8))
        ops = rng.integers(1, 40, 300).astype(np.uint32) << 4 | rng.choice(np.array([0, 0, 0, 1, 2], np.uint32), 300)
        parts.append(ops.tobytes())
        nib = rng.choice(np.array([1, 2, 4, 8], np.uint8), 3000)              # BAM 4-bit bases A C G T, two per byte
        parts.append((nib[0::2] << 4 | nib[1::2]).astype(np.uint8).tobytes())
        parts.append(np.clip(rng.normal(20, 8, 3000), 0, 60).astype(np.uint8).tobytes())
    return b"".join(parts)[:n]


def streams():
    """name -> (bgzf bytes, expected inflated bytes)"""
    rng = np.random.default_rng(9)
    out = {}
    texts = {
        "bam_like": bam_like(rng, 400_000),
        "zeros": bytes(200_000),                                  # one long run: distance-1 matches of length 258
        "random": rng.integers(0, 256, 150_000).astype(np.uint8).tobytes(),   # incompressible: stored blocks at every level
        "text": (b"the quick brown fox jumps over the lazy dog; " * 6000)[:250_000],
        "period3": (b"abc" * 70_000)[:200_000],
        "far_matches": b"".join([rng.integers(0, 256, 30_000).astype(np.uint8).tobytes()] * 2) * 3,   # distances near 30 000
    }
    for name, payload in texts.items():
        for level, strategy, tag in ((1, zlib.Z_DEFAULT_STRATEGY, "l1"), (6, zlib.Z_DEFAULT_STRATEGY, "l6"), (9, zlib.Z_DEFAULT_STRATEGY, "l9"),
                                     (0, zlib.Z_DEFAULT_STRATEGY, "stored"), (6, zlib.Z_FIXED, "fixed"), (6, zlib.Z_HUFFMAN_ONLY, "huff"),
                                     (6, zlib.Z_RLE, "rle")):
            chunk = 65280 if tag != "stored" else 60000
            sizes = [chunk] * (len(payload) // chunk) + [len(payload) % chunk]
            data, o = b"", 0
            for k, sz in enumerate(sizes):
                if k % 5 == 4:
                    data += EOF_MEMBER                            # empty members in the middle are legal
                data += member(payload[o:o + sz], level, strategy)
                o += sz
            out[f"{name}_{tag}"] = (data + EOF_MEMBER, payload)
    # many small members and single bytes
    small = [bytes([i % 256]) * (i % 7 + 1) for i in range(300)]
    out["tiny_members"] = (b"".join(member(x, 6) for x in small) + EOF_MEMBER, b"".join(small))
    return out
