"""The host half of lps_phase_solve (lps_sweep_votes = edgeConnectResult over one-byte edge summaries) against the oracle's
sweep over the float edge table, on every SIMD path.  No GPU: the vote bytes are derived here, in numpy, from the oracle's own
weights with the rules of VariantEdge::findBestEdgePair (reference src/phase/PhasingGraph.cpp:166-228)."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

from tests import cases

ffi = importlib.import_module("longphase_s_b200._ffi")
pyoracle = importlib.import_module("oracle.pyoracle")


def vote_bytes(weights, edge_threshold):
    """[n][W][4] float32 (rr, ra, ar, aa) -> [n][W] summaries; float32 sums, double ratio, like the reference."""
    rr, ra, ar, aa = (weights[..., i] for i in range(4))
    para, cross = rr + aa, ar + ra                                 # float32 adds
    with np.errstate(invalid="ignore", divide="ignore"):
        esr = np.minimum(para, cross).astype(np.float64) / np.maximum(para, cross).astype(np.float64)
    link = np.where(para > cross, 1, np.where(para < cross, 2, 0))
    link = np.where(esr > edge_threshold, 0, link)                 # NaN (empty cell) compares false: the link stays 0 anyway
    info = link.astype(np.uint8)
    heavy = ((esr <= 0.1) & ((rr + aa + ra + ar) >= 1)) | ((para < 1) & (cross >= 1)) | ((para >= 1) & (cross < 1))
    info |= np.where(heavy, 4, 0).astype(np.uint8)
    info |= np.where((para + cross) <= 1, 8, 0).astype(np.uint8)
    info |= np.where(esr < 0.2, 16, 0).astype(np.uint8)
    n, w = info.shape
    k, d = np.meshgrid(np.arange(n), np.arange(w), indexing="ij")
    info[k + 1 + d >= n] = 0                                       # successors past the last node
    return np.ascontiguousarray(info)


def run_sweep(params, node_pos, node_type, votes, mode):
    lib = ffi.load_library()
    n, w = votes.shape
    ps, hap = np.zeros(n, np.int32), np.zeros(n, np.int8)
    old = os.environ.get("LPS_SWEEP")
    os.environ["LPS_SWEEP"] = mode
    try:
        rc = lib.lps_sweep_votes(C.byref(params), n, w, ffi.ptr(node_pos, ffi.i32p), ffi.ptr(node_type, ffi.u8p), ffi.ptr(votes, ffi.u8p),
                                 ffi.ptr(ps, ffi.i32p), ffi.ptr(hap, ffi.i8p))
    finally:
        if old is None:
            del os.environ["LPS_SWEEP"]
        else:
            os.environ["LPS_SWEEP"] = old
    assert rc >= 0
    return rc, ps, hap


@pytest.mark.parametrize("name", ["snp_indel", "dense_indel_noseq", "deep_long_reads", "short_reads_sparse"])
def test_sweep_votes_matches_oracle(name):
    contig, params = cases.get(name)
    orc = pyoracle.OraclePhase(contig, params)
    votes = vote_bytes(orc.weights, params.edge_threshold)
    node_pos = np.ascontiguousarray(contig.var_pos[orc.node_var].astype(np.int32))
    node_type = np.ascontiguousarray(orc.node_type)
    want_ps, want_hap = orc.ps_sweep[orc.node_var], orc.hap_ref_sweep[orc.node_var]
    paths = set()
    for mode in ("scalar", "avx2", "avx512"):
        path, ps, hap = run_sweep(params, node_pos, node_type, votes, mode)
        paths.add(path)
        assert np.array_equal(ps, want_ps), (name, mode)
        assert np.array_equal(hap, want_hap), (name, mode)
    assert 0 in paths


def test_sweep_votes_small_windows_and_edges():
    """Window sizes around the 16-node blocks, a distance break and the N < 2 cases; every SIMD path must agree with the scalar one."""
    rng = np.random.default_rng(5)
    params = ffi.default_phase_params(True)
    for w in (1, 2, 15, 16, 17, 33, 35, 48, 49, 100, 127):
        for n in (0, 1, 2, 3, 17, 400):
            votes = np.zeros((n, w), np.uint8)
            if n:
                link = rng.integers(0, 3, (n, w)).astype(np.uint8)
                link[rng.random((n, w)) < 0.3] = 0
                votes = (link | (rng.integers(0, 8, (n, w)).astype(np.uint8) << 2)).astype(np.uint8)
                k, d = np.meshgrid(np.arange(n), np.arange(w), indexing="ij")
                votes[k + 1 + d >= n] = 0
            node_pos = np.cumsum(rng.integers(1, 2000, n)).astype(np.int32)
            if n > 200:
                node_pos[200:] += 400_000          # farther than params.distance: the chain restarts
            node_type = rng.choice(np.array([0, 0, 0, 3, 4], np.uint8), n)
            _, ps0, hap0 = run_sweep(params, node_pos, node_type, np.ascontiguousarray(votes), "scalar")
            for mode in ("avx2", "avx512"):
                _, ps, hap = run_sweep(params, node_pos, node_type, np.ascontiguousarray(votes), mode)
                assert np.array_equal(ps, ps0) and np.array_equal(hap, hap0), (w, n, mode)


def test_sweep_votes_rejects_bad_arguments():
    lib = ffi.load_library()
    params = ffi.default_phase_params(True)
    z32, z8, zi8 = np.zeros(4, np.int32), np.zeros(4 * 200, np.uint8), np.zeros(4, np.int8)
    args = (ffi.ptr(z32, ffi.i32p), ffi.ptr(z8, ffi.u8p), ffi.ptr(z8, ffi.u8p), ffi.ptr(z32, ffi.i32p), ffi.ptr(zi8, ffi.i8p))
    assert lib.lps_sweep_votes(C.byref(params), 4, 0, *args) < 0
    assert lib.lps_sweep_votes(C.byref(params), 4, 128, *args) < 0
    assert lib.lps_sweep_votes(C.byref(params), -1, 35, *args) < 0
