"""The C-ABI library loads without a GPU and exports every symbol include/lps.h declares; no compute calls here."""
import ctypes as C
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ffi = importlib.import_module("longphase_s_b200._ffi")


def declared_functions():
    text = open(os.path.join(ROOT, "include", "lps.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lps_[a-z_0-9]+)\s*\(", text)))


def test_header_and_ffi_agree():
    assert declared_functions() == sorted(ffi.SYMBOLS), "include/lps.h and _ffi.SYMBOLS list different entry points"


def test_library_exports_every_declared_symbol():
    lib = ffi.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), f"liblps_b200.so does not export {name}"
    assert b"sm_100a" in lib.lps_version()


def test_struct_layouts_match_the_header():
    # sizes the C compiler produces for the structs of include/lps.h (LP64)
    assert C.sizeof(ffi.LpsCall) == 8
    assert C.sizeof(ffi.LpsVariants) == 72
    assert C.sizeof(ffi.LpsReadBatch) == 208
    assert C.sizeof(ffi.LpsPhaseParams) == 64
    assert ffi.CALL_DTYPE.itemsize == 8


def test_host_library_was_built_against_this_header():
    """liblps_host.so embeds lps_read_batch by value (lpsh_packed): a host library compiled before the struct grew would hand the
    kernels' library uninitialised trailing fields."""
    path = os.path.join(ROOT, "longphase-s_b200", "liblps_host.so")
    if not os.path.exists(path):
        pytest.skip("the C++ host is not built here")
    hostlib = C.CDLL(path)
    assert hasattr(hostlib, "lpsh_sizeof_read_batch"), "stale liblps_host.so: rebuild with make -C longphase-s_b200/host"
    assert hostlib.lpsh_sizeof_read_batch() == C.sizeof(ffi.LpsReadBatch)


def test_no_cpu_fallback():
    """Without a CUDA device the context cannot be created: the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    host = importlib.import_module("longphase_s_b200.host")
    with pytest.raises(host.LpsError):
        host.Context(0)


def test_argument_checks_without_gpu():
    lib = ffi.load_library()
    assert lib.lps_ctx_create(0, None) == -1          # LPS_E_ARG
    assert lib.lps_get_stats(None, None) == -1
    assert lib.lps_last_error(None) == b"null context"


def test_product_does_not_touch_the_oracle():
    """Nothing under longphase-s_b200/ may import, link or execute oracle/ (tests, smoke and bench's baseline legs may)."""
    pkg = os.path.join(ROOT, "longphase-s_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".c", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("oracle's", ""), f"{f} mentions the oracle"
