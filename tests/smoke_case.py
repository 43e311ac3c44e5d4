"""__graft_entry__.smoke(): one small contig through the CUDA hot path on cuda:0, checked against the oracle."""
import importlib

import __graft_entry__ as entry


def run_smoke():
    entry.load_package()
    synth = importlib.import_module("longphase_s_b200.synth")
    ffi = importlib.import_module("longphase_s_b200._ffi")
    from . import parity
    contig = synth.Contig(seed=5, contig_len=300_000, indel_frac=0.1)
    info = parity.check_phase(contig, ffi.default_phase_params(True))
    print("smoke ok:", {k: v for k, v in info.items() if k != "stats"}, "kernel launches", info["stats"]["kernel_launches"])
