"""bench.py's host-side helpers that only matter with several ranks on one box (no GPU here): the NUMA binding decision."""
import builtins
import io
import os
import types

import bench


class FakeCuda:
    def __init__(self, n):
        self.n = n

    def device_count(self):
        return self.n

    def get_device_properties(self, d):
        return types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x10 + d, pci_device_id=0)


def run_bind(monkeypatch, affinity, cpulists, local_rank, world, per_rank):
    """cpulists[d] = (local_cpulist, numa_node) of GPU d"""
    real_open = builtins.open

    def fake_open(path, *a, **k):
        if isinstance(path, str) and path.startswith("/sys/bus/pci/devices/"):
            d = int(path.split(":")[1], 16) - 0x10
            return io.StringIO(cpulists[d][0] + "\n" if path.endswith("local_cpulist") else cpulists[d][1] + "\n")
        return real_open(path, *a, **k)
    bound = {}
    monkeypatch.setattr(builtins, "open", fake_open)
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(affinity))
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: bound.setdefault("cpus", set(cpus)))
    text = bench.bind_to_gpu_numa(types.SimpleNamespace(cuda=FakeCuda(len(cpulists))), local_rank, world, per_rank)
    return bound.get("cpus"), text


def test_cpu_list_parser():
    assert bench.cpus_of_list("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert bench.cpus_of_list("\n") == set()


def test_binds_when_every_rank_of_the_node_keeps_its_share(monkeypatch):
    two_nodes = [("0-15", "0")] * 4 + [("16-31", "1")] * 4
    cpus, text = run_bind(monkeypatch, range(32), two_nodes, 5, 8, 4)
    assert cpus == set(range(16, 32)) and "node 1" in text and "4 rank(s)" in text


def test_leaves_the_mask_alone_when_the_node_is_short_of_cores(monkeypatch):
    two_nodes = [("0-23", "0")] * 4 + [("24-31", "1")] * 4          # 8 local cores for 4 ranks that want 4 each
    cpus, text = run_bind(monkeypatch, range(32), two_nodes, 6, 8, 4)
    assert cpus is None and text.startswith("unchanged")
    # all cores of the container on one node: nothing to gain for anybody
    one_node = [("0-31", "0")] * 4 + [("32-63", "1")] * 4
    for r in (0, 7):
        cpus, text = run_bind(monkeypatch, range(32), one_node, r, 8, 4)
        assert cpus is None and text.startswith("unchanged")


def test_missing_sysfs_is_not_an_error(monkeypatch):
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(8)))

    class Broken:
        def device_count(self):
            return 2

        def get_device_properties(self, d):
            raise RuntimeError("no CUDA")
    assert bench.bind_to_gpu_numa(types.SimpleNamespace(cuda=Broken()), 0, 2, 4).startswith("unchanged")


def test_deflate_leg_runs_against_a_stand_in_context():
    """bench.py's BGZF deflate leg end to end, with the host-compiled member encoder standing in for the device call: the leg's own
    checks (members inflate to the input, sizes against zlib) and the shape of what it reports."""
    import ctypes as C
    import importlib

    import numpy as np
    bench.entry.load_package()
    ffi = importlib.import_module("longphase_s_b200._ffi")
    lib = ffi.load_library()

    class StandIn:
        def bgzf_deflate(self, data, block_bytes=0xff00):
            out, slot, n = [], np.zeros(65312, np.uint8), C.c_uint32(0)
            for i in range(0, len(data), block_bytes):
                piece = np.ascontiguousarray(data[i:i + block_bytes])
                assert lib.lps_bgzf_deflate_block_host(ffi.ptr(piece, ffi.u8p), len(piece), ffi.ptr(slot, ffi.u8p), 65312, C.byref(n)) == 0
                out.append(slot[:n.value].copy())
            return np.concatenate(out) if out else np.zeros(0, np.uint8)

        def stats(self):
            return {"ms_kernel_bgzf": 2.0}
    r = bench.bench_bgzf_deflate(StandIn(), 2, types.SimpleNamespace(steps=1), 6534.8, mb=8)
    assert 0.5 < r["ratio"] < 0.8 and r["ratio"] < 1.1 * r["cpu_zlib_level6"]["ratio"]
    assert r["roofline"]["kernel"] == "k_bgzf_deflate" and r["roofline"]["algorithmic_bytes_per_launch"] > 2 * (8 << 20)
