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
