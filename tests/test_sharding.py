"""Multi-GPU path on CPU: contig sharding (LPT, no collective on the data path) and the host-side merge, with two `gloo`
processes standing in for two ranks.  The per-contig work is done by the oracle here (the checker), because the product
kernels need a GPU; what is under test is the partition, the rank-local loop and the merge."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from . import cases  # noqa: F401  (loads the package)

shard = importlib.import_module("longphase_s_b200.shard")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONTIGS = {"ctgA": dict(seed=31, contig_len=200_000), "ctgB": dict(seed=32, contig_len=120_000, indel_frac=0.1),
           "ctgC": dict(seed=33, contig_len=90_000), "ctgD": dict(seed=34, contig_len=60_000, indel_frac=0.2),
           "ctgE": dict(seed=35, contig_len=40_000)}


def test_lpt_partition_balances_a_genome():
    w = shard.GRCH38_MB
    for n in (1, 2, 4, 8):
        bins = shard.lpt_partition(w, n)
        assert sorted(k for b in bins for k in b) == sorted(w)
        loads = shard.bin_loads(w, bins)
        # the heaviest rank bounds the step: with 24 human contigs LPT stays within 6 % of the ideal split up to 8 ranks
        assert max(loads) <= 1.06 * sum(loads) / n, (n, loads)
    assert shard.lpt_partition([5, 5, 5], 2) == [[0, 2], [1]]
    assert shard.lpt_partition({}, 3) == [[], [], []]


def _phase_one(name):
    from oracle import pyoracle as po
    synth = importlib.import_module("longphase_s_b200.synth")
    ffi = importlib.import_module("longphase_s_b200._ffi")
    c = synth.Contig(**CONTIGS[name])
    o = po.OraclePhase(c, ffi.default_phase_params(True))
    return shard.export_phasing_result(name, c.var_pos, o.ps, o.hap_ref), dict(reads=c.n_reads, calls=len(o.calls))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    weights = {k: v["contig_len"] for k, v in CONTIGS.items()}
    mine = shard.contigs_of_rank(weights, world, rank)
    local = [_phase_one(n) for n in mine]                                   # rank-local loop, no communication
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, local))                         # result merge only (host objects)
    dist.barrier()
    if rank == 0:
        names = [n for g in gathered for n in g[0]]
        merged = shard.merge_phasing_results([r for g in gathered for r, _ in g[1]])
        stats = shard.merge_read_statistics([s for g in gathered for _, s in g[1]])
        q.put((names, merged, stats))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_merge_to_the_single_process_result():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    names, merged, stats = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(names) == sorted(CONTIGS)
    single = [_phase_one(n) for n in CONTIGS]
    assert merged == shard.merge_phasing_results([r for r, _ in single])
    assert stats == shard.merge_read_statistics([s for _, s in single])
    assert len(merged) > 100 and all(k.split("_")[0] in CONTIGS for k in merged)
