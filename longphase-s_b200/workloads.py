"""The synthetic workloads bench.py times, defined ONCE so that the parity tests (`tests/test_gpu_large.py`), the committed
known-answer digests (`tests/golden/bench_digests.json`, written by `tools/make_bench_digests.py` from the CPU checker) and the bench
itself build byte-identical contigs.  BASELINE.json configs: C2 shard (phase SNP+indel, 64 Mb contigs, 30x ONT-like 20 kb reads), C4 shard (tumor 50x /
normal 25x pair), and the whole-genome shape of C2 (24 GRCh38-proportioned contigs, strong scaling by LPT over the ranks).

No torch / CUDA imports: host logic, usable from the CPU tests.
"""
import hashlib
import json
import os

import numpy as np

from . import shard

HERE = os.path.dirname(os.path.abspath(__file__))
DIGEST_FILE = os.path.join(os.path.dirname(HERE), "tests", "golden", "bench_digests.json")


def phase_kwargs(seed, contig_mb=64.0, depth=30.0, mean_len=20000.0, variant_spacing=1000.0, scale=1.0):
    """synth.Contig keyword arguments of one phase contig of the bench."""
    return dict(seed=int(seed), contig_len=int(contig_mb * 1_000_000 * scale), indel_frac=0.1, depth=float(depth), mean_len=float(mean_len),
                variant_rate=1.0 / float(variant_spacing))


def weak_seed(rank, i):
    """Contig i of rank `rank` in the weak-scaling run (every rank phases its own equal contigs)."""
    return 100 + 16 * int(rank) + int(i)


def genome_contigs(genome_mb):
    """The 24 GRCh38-proportioned contigs scaled to `genome_mb` megabases in total: [(name, seed, contig_mb)], in karyotype order."""
    total = sum(shard.GRCH38_MB.values())
    return [(name, 2000 + k, mb * genome_mb / total) for k, (name, mb) in enumerate(shard.GRCH38_MB.items())]


def genome_partition(genome_mb, world):
    """LPT partition (shard.lpt_partition) of the genome's contigs over `world` ranks by expected read count (proportional to the
    contig length at constant depth; a real run takes the counts from the BAM index).  Returns `world` lists of (name, seed, mb)."""
    contigs = genome_contigs(genome_mb)
    by_name = {c[0]: c for c in contigs}
    bins = shard.lpt_partition({c[0]: c[2] for c in contigs}, world)
    return [[by_name[n] for n in b] for b in bins]


def c4_pair_kwargs(contig_mb=32.0, depth=30.0, mean_len=20000.0, variant_spacing=1000.0):
    """(normal kwargs, tumor kwargs) of the C4 shard: tumor 50x (purity 0.6) / normal 25x over one contig, ~3000 somatic SNV+indel per 64 Mb."""
    kw = phase_kwargs(900, contig_mb, depth, mean_len, variant_spacing)
    kw.update(somatic_rate=3000.0 / 64e6, indel_frac=0.1)
    kn, kt = dict(kw), dict(kw)
    kn.update(depth=25.0, purity=0.0, read_seed=901)
    kt.update(depth=50.0, purity=0.6, read_seed=902)
    return kn, kt


def c4_pair(synth_mod, contig_mb=32.0, **kw):
    """(union map with the normal reads, union map with the tumor reads) of the C4 shard."""
    kn, kt = c4_pair_kwargs(contig_mb, **kw)
    cn, ct = synth_mod.Contig(**kn), synth_mod.Contig(**kt)
    un = cn.somatic_union(seed=9)
    return un, un.with_reads_of(ct)


def key_of(kwargs):
    """Stable text key of a synth configuration (the digest file is keyed by it)."""
    return json.dumps({k: kwargs[k] for k in sorted(kwargs)}, separators=(",", ":"))


def phase_digest(ps, hap_ref, read_hp, hp_counts):
    """sha256 over the final products of the phase path of one contig (per-variant PS and REF haplotype, per-read haplotype,
    per-variant hp x allele counters), in their C ABI types."""
    h = hashlib.sha256()
    for a, dt in ((ps, np.int32), (hap_ref, np.int8), (read_hp, np.int8), (hp_counts, np.int32)):
        h.update(np.ascontiguousarray(np.asarray(a), dtype=dt).tobytes())
    return h.hexdigest()


def load_digests():
    if not os.path.exists(DIGEST_FILE):
        return {}
    with open(DIGEST_FILE) as f:
        return json.load(f)
