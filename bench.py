#!/usr/bin/env python
"""bench.py — throughput of the LongPhase-S `phase` read-to-variant hot path on B200.

One "step" = one pass of the whole hot path (allele calling -> overlap filter -> graph construction + edge fold ->
edgeConnectResult sweep -> read correction) over the contigs of this GPU.

Workloads (longphase-s_b200/workloads.py; BASELINE.json config C2, "phase SNP+indel whole-genome ... contig-sharded 1/2/4/8 B200"):
  --workload genome (default): 24 GRCh38-proportioned contigs, --genome-mb megabases in total, 30x ONT-like 20 kb reads.  The contigs
      are dealt to the ranks by a greedy LPT partition of their read counts (shard.lpt_partition), so the total work is FIXED as the
      number of GPUs grows: "scaling": "strong".  The whole 3.1 Gb genome does not fit one GPU next to its scratch (165 GB of reads),
      so the default is the half-scale genome (1536 Mb, chr1 = 124 Mb ... chr21 = 23 Mb): the largest single-GPU configuration.
  --workload weak: --contigs-per-gpu equal contigs of --contig-mb per GPU (round 1's shape), "scaling": "weak".
  value : reads/s with every read batch already resident in HBM (lps_batch_submit_device, 16-bit CIGAR stream)
  e2e   : reads/s through the C ABI with pinned HOST buffers, H2D + D2H inside the timed region
  roofline / rooflines : the dominant kernel (k_call_alleles) and the other kernels against the measured HBM copy peak
  cpu_baseline / --impl reference : the unmodified reference (oracle/_ref) on the host cores
Before a value is printed the phase result of EVERY contig of every rank is digested (workloads.phase_digest) and compared with
the committed digest of the CPU checker for that contig (tests/golden/bench_digests.json); a mismatch aborts the run.
Launch: `python bench.py --gpus 1` or torchrun for N > 1 (one rank per GPU, contigs are independent: no collective on the data
path; NCCL is only used for the barrier and the max-over-ranks of the time).
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "phase_hot_path_reads_per_s"


def cpus_of_list(text):
    out = set()
    for part in text.strip().split(","):
        if part:
            a, _, b = part.partition("-")
            out.update(range(int(a), int(b or a) + 1))
    return out


def bind_to_gpu_numa(torch, local_rank, world, cores_per_rank):
    """Several ranks on one box: keep this rank's host threads on the cores next to its GPU (sysfs local_cpulist of the GPU's PCI
    function), so that the pages they touch first - the pinned buffers the GPU reads over PCIe - live on that NUMA node.  Only when
    those cores, shared with the other ranks whose GPUs sit on the same node, still leave every rank `cores_per_rank` of them;
    otherwise nothing changes.  Returns what was done."""
    try:
        def local_cpus(dev):
            pr = torch.cuda.get_device_properties(dev)
            bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            return (cpus_of_list(open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read()),
                    open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        local, node = local_cpus(local_rank)
        sharing = sum(1 for d in range(min(world, torch.cuda.device_count())) if local_cpus(d)[0] == local)
        mine = os.sched_getaffinity(0)
        both = mine & local
        if len(both) >= max(2, cores_per_rank) * max(1, sharing) and len(both) < len(mine):
            os.sched_setaffinity(0, both)
            return "node %s: %d of %d cores, shared by %d rank(s)" % (node, len(both), len(mine), sharing)
        return "unchanged (node %s, %d local cores among this process's %d, %d rank(s) on the node)" % (node, len(both), len(mine), sharing)
    except Exception as e:          # no sysfs entry, no permission: the run goes on unbound
        return "unchanged (%s)" % type(e).__name__


def workload_text(args, world, n_contigs_rank0):
    if args.workload == "genome":
        return ("C2 whole-genome shape at %.0f Mb: phase SNP+indel over 24 GRCh38-proportioned contigs (%.1f .. %.1f Mb), %gx ONT-like %g kb reads, "
                "1 het variant / %g bp (10%% indels), ONT error model; contigs dealt to %d GPU(s) by LPT on read counts (%d on rank 0)"
                % (args.genome_mb, args.genome_mb * 46.71 / 3088.27, args.genome_mb * 248.96 / 3088.27, args.depth, args.mean_len / 1e3,
                   args.variant_spacing, world, n_contigs_rank0))
    return ("C2 shard: phase SNP+indel, %d x %g Mb contigs per GPU (%g Mb), %gx ONT-like %g kb reads, 1 het variant / %g bp (10%% indels), "
            "ONT error model" % (args.contigs_per_gpu, args.contig_mb, args.contigs_per_gpu * args.contig_mb, args.depth, args.mean_len / 1e3,
                                 args.variant_spacing))


def algorithmic_bytes_k1(contig, status, n_calls):
    """B1 = sum_reads (16 + 4 n_cigar) + 18 n_calls; reads rejected by the flag/MAPQ filter
    are never walked, so only their 16-byte record counts."""
    walked = status != 2
    return int(16 * contig.n_reads + 4 * int(contig.n_cigar[walked].sum()) + 18 * n_calls)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.stop_flag, self.th = device, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i",
                                      str(self.device)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def run_reference_cpu(args, n_threads, rounds, warmup):
    """The UNMODIFIED reference (oracle/_ref/libref_tap.so: BamParser::get_snp over every record, SnpParser::filterSNP, Clip,
    VairiantGraph::addEdge, edgeConnectResult, readCorrection, exportResult) on one contig of the workload (the contig the GPU arm
    phases first in the weak shape), every host thread running its own replica - the reference parallelises over contigs the same
    way (PhasingProcess.cpp:113).  The time of a replica is the sum of the seconds spent INSIDE the reference's own calls (the tap's
    timing mode: nothing is flattened or copied for the caller); a round's throughput is replicas x reads / mean replica time.
    Falls back to the C restatement under oracle/ when the reference library is absent."""
    from oracle import pyoracle as po
    synth = importlib.import_module("longphase_s_b200.synth")
    ffi = importlib.import_module("longphase_s_b200._ffi")
    wl = importlib.import_module("longphase_s_b200.workloads")
    contig = synth.Contig(**wl.phase_kwargs(wl.weak_seed(0, 0), args.contig_mb, args.depth, args.mean_len, args.variant_spacing))
    params = ffi.default_phase_params(True)
    kind = "reference" if po.tap_available() else "port"

    def one_round():
        secs, calls = [0.0] * n_threads, [0] * n_threads

        def work(k):
            if kind == "reference":
                r = po.ReferencePhaseTimed(contig, params)
                secs[k], calls[k] = r.seconds, r.n_calls
            else:
                t0 = time.perf_counter()
                r = po.OraclePhase(contig, params)
                secs[k], calls[k] = time.perf_counter() - t0, len(r.calls)
        ths = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return float(np.mean(secs)), time.perf_counter() - t0, calls[0]

    for _ in range(warmup):
        one_round()
    res = [one_round() for _ in range(max(1, rounds))]
    per = [contig.n_reads * n_threads / r[0] for r in res]
    value = float(np.mean(per))
    mean_s = float(np.mean([r[0] for r in res]))
    spread = 100.0 * (max(per) - min(per)) / value if value else 0.0
    return dict(value=value, unit="reads/s", cores=n_threads, kind=kind,
                sample="%d replicas of one %g Mb contig of the workload (%d reads each), %.1f s inside the reference's calls per replica, %d round(s), "
                       "spread %.1f%%" % (n_threads, args.contig_mb, contig.n_reads, mean_s, len(res), spread),
                allele_calls_per_s=res[0][2] * n_threads / mean_s, ms_per_step=1e3 * float(np.mean([r[1] for r in res])), spread_pct=spread)


def bench_bgzf(ctx, host, ffi, torch, ncores, args, mb=1024):
    """inflation of BGZF blocks (htslib bgzf_read_block / inflate_block).  BAM-like synthetic bytes, deflated
    by zlib level 6 in 65280-byte members like htslib writes them; kernel-resident, end-to-end (host in, host out) and zlib on all
    host cores (what the reference's htslib build calls) on the same members."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    from tests import bgzf_cases
    rng = np.random.default_rng(1)
    unit = bgzf_cases.bam_like(rng, 8 << 20)
    chunks = [unit[o:o + 65280] for o in range(0, len(unit), 65280)]
    with ThreadPoolExecutor(max_workers=ncores) as tp:
        members = list(tp.map(bgzf_cases.member, chunks))
    reps = max(1, (mb << 20) // len(unit))
    data = np.frombuffer(b"".join(members) * reps + bgzf_cases.EOF_MEMBER, np.uint8)
    blocks, out_bytes = host.bgzf_scan(data)
    d_data, d_blocks = torch.from_numpy(data.copy()).cuda(), torch.from_numpy(blocks.view(np.uint8).copy()).cuda()
    d_out = torch.empty(out_bytes + 16, dtype=torch.uint8, device="cuda")
    ms = []
    for _ in range(args.warmup + args.steps):
        rc = ctx.lib.lps_bgzf_inflate_device(ctx.h, d_data.data_ptr(), d_blocks.data_ptr(), len(blocks), d_out.data_ptr())
        if rc != 0:
            raise RuntimeError(ctx.lib.lps_last_error(ctx.h).decode())
        ms.append(ctx.stats()["ms_kernel_bgzf"])
    k_ms = float(np.mean(ms[args.warmup:]))
    got = d_out[:len(unit)].cpu().numpy().tobytes()
    assert got == unit, "device inflation differs from the input text"
    pin_in = torch.from_numpy(data.copy()).pin_memory()
    pin_out = torch.empty(out_bytes + 16, dtype=torch.uint8).pin_memory()
    pblocks = blocks.ctypes.data_as(C.POINTER(ffi.LpsBgzfBlock))
    e2e = []
    for _ in range(2 + args.steps):
        t0 = time.perf_counter()
        rc = ctx.lib.lps_bgzf_inflate(ctx.h, C.cast(pin_in.data_ptr(), ffi.u8p), len(data), pblocks, len(blocks), C.cast(pin_out.data_ptr(), ffi.u8p),
                                      out_bytes, 0)
        if rc != 0:
            raise RuntimeError(ctx.lib.lps_last_error(ctx.h).decode())
        e2e.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = float(np.mean(e2e[2:]))
    raw = [m[18:-8] for m in members]

    def cpu_pass(_):
        n = 0
        for r in raw:
            n += len(zlib.decompress(r, -15))
        return n
    with ThreadPoolExecutor(max_workers=ncores) as tp:
        t0 = time.perf_counter()
        done = sum(tp.map(cpu_pass, range(ncores)))
        cpu_s = time.perf_counter() - t0
    alg = len(data) + out_bytes
    return {"config": "%d MB of BAM-like bytes in %d BGZF members (zlib level 6, 65280 bytes each), ratio %.2f" % (out_bytes >> 20, len(blocks), out_bytes / len(data)),
            "kernel_ms": k_ms, "out_gb_per_s": out_bytes / (k_ms * 1e-3) / 1e9, "algorithmic_bytes": alg,
            "e2e_ms": e2e_ms, "e2e_out_gb_per_s": out_bytes / (e2e_ms * 1e-3) / 1e9, "h2d_bytes": int(len(data)), "d2h_bytes": int(out_bytes),
            "cpu_zlib_out_gb_per_s": done / cpu_s / 1e9, "cpu_cores": ncores, "cpu_sample": "%d MB inflated per core, python zlib (libz inflate, GIL released)" % (len(unit) >> 20)}


def bench_bgzf_deflate(ctx, ncores, args, peak, mb=256):
    """SURVEY 8f rank 1, the writer side: deflation of BAM-like bytes into BGZF members (htslib bgzf_write / deflate_block).  Kernel
    time and the time of the whole call (pageable host buffers in and out), size and speed of zlib on all host cores beside it, and a
    round trip of the first members through zlib (CRC32 and ISIZE checked by the gzip reader)."""
    import gzip
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    from tests import bgzf_cases
    rng = np.random.default_rng(3)
    tile = bgzf_cases.bam_like(rng, 8 << 20)
    data = np.frombuffer((tile * (mb // 8 + 1))[:mb << 20], np.uint8).copy()
    data[::4099] ^= rng.integers(0, 255, len(data[::4099])).astype(np.uint8)      # the tiles are not identical
    ctx.bgzf_deflate(data[:1 << 20])
    wall, kern, comp = [], [], None
    for _ in range(max(1, min(args.steps, 3))):
        t0 = time.perf_counter()
        comp = ctx.bgzf_deflate(data)
        wall.append((time.perf_counter() - t0) * 1e3)
        kern.append(ctx.stats()["ms_kernel_bgzf"])
    buf, pos, members = comp.tobytes(), 0, 0
    while pos < len(buf) and members < 32:
        pos += int.from_bytes(buf[pos + 16:pos + 18], "little") + 1
        members += 1
    if gzip.decompress(buf[:pos]) != data[:members * 65280].tobytes():
        raise RuntimeError("the deflated members do not inflate to the input")
    sample = data[:min(len(data), 64 << 20)].tobytes()

    def one(i):
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        return len(co.compress(sample[i:i + 65280]) + co.flush()) + 26
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=ncores) as tp:
        z_bytes = sum(tp.map(one, range(0, len(sample), 65280)))
    z_s = time.perf_counter() - t0
    k, w = float(np.mean(kern)), float(np.mean(wall))
    alg = 2 * len(data) + 2 * len(comp)
    return {"config": "%d MB of BAM-like bytes, %d members of 65280 bytes, one dynamic-Huffman block of literals each" % (mb, (len(data) + 65279) // 65280),
            "kernel_ms": k, "in_gb_per_s": len(data) / (k * 1e-3) / 1e9, "call_ms_host_buffers": w, "call_in_gb_per_s": len(data) / (w * 1e-3) / 1e9,
            "ratio": len(comp) / len(data), "cpu_zlib_level6": {"ratio": z_bytes / len(sample), "in_gb_per_s": len(sample) / z_s / 1e9, "cores": ncores},
            "roofline": {"bound": "hbm", "kernel": "k_bgzf_deflate", "achieved": alg / (k * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (k * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": int(alg), "kernel_ms": k,
                         "note": "two passes over the input, the slots written once, one copy into the stream; one thread per member: the time is "
                                 "a member's latency chain through its local-memory tables, not bytes"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="genome", choices=["genome", "weak"])
    ap.add_argument("--genome-mb", type=float, default=1536.0)
    ap.add_argument("--contig-mb", type=float, default=64.0)            # weak shape; also the contig of the CPU reference arm
    ap.add_argument("--contigs-per-gpu", type=int, default=8)          # weak shape
    ap.add_argument("--threads", type=int, default=0)                  # host threads per rank driving the contigs; 0: min(contigs, cores / ranks, 8)
    ap.add_argument("--depth", type=float, default=30.0)               # C5 stress: --depth 120 --mean-len 50000 --variant-spacing 300
    ap.add_argument("--mean-len", type=float, default=20000.0)
    ap.add_argument("--variant-spacing", type=float, default=1000.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-paths", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sq", action="store_true")            # end-to-end leg: SEQ and QUAL as BAM's two arrays instead of the interleaved rows (lps_read_batch.sq)
    ap.add_argument("--no-numa", action="store_true")          # several ranks: do not bind the host threads to the GPU's NUMA node
    ap.add_argument("--e2e-ab", action="store_true")           # end-to-end leg: time the other SEQ / QUAL wire format as well (reported beside the headline one)
    ap.add_argument("--resident-sq", action="store_true")      # resident leg: the interleaved rows in HBM instead of the two arrays (experiment; the line says so)
    ap.add_argument("--sync", default="auto", choices=["auto", "spin", "block", "yield"])
    ap.add_argument("--cigar16", action="store_true")    # end-to-end leg: send the 16-bit CIGAR stream instead of the 8-bit wire format
    ap.add_argument("--cigar32", action="store_true")    # end-to-end leg: send BAM's uint32 CIGAR ops instead of the compact 16-bit stream
    ap.add_argument("--allow-unpinned", action="store_true")   # run contigs that have no committed digest (custom shapes); the line says so
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # torchrun exports OMP_NUM_THREADS=1; the synthetic generator (OpenMP) gets this rank's share of the host cores.  Must be
    # set before libgomp is loaded (torch pulls it in).
    os.environ["OMP_NUM_THREADS"] = str(max(1, ncores // world))
    # stdout carries exactly ONE line (the JSON); anything libraries print there (NCCL's version banner) goes to stderr
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    entry.load_package()
    wl = importlib.import_module("longphase_s_b200.workloads")

    if args.impl == "reference":
        if rank != 0:
            return 0
        os.environ["OMP_NUM_THREADS"] = str(ncores)
        res = run_reference_cpu(args, min(ncores, 64), args.steps, min(args.warmup, 1))
        n0 = len(wl.genome_partition(args.genome_mb, max(1, args.gpus))[0]) if args.workload == "genome" else args.contigs_per_gpu
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": "reads/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "strong" if args.workload == "genome" else "weak", "vs_baseline": None, "dtype": "int32/f32",
                "data": "synthetic", "allele_calls_per_s": res["allele_calls_per_s"],
                "config": {"workload": workload_text(args, max(1, args.gpus), n0), "sample": res["sample"]},
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    ffi = importlib.import_module("longphase_s_b200._ffi")
    # ---- this rank's contigs ----
    if args.workload == "genome":
        mine = wl.genome_partition(args.genome_mb, world)[rank]          # [(name, seed, mb)], heaviest first
        specs = [(name, wl.phase_kwargs(seed, mb, args.depth, args.mean_len, args.variant_spacing)) for name, seed, mb in mine]
    else:
        specs = [("w%d_%d" % (rank, i), wl.phase_kwargs(wl.weak_seed(rank, i), args.contig_mb, args.depth, args.mean_len, args.variant_spacing))
                 for i in range(args.contigs_per_gpu)]
    n_ctg = len(specs)
    T_ = args.threads if args.threads > 0 else max(1, min(n_ctg, max(1, ncores // world), 8))
    # waiting host threads spin while each has a core of its own, otherwise they yield it
    if args.sync == "auto":
        args.sync = "yield" if ncores // world < T_ else "spin"
    blocking = args.sync == "block"
    if blocking or args.sync == "yield":
        ffi.load_library().lps_set_blocking_sync(local_rank, 1 if blocking else 2)
    numa = bind_to_gpu_numa(torch, local_rank, world, ncores // world) if world > 1 and not args.no_numa else "not applicable"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    synth_mod = importlib.import_module("longphase_s_b200.synth")
    host = importlib.import_module("longphase_s_b200.host")
    from concurrent.futures import ThreadPoolExecutor

    t_gen = time.time()
    contigs = [synth_mod.Contig(**kw) for _, kw in specs]
    t_gen = time.time() - t_gen
    digests = wl.load_digests()
    want = [digests.get(wl.key_of(kw)) for _, kw in specs]
    if any(w is None for w in want) and not args.allow_unpinned:
        raise SystemExit("bench.py: no committed digest for %s; run tools/make_bench_digests.py for this shape or pass --allow-unpinned"
                         % [n for (n, _), w in zip(specs, want) if w is None])
    params = ffi.default_phase_params(True)
    # one context (own stream, own scratch, resident reference + variant table) per contig, like a job that keeps every contig of
    # its shard on the device; T_ host threads drive them, each its own share, heaviest first (the reference runs its contig loop
    # the same way: `#pragma omp parallel for schedule(dynamic)`, PhasingProcess.cpp:113)
    ctxs = [host.Context(local_rank) for _ in range(n_ctg)]
    keep = []
    for ctx, contig in zip(ctxs, contigs):
        ctx.set_reference(contig.ref)
        vs = contig.variants_struct()
        keep.append(vs)
        ctx.set_variants(vs, True)
    share = [list(range(t, n_ctg, T_)) for t in range(T_)]

    # ---- device-resident copy of every batch (torch owns the memory): the 16-bit CIGAR stream + its side table, SEQ, QUAL, records ----
    def dev(a):
        view = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64}.get(a.dtype)
        return torch.from_numpy(a.view(view) if view else a).cuda()

    names = ["ref_start", "l_qseq", "n_cigar", "cigar_off", "seq_off", "qual_off", "flag", "mapq", "name_rank", "seq4", "qual"]
    ptypes = dict(ref_start=ffi.i32p, l_qseq=ffi.i32p, n_cigar=ffi.u32p, cigar_off=ffi.u64p, seq_off=ffi.u64p, qual_off=ffi.u64p,
                  flag=ffi.u16p, mapq=ffi.u8p, name_rank=ffi.i32p, cigar=ffi.u32p, seq4=ffi.u8p, qual=ffi.u8p)
    packed = [c.pack_cigar16() for c in contigs]

    def batch_from(contig, ptr_of, pk):
        b = ffi.LpsReadBatch(n_reads=contig.n_reads, cigar_len=len(contig.cigar), seq_bytes=len(contig.seq4), qual_bytes=len(contig.qual),
                             **{k: C.cast(ptr_of(k), ptypes[k]) for k in names})
        b.cigar16 = C.cast(ptr_of("cigar16"), ffi.u16p)
        b.n_cigar_long = len(pk[1])
        if len(pk[1]):
            b.cigar_long_len = C.cast(ptr_of("cigar_long_len"), ffi.u32p)
            b.cigar_long_at = C.cast(ptr_of("cigar_long_at"), ffi.u64p)
        return b

    dtens, dev_batches = [], []
    for c, pk in zip(contigs, packed):
        if args.resident_sq:
            d = {k: dev(getattr(c, k)) for k in names if k not in ("seq4", "qual", "qual_off", "seq_off")}
            sq, sq_off = c.pack_sq(threads=min(ncores, 16))
            d["sq"], d["seq_off"] = dev(sq), dev(sq_off)
            del sq
        else:
            d = {k: dev(getattr(c, k)) for k in names}
        d["cigar16"] = dev(pk[0])
        d["cigar_long_len"] = dev(np.ascontiguousarray(pk[1])) if len(pk[1]) else None
        d["cigar_long_at"] = dev(np.ascontiguousarray(pk[2])) if len(pk[2]) else None
        dtens.append(d)
        b = batch_from(c, lambda k, d=d: d[k].data_ptr() if d.get(k) is not None else None, pk)
        if args.resident_sq:
            b.sq, b.sq_bytes, b.seq_bytes, b.qual_bytes = C.cast(d["sq"].data_ptr(), ffi.u8p), d["sq"].numel(), 0, 0
        dev_batches.append(b)
    resident_bytes = int(sum(t.numel() * t.element_size() for d in dtens for t in d.values() if t is not None))
    n_reads_gpu = int(sum(c.n_reads for c in contigs))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    max_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.MAX if world > 1 else None)  # noqa: E731
    sum_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.SUM if world > 1 else None)  # noqa: E731
    min_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.MIN if world > 1 else None)  # noqa: E731

    pool = ThreadPoolExecutor(max_workers=T_)
    results = [None] * n_ctg

    def run_threads(fn, steps):
        """Every host thread runs `steps` passes over its share of the contigs back to back (no barrier between steps or contigs: the
        host parts of one contig overlap the kernels of the others); the last pass keeps the results."""
        def body(t):
            for k in range(steps):
                for i in share[t]:
                    r = fn(i, k == steps - 1)
                    if r is not None:
                        results[i] = r
        for f in [pool.submit(body, t) for t in range(T_)]:
            f.result()

    def step_resident(i, last):
        return ctxs[i].phase_contig(params, copy=last)      # the batch was registered once with lps_batch_submit_device (no copy)

    def timed(fn, steps, slot):
        """device time between the events of every context's stream (the streams the kernels run on), max over them"""
        barrier()
        for ctx in ctxs:
            ctx.event_record(slot)
        t0 = time.perf_counter()
        run_threads(fn, steps)
        for ctx in ctxs:
            ctx.event_record(slot + 1)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        return max(ctx.event_elapsed_ms(slot, slot + 1) for ctx in ctxs) / steps, wall / steps

    # ---- kernel-resident leg ----
    for i in range(n_ctg):
        ctxs[i].submit_device(dev_batches[i])
    if args.warmup:
        run_threads(step_resident, args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()    # the JSON line is rank 0's: only its GPU is sampled (every nvidia-smi call costs host CPU)
    s0 = [ctx.stats() for ctx in ctxs]
    dev_ms, wall_ms = timed(step_resident, args.steps, 0)
    s1 = [ctx.stats() for ctx in ctxs]
    clocks = None                 # the sampler keeps running through the end-to-end leg (the resident leg alone lasts ~0.1 s)
    ms_step = max_over_ranks(dev_ms)
    launches = int(sum(b_["kernel_launches"] - a_["kernel_launches"] for a_, b_ in zip(s0, s1)))
    stage_keys = ("ms_call_alleles", "ms_build_edges", "ms_sweep", "ms_read_correction", "ms_wall_call_alleles", "ms_wall_build_edges",
                  "ms_host_filters", "ms_host_sweep", "ms_kernel_call_alleles", "ms_kernel_fold_edges")
    stage = {k[3:]: float(np.mean([st[k] for st in s1])) for k in stage_keys}
    stage["sweep_fallbacks"] = int(sum(b_["sweep_fallbacks"] - a_["sweep_fallbacks"] for a_, b_ in zip(s0, s1)))
    stage["slow_path_contigs"] = int(sum(b_["slow_path_contigs"] - a_["slow_path_contigs"] for a_, b_ in zip(s0, s1)))
    res_resident = list(results)

    # ---- parity gate: the digest of every contig's result against the committed digest of the CPU checker ----
    def digest_ok(res_list):
        ok, checked = True, 0
        for r, w in zip(res_list, want):
            if w is None:
                continue
            checked += 1
            ok = ok and wl.phase_digest(r["ps"], r["hap_ref"], r["read_hp"], r["hp_counts"]) == w["digest"]
        return ok, checked
    ok_resident, n_checked = digest_ok(res_resident)
    debug_no_gate = os.environ.get("LPS_BENCH_DEBUG_NO_GATE") == "1"     # timing experiments with a deliberately wrong kernel build: the line is marked invalid
    if min_over_ranks(1.0 if ok_resident else 0.0) < 1.0 and not debug_no_gate:
        raise SystemExit("bench.py: the phase result of a timed contig differs from the committed digest of the CPU checker: no value is reported")

    # ---- the kernels timed alone (one context, nothing else on the GPU): rooflines ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy; the kernel is timed alone)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    big = int(np.argmax([c.n_reads for c in contigs]))
    ctx0, contig0 = ctxs[big], contigs[big]
    calls = ctx0.call_alleles(params, want_host=True)
    n_calls0 = int(calls["n_calls"])
    b1 = algorithmic_bytes_k1(contig0, calls["read_status"], n_calls0)
    k1_ms, fold_ms = [], []
    for _ in range(max(args.steps, 5)):
        ctx0.call_alleles(params, want_host=False)
        k1_ms.append(ctx0.stats()["ms_kernel_call_alleles"])
    edges = None
    for _ in range(3):
        ctx0.call_alleles(params, want_host=False)
        edges = ctx0.build_edges(params, want_host=False)
        fold_ms.append(ctx0.stats()["ms_kernel_fold_edges"])
    k1, kf = float(np.mean(k1_ms)), float(np.mean(fold_ms))
    # 8d: B2 = 24 n_contrib + 16 n_cells, cells = node x successor-in-window pairs the fold writes (4 floats each)
    n_cells = int(edges["n_nodes"]) * int(edges["window"])
    b2 = 24 * int(edges["n_contrib"]) + 16 * n_cells
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_k1_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))                     # a list of captured launches; the one of this contig (same read count) counts
        for e in (tj if isinstance(tj, list) else [tj]):
            if e.get("reads") == contig0.n_reads:
                traffic, traffic_src = e["dram_bytes_per_launch"], e["source"]
    achieved = b1 / (k1 * 1e-3) / 1e9 if k1 > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_call_alleles", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": b1, "kernel_ms": k1,
                "peak_source": peak_src + " (of measured)", "launch": "the largest contig of rank 0 (%d reads, %d allele calls), timed alone" % (contig0.n_reads, n_calls0),
                "note": "algorithmic bytes follow SURVEY 8d (4 bytes per CIGAR op); the kernel reads the stream in 16 bits per op, so its "
                        "DRAM traffic is below the algorithmic figure, and it is bound by instruction issue / latency, not by HBM (profiles/)"}
    rooflines = [dict(roofline), {"bound": "hbm", "kernel": "k_fold_edges", "achieved": b2 / (kf * 1e-3) / 1e9 if kf > 0 else 0.0, "peak": peak, "unit": "GB/s",
                                   "frac": (b2 / (kf * 1e-3) / 1e9 / peak) if kf > 0 else 0.0, "algorithmic_bytes_per_launch": b2, "kernel_ms": kf,
                                   "note": "24 B per pair contribution + 16 B per (node, successor) cell; the ordered float fold is bound by issue, not by HBM"}]
    ctxs[big].submit_device(dev_batches[big])

    # ---- end-to-end leg: pinned host buffers through the C ABI ----
    e2e = e2e_other = None
    if not args.no_e2e:
        # The host loop packs the CIGAR ops of every record into the 8-bit wire format while it appends the record to the batch
        # (lps_pack_cigar8; include/lps.h); the device expands it into the resident 16-bit stream.  The big streams are pinned in place (cudaHostRegister), the small ones copied.
        cudart = torch.cuda.cudart()

        def run_e2e(use_sq):
            pinned_in_place, ptens, pin_batches = [], [], []

            def pin(a):
                a = np.ascontiguousarray(a)
                if a.nbytes >= (1 << 20):
                    rc = cudart.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
                    if int(rc) == 0:
                        pinned_in_place.append(a)
                        return a, a.ctypes.data
                t = torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory()
                return t, t.data_ptr()
            for c, pk in zip(contigs, packed):
                d, ptr = {}, {}
                for k in names:
                    if use_sq and k in ("seq4", "qual", "qual_off", "seq_off"):
                        continue
                    d[k], ptr[k] = pin(getattr(c, k))
                if use_sq:
                    # the host loop writes a record's bases and qualities as one interleaved row (lps_pack_sq) while it appends the record
                    sq, sq_off = c.pack_sq(threads=min(ncores, 16))
                    d["sq"], ptr["sq"] = pin(sq)
                    d["seq_off"], ptr["seq_off"] = pin(sq_off)
                    del sq
                    if not args.e2e_ab:
                        # nothing reads the generator's two host arrays from here on (the resident copy lives in HBM): keeps the
                        # host footprint at one copy of the bases and qualities
                        c.seq4, c.qual = c.seq4[:0].copy(), c.qual[:0].copy()
                if args.cigar32:
                    d["cigar"], ptr["cigar"] = pin(c.cigar)
                    b = ffi.LpsReadBatch(n_reads=c.n_reads, cigar_len=len(c.cigar), seq_bytes=len(c.seq4), qual_bytes=len(c.qual),
                                         **{k: C.cast(ptr[k], ptypes[k]) for k in names + ["cigar"]})
                elif args.cigar16:
                    d["cigar16"], ptr["cigar16"] = pin(pk[0])
                    if len(pk[1]):
                        d["cigar_long_len"], ptr["cigar_long_len"] = pin(pk[1])
                        d["cigar_long_at"], ptr["cigar_long_at"] = pin(pk[2])
                    b = batch_from(c, lambda k, ptr=ptr: ptr.get(k), pk)
                else:
                    c8, esc16, esc_blk, long_len, long_at = c.pack_cigar8()
                    d["cigar8"], ptr["cigar8"] = pin(c8)
                    d["cigar_esc_blk"], ptr["cigar_esc_blk"] = pin(esc_blk)
                    if use_sq:
                        b = ffi.LpsReadBatch(n_reads=c.n_reads, cigar_len=len(c.cigar), sq=C.cast(ptr["sq"], ffi.u8p), sq_bytes=len(d["sq"]),
                                             **{k: C.cast(ptr[k], ptypes[k]) for k in names if k in ptr})
                    else:
                        b = ffi.LpsReadBatch(n_reads=c.n_reads, cigar_len=len(c.cigar), seq_bytes=len(c.seq4), qual_bytes=len(c.qual),
                                             **{k: C.cast(ptr[k], ptypes[k]) for k in names})
                    b.cigar8, b.cigar_esc_blk = C.cast(ptr["cigar8"], ffi.u8p), C.cast(ptr["cigar_esc_blk"], ffi.u32p)
                    b.n_cigar_esc, b.n_cigar_long = len(esc16), len(long_len)
                    if len(esc16):
                        d["cigar_esc16"], ptr["cigar_esc16"] = pin(esc16)
                        b.cigar_esc16 = C.cast(ptr["cigar_esc16"], ffi.u16p)
                    if len(long_len):
                        d["cigar_long_len"], ptr["cigar_long_len"] = pin(long_len)
                        d["cigar_long_at"], ptr["cigar_long_at"] = pin(long_at)
                        b.cigar_long_len, b.cigar_long_at = C.cast(ptr["cigar_long_len"], ffi.u32p), C.cast(ptr["cigar_long_at"], ffi.u64p)
                ptens.append(d)
                pin_batches.append(b)
            host_bytes = int(sum((t.nbytes if isinstance(t, np.ndarray) else t.numel()) for d in ptens for t in d.values()))

            def step_e2e(i, last):
                ctxs[i].submit(pin_batches[i])
                return ctxs[i].phase_contig(params, copy=last)
            run_threads(step_e2e, min(args.warmup, 2))
            s2 = [ctx.stats() for ctx in ctxs]
            e2e_dev_ms, e2e_wall_ms = timed(step_e2e, args.steps, 2)
            e2e_ms = max_over_ranks(e2e_dev_ms)
            s3 = [ctx.stats() for ctx in ctxs]
            d2h_step = int(sum(b_["d2h_bytes"] - a_["d2h_bytes"] for a_, b_ in zip(s2, s3)) / args.steps)
            h2d_step = int(sum(b_["h2d_bytes"] - a_["h2d_bytes"] for a_, b_ in zip(s2, s3)) / args.steps)
            ok_e2e, _ = digest_ok(results)
            if min_over_ranks(1.0 if ok_e2e else 0.0) < 1.0 and not debug_no_gate:
                raise SystemExit("bench.py: the end-to-end leg's result differs from the committed digest: no value is reported")
            out = (e2e_ms, e2e_wall_ms, h2d_step, d2h_step, host_bytes, use_sq)
            for i in range(n_ctg):
                ctxs[i].submit_device(dev_batches[i])       # nothing refers to the host buffers any more
            for a in pinned_in_place:
                cudart.cudaHostUnregister(a.ctypes.data)
            return out
        use_sq = not args.no_sq and not args.cigar32 and not args.cigar16
        e2e = run_e2e(use_sq)
        e2e_other = run_e2e(not use_sq) if args.e2e_ab and not args.cigar32 and not args.cigar16 else None

    if rank == 0:
        clocks = sampler.stop()
    calls_gpu = 0
    for i in range(n_ctg):
        ctxs[i].submit_device(dev_batches[i])
        calls_gpu += int(ctxs[i].call_alleles(params, want_host=False)["n_calls"])
    total_reads = sum_over_ranks(float(n_reads_gpu))
    total_calls = sum_over_ranks(float(calls_gpu))
    max_reads = max_over_ranks(float(n_reads_gpu))
    line = {
        "metric": METRIC, "value": total_reads / (ms_step * 1e-3), "unit": "reads/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if args.workload == "genome" else "weak", "vs_baseline": None, "dtype": "int32/f32", "data": "synthetic",
        "allele_calls_per_s": total_calls / (ms_step * 1e-3), "parity_digest_ok": bool(ok_resident),
        **({"resident_seq_qual": "interleaved rows (lps_read_batch.sq) in HBM: an experiment, not the default layout"} if args.resident_sq else {}), **({"invalid": "LPS_BENCH_DEBUG_NO_GATE=1: a debugging run, not a measurement"} if debug_no_gate else {}), "parity_digests_checked_rank0": n_checked,
        "config": {"workload": workload_text(args, world, n_ctg), "contigs_rank0": n_ctg, "host_threads_per_rank": T_,
                   "reads_total": int(total_reads), "reads_rank0": n_reads_gpu, "largest_rank_share_of_reads": max_reads / total_reads * world,
                   "variants_rank0": int(sum(c.n_var for c in contigs)), "allele_calls_total": int(total_calls),
                   "cigar_ops_per_read": float(np.mean([c.n_cigar.mean() for c in contigs])), "resident_bytes_rank0": resident_bytes,
                   "l2": "inputs (%.1f GB on rank 0) are far larger than the 126 MB L2; no flush needed" % (resident_bytes / 1e9),
                   "host_sync": "blocking" if blocking else ("yield" if args.sync == "yield" else "spin"), "host_cores": ncores, "numa_bind": numa,
                   "parallelism": "contigs sharded over %d GPU(s) by LPT on read counts, one lps_ctx (stream + scratch) per contig, %d host thread(s) per "
                                  "rank, no collective" % (world, T_),
                   "timing": "CUDA events on every context's stream (the streams the kernels run on), max over contexts and ranks",
                   "wall_ms_per_step_rank0": wall_ms, "synth_seconds": t_gen},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "rooflines": rooflines,
        "stage_ms": dict(k_call_alleles_alone=k1, k_fold_edges_alone=kf, **stage),
    }
    if e2e is not None:
        e2e_ms, e2e_wall_ms, h2d_step, d2h_step, host_bytes, use_sq = e2e
        line["e2e"] = {"value": total_reads / (e2e_ms * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                       "ms_per_step": e2e_ms, "wall_ms_per_step_rank0": e2e_wall_ms, "host_buffer_bytes": host_bytes, "parity_digest_ok": True,
                       "note": "pinned SEQ/QUAL stay on the host; the kernel gathers the sectors it needs over PCIe (zero-copy), "
                               "CIGAR and per-read records are copied; h2d bytes are the library's own count",
                       "seq_qual_wire_format": ("interleaved rows (lps_pack_sq / lps_read_batch.sq): ten qualities + ten 4-bit bases per 16-byte unit, "
                                                "one PCIe read request per allele call" if use_sq else
                                                "BAM's two arrays (seq4, qual): two PCIe read requests per allele call"),
                       "cigar_wire_format": ("uint32 (BAM), narrowed on the device" if args.cigar32 else
                                             "16-bit stream (lps_pack_cigar16), used as it arrives" if args.cigar16 else
                                             "8-bit stream (lps_pack_cigar8), expanded into the resident 16-bit stream by k_expand_cigar8")}
    if e2e is not None and e2e_other is not None:
        line["e2e"]["other_wire_format"] = {"seq_qual": "BAM's two arrays" if e2e[5] else "interleaved rows", "ms_per_step": e2e_other[0],
                                            "value": total_reads / (e2e_other[0] * 1e-3), "h2d_bytes_per_step": e2e_other[2]}
    if rank == 0 and world == 1 and not args.no_other_paths:
        # ---- the other dialects of the hot path (BASELINE configs C3 / C4), kernel-resident device time of one call each ----
        other = {}
        try:
            small = int(np.argmin([abs(c.n_reads - 99000) for c in contigs]))
            cs, ctx_s = contigs[small], ctxs[small]
            tp = ffi.default_tag_params()
            phased = cs.phased(res_resident[small]["ps"], res_resident[small]["hap_ref"] == 1)
            pv = phased.variants_struct()
            ctx_s.set_variants(pv, 0)
            ctx_s.submit_device(dev_batches[small])
            ms = []
            for _ in range(args.warmup + args.steps):
                r_tag = ctx_s.tag_reads(tp, want_calls=False)
                ms.append(ctx_s.stats()["ms_tag_reads"])
            ms = float(np.mean(ms[args.warmup:]))
            other["haplotag"] = {"config": "C3: germline haplotag of one %.0f Mb contig with the phase set of this run" % (len(cs.ref) / 1e6),
                                 "reads_per_s": cs.n_reads / (ms * 1e-3), "device_ms": ms, "alignments": cs.n_reads,
                                 "tagged": int((r_tag["hp"] != 0).sum())}
            un, ut = wl.c4_pair(synth_mod, 32.0)
            sp = ffi.LpsTagParams(mapping_quality=20, mapq_filter=0, tag_supplementary=1, have_reference=1, percentage_threshold=0.6)
            for name, c, cls in (("extract_normal", un, host.ExtractNorDataChrProcessor), ("extract_tumor", ut, host.ExtractTumDataChrProcessor),
                                 ("somatic_tag", ut, host.SomaticHaplotagChrProcessor)):
                proc = cls(ctx_s, c, sp)
                ms, wd = [], []
                for _ in range(args.warmup + args.steps):
                    r_s = proc.processSingleChrom(c)
                    st = ctx_s.stats()
                    ms.append(st["ms_tag_reads"]); wd.append(st["ms_kernel_window_diff"])
                ms = float(np.mean(ms[args.warmup:]))
                other[name] = {"reads_per_s": c.n_reads / (ms * 1e-3), "device_ms": ms, "alignments": c.n_reads, "tumor_positions": int(r_s["n_tum"])}
                if name == "extract_tumor":
                    wdm = float(np.mean(wd[args.warmup:]))
                    # 250 B of SEQ / reference / CIGAR + 8 B of histogram update per (tumor position, alignment) pair
                    gbs = 258.0 * r_s["n_window_items"] / (wdm * 1e-3) / 1e9 if wdm > 0 else None
                    other[name].update(window_items=r_s["n_window_items"], k_window_diff_ms=wdm, k_window_diff_gbs=gbs)
                    if gbs:
                        line["rooflines"].append({"bound": "hbm", "kernel": "k_window_diff", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                                  "algorithmic_bytes_per_launch": int(258 * r_s["n_window_items"]), "kernel_ms": wdm})
            other["somatic_config"] = "C4 shard: tumor 50x (purity 0.6) / normal 25x pair of one 32 Mb contig, ~%d somatic SNV+indel, union map of %d positions" % (
                int(un.var_is_somatic.sum()), un.n_var)
        except Exception as e:  # secondary numbers: never fail the headline line
            other["error"] = repr(e)
        try:
            other["bgzf_inflate"] = bench_bgzf(ctxs[0], host, ffi, torch, ncores, args)
        except Exception as e:
            other["bgzf_inflate"] = {"error": repr(e)}
        try:
            other["bgzf_deflate"] = bench_bgzf_deflate(ctxs[0], ncores, args, peak)
            line["rooflines"].append(other["bgzf_deflate"]["roofline"])
        except Exception as e:
            other["bgzf_deflate"] = {"error": repr(e)}
        line["other_paths"] = other
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = run_reference_cpu(args, min(ncores, 64), 1, 0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
    if rank == 0:
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    pool.shutdown()
    for ctx in ctxs:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
