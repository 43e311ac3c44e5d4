#!/usr/bin/env python
"""bench.py — throughput of the LongPhase-S `phase` read-to-variant hot path on B200.

One "step" = one pass of the whole hot path (allele calling -> host filters -> edge fold -> host sweep
-> read correction) over the synthetic contigs of this GPU (BASELINE.json config C2, contig-sharded; by default 4 x 64 Mb
per GPU, in flight concurrently on one lps_ctx + host thread each, like the reference's omp-parallel contig loop).
  value : reads/s with the read batch already resident in HBM (lps_batch_submit_device)
  e2e   : reads/s through the C ABI with pinned HOST buffers, H2D + D2H inside the timed region
  roofline : the dominant kernel (k_call_alleles) against the measured HBM copy peak
  cpu_baseline / --impl reference : the unmodified reference (oracle/_ref) on the host cores
Launch: `python bench.py --gpus 1` or torchrun for N > 1 (one rank per GPU, contigs are independent:
no collective on the data path; NCCL is only used for the barrier and the max-over-ranks of the time).
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

WORKLOAD = ("C2 shard: phase SNP+indel, {n} x {mb} Mb contigs per GPU ({tot} Mb), {depth:g}x ONT-like "
            "{kb:g} kb reads, 1 het variant / {sp:g} bp (10% indels), ONT error model")


def synth_kwargs(args, seed, scale=1.0):
    return dict(seed=seed, contig_len=int(args.contig_mb * 1_000_000 * scale), indel_frac=0.1, depth=args.depth, mean_len=args.mean_len,
                variant_rate=1.0 / args.variant_spacing)


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
            time.sleep(0.4)

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


def run_reference_cpu(args, n_threads, sample_mb):
    """The UNMODIFIED reference (oracle/_ref/libref_tap.so: get_snp, filterSNP, Clip, addEdge, edgeConnectResult,
    readCorrection, exportResult) on one bounded sample contig, every host thread running its own replica —
    the reference parallelises over contigs the same way (PhasingProcess.cpp:113)."""
    from oracle import pyoracle as po
    synth = importlib.import_module("longphase_s_b200.synth")
    ffi = importlib.import_module("longphase_s_b200._ffi")
    kw = synth_kwargs(args, 1000)
    kw["contig_len"] = int(sample_mb * 1_000_000)
    contig = synth.Contig(**kw)
    params = ffi.default_phase_params(True)
    if not po.tap_available():
        kind = "port"
        runner = lambda: po.OraclePhase(contig, params)  # noqa: E731
    else:
        kind = "reference"
        runner = lambda: po.ReferencePhase(contig, params)  # noqa: E731
    calls = []

    def work():
        r = runner()
        if kind == "reference":
            calls.append(int(r.stage_a["off"][-1]))
        else:
            calls.append(len(r.calls))

    def one_round():
        ths = [threading.Thread(target=work) for _ in range(n_threads)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0

    times = []
    for _ in range(args.warmup):
        one_round()
    for _ in range(args.steps):
        times.append(one_round())
    dt = float(np.mean(times))
    reads = contig.n_reads * n_threads
    return dict(value=reads / dt, unit="reads/s", cores=n_threads, kind=kind,
                sample=f"{n_threads} replicas of one {sample_mb} Mb contig ({contig.n_reads} reads each), "
                       f"{dt:.2f} s per round, {args.steps} rounds",
                allele_calls_per_s=(calls[0] * n_threads / dt) if calls else None, ms_per_step=dt * 1e3)


def bench_bgzf(ctx, host, ffi, torch, ncores, args, mb=1024):
    """inflation of BGZF blocks (htslib bgzf_read_block / inflate_block).  BAM-like synthetic bytes, deflated
    by zlib level 6 in 65280-byte members like htslib writes them; kernel-resident, end-to-end (host in, host out) and zlib on all
    host cores (what the reference's htslib build calls) on the same members."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, ROOT)
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
    os.environ["LPS_BGZF_SPECULATE"] = "0"           # A/B: the plain decoder (one table lookup per symbol)
    ms0 = []
    for _ in range(3):
        ctx.lib.lps_bgzf_inflate_device(ctx.h, d_data.data_ptr(), d_blocks.data_ptr(), len(blocks), d_out.data_ptr())
        ms0.append(ctx.stats()["ms_kernel_bgzf"])
    del os.environ["LPS_BGZF_SPECULATE"]
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
    # CPU: zlib inflate of the same members on every host core (a bounded sample: one pass over `unit`'s members per thread)
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
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    alg = len(data) + out_bytes
    return {"config": "%d MB of BAM-like bytes in %d BGZF members (zlib level 6, 65280 bytes each), ratio %.2f" % (out_bytes >> 20, len(blocks), out_bytes / len(data)),
            "kernel_ms": k_ms, "kernel_ms_plain_decoder": float(np.mean(ms0[1:])), "out_gb_per_s": out_bytes / (k_ms * 1e-3) / 1e9, "algorithmic_bytes": alg,
            "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (k_ms * 1e-3) / 1e9 / peak,
                         "note": "compressed bytes read + inflated bytes written; the kernel is bound by the serial Huffman chain of each block, not by HBM"},
            "e2e_ms": e2e_ms, "e2e_out_gb_per_s": out_bytes / (e2e_ms * 1e-3) / 1e9, "h2d_bytes": int(len(data)), "d2h_bytes": int(out_bytes),
            "cpu_zlib_out_gb_per_s": done / cpu_s / 1e9, "cpu_cores": ncores, "cpu_sample": "%d MB inflated per core, python zlib (libz inflate, GIL released)" % (len(unit) >> 20)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--contig-mb", type=float, default=64.0)
    ap.add_argument("--contigs-per-gpu", type=int, default=8)       # in flight per GPU; capped by this rank's share of the host cores
    ap.add_argument("--cpu-sample-mb", type=float, default=8.0)
    ap.add_argument("--depth", type=float, default=30.0)               # C5 stress: --depth 120 --mean-len 50000 --variant-spacing 300
    ap.add_argument("--mean-len", type=float, default=20000.0)
    ap.add_argument("--variant-spacing", type=float, default=1000.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-paths", action="store_true")
    ap.add_argument("--sync", default="auto", choices=["auto", "spin", "block", "yield"])   # yield: waits give the core away, contigs in flight are not capped by the cores
    ap.add_argument("--cigar32", action="store_true")    # end-to-end leg: send BAM's uint32 CIGAR ops instead of the compact 16-bit stream
    ap.add_argument("--unequal", type=int, default=0)   # 1: contig sizes +-25 % around --contig-mb (the largest one then bounds the step)
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

    if args.impl == "reference":
        if rank != 0:
            return 0
        os.environ["OMP_NUM_THREADS"] = str(ncores)
        res = run_reference_cpu(args, min(ncores, 64), args.cpu_sample_mb)
        line = {"impl": "reference", "metric": "phase_hot_path_reads_per_s", "value": res["value"], "unit": "reads/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f32", "data": "synthetic",
                "allele_calls_per_s": res["allele_calls_per_s"],
                "config": {"workload": WORKLOAD.format(n=args.contigs_per_gpu, mb=args.contig_mb, tot=args.contigs_per_gpu * args.contig_mb, depth=args.depth, kb=args.mean_len / 1e3, sp=args.variant_spacing),
                           "sample": res["sample"]},
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    ffi0 = importlib.import_module("longphase_s_b200._ffi")
    # "auto": spin while every contig in flight has a host core of its own; with fewer cores (8 ranks on a 32-core box) keep all the
    # contigs in flight anyway and let waiting threads yield their core.  Measured on 8 GPUs / 32 cores: 4 contigs per rank
    # spinning 359.6 M reads/s, 8 per rank yielding 433.7 M; blocking waits double the step (21.1 vs 9.8 ms).  On 1 GPU / 16 cores
    # spinning wins (80.3 M against 69.6 M with 8 yielding threads).
    if args.sync == "auto":
        args.sync = "yield" if ncores // world < args.contigs_per_gpu else "spin"
    blocking = args.sync == "block"
    if blocking or args.sync == "yield":
        ffi0.load_library().lps_set_blocking_sync(local_rank, 1 if blocking else 2)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    synth_mod = importlib.import_module("longphase_s_b200.synth")
    host = importlib.import_module("longphase_s_b200.host")
    ffi = importlib.import_module("longphase_s_b200._ffi")

    from concurrent.futures import ThreadPoolExecutor
    # contigs in flight per GPU: each needs a host thread, so never more than this rank's share of the host cores
    C_ = max(1, args.contigs_per_gpu if args.sync == "yield" else min(args.contigs_per_gpu, ncores // world))
    t_gen = time.time()
    # unequal contigs (a genome's are): same total, +-25 % around --contig-mb
    shape = [1.25, 0.9, 1.1, 0.75] if args.unequal and C_ % 4 == 0 else [1.0]
    contigs = [synth_mod.Contig(**synth_kwargs(args, 100 + 16 * rank + i, shape[i % len(shape)])) for i in range(C_)]
    t_gen = time.time() - t_gen
    params = ffi.default_phase_params(True)
    # one context (own stream, own scratch) per contig in flight, driven by its own host thread: the reference runs its contig
    # loop the same way (`#pragma omp parallel for`, PhasingProcess.cpp:113), and the host parts of one contig (overlap filter,
    # edgeConnectResult chain) overlap the kernels of the others
    ctxs = [host.Context(local_rank) for _ in range(C_)]
    keep = []
    for ctx, contig in zip(ctxs, contigs):
        ctx.set_reference(contig.ref)
        vs = contig.variants_struct()
        keep.append(vs)
        ctx.set_variants(vs, True)

    # ---- device-resident copy of every batch (torch owns the memory) ----
    def dev(a):
        view = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64}.get(a.dtype)
        return torch.from_numpy(a.view(view) if view else a).cuda()

    names = ["ref_start", "l_qseq", "n_cigar", "cigar_off", "seq_off", "qual_off", "flag", "mapq", "name_rank", "cigar", "seq4", "qual"]
    ptypes = dict(ref_start=ffi.i32p, l_qseq=ffi.i32p, n_cigar=ffi.u32p, cigar_off=ffi.u64p, seq_off=ffi.u64p, qual_off=ffi.u64p,
                  flag=ffi.u16p, mapq=ffi.u8p, name_rank=ffi.i32p, cigar=ffi.u32p, seq4=ffi.u8p, qual=ffi.u8p)

    def batch_from(contig, ptr_of):
        return ffi.LpsReadBatch(n_reads=contig.n_reads, cigar_len=len(contig.cigar), seq_bytes=len(contig.seq4),
                                qual_bytes=len(contig.qual), **{k: C.cast(ptr_of(k), ptypes[k]) for k in names})

    dtens = [{k: dev(getattr(c, k)) for k in names} for c in contigs]
    dev_batches = [batch_from(c, lambda k, d=d: d[k].data_ptr()) for c, d in zip(contigs, dtens)]
    # ---- pinned host copies for the end-to-end leg ----
    # The host loop packs the CIGAR ops of every record into the compact 16-bit wire format while it appends the record to the
    # batch (lps_pack_cigar16; include/lps.h): the pinned batch of the end-to-end leg holds that stream, not the uint32 ops.
    e2e_names = list(names)
    packed = None
    if not args.cigar32:
        packed = [c.pack_cigar16() for c in contigs]
        e2e_names.remove("cigar")
    ptens = [{k: torch.from_numpy(getattr(c, k).view(np.uint8).reshape(-1)).pin_memory() for k in e2e_names} for c in contigs]
    pin_batches = []
    for i, (c, d) in enumerate(zip(contigs, ptens)):
        if packed is None:
            pin_batches.append(batch_from(c, lambda k, d=d: d[k].data_ptr()))
            continue
        c16, long_len, long_at = packed[i]
        d["cigar16"] = torch.from_numpy(c16.view(np.uint8)).pin_memory()
        d["cigar_long_len"] = torch.from_numpy(np.ascontiguousarray(long_len).view(np.uint8)).pin_memory()
        d["cigar_long_at"] = torch.from_numpy(np.ascontiguousarray(long_at).view(np.uint8)).pin_memory()
        b = batch_from(c, lambda k, d=d: d[k].data_ptr() if k != "cigar" else None)
        b.cigar16 = C.cast(d["cigar16"].data_ptr(), ffi.u16p)
        b.n_cigar_long = len(long_len)
        if len(long_len):
            b.cigar_long_len = C.cast(d["cigar_long_len"].data_ptr(), ffi.u32p)
            b.cigar_long_at = C.cast(d["cigar_long_at"].data_ptr(), ffi.u64p)
        pin_batches.append(b)
    host_bytes = int(sum(t.numel() for d in ptens for t in d.values()))
    input_bytes = host_bytes
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

    pool = ThreadPoolExecutor(max_workers=C_)

    def run_all(fn):
        return [f.result() for f in [pool.submit(fn, i) for i in range(C_)]]

    # `last`: only the final pass of a timed region turns the result into numpy arrays (Python work under the GIL, serialised over
    # the host threads of a rank); every pass brings the result to host memory inside lps_phase_solve
    def step_resident(i, last=True):
        return ctxs[i].phase_contig(params, copy=last)      # the batch was registered once with lps_batch_submit_device (no copy)

    def step_e2e(i, last=True):
        ctxs[i].submit(pin_batches[i])
        return ctxs[i].phase_contig(params, copy=last)

    def timed(fn, steps, slot):
        """`steps` passes over all contigs of this GPU; device time between the events of every context's stream, max over them."""
        barrier()
        for ctx in ctxs:
            ctx.event_record(slot)
        t0 = time.perf_counter()
        # every host thread runs its `steps` passes back to back, with no barrier between steps: like the reference's contig loop,
        # the threads drift out of lockstep, so the host phases of one contig overlap the kernels of another
        out = run_all(lambda i: [fn(i, k == steps - 1) for k in range(steps)][-1])
        for ctx in ctxs:
            ctx.event_record(slot + 1)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        return max(ctx.event_elapsed_ms(slot, slot + 1) for ctx in ctxs) / steps, wall / steps, out

    # ---- kernel-resident leg ----
    for i in range(C_):
        ctxs[i].submit_device(dev_batches[i])
    for _ in range(args.warmup):
        run_all(step_resident)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()    # the JSON line is rank 0's: only its GPU is sampled (every nvidia-smi call costs host CPU)
    s0 = [ctx.stats() for ctx in ctxs]
    dev_ms, wall_ms, res = timed(step_resident, args.steps, 0)
    s1 = [ctx.stats() for ctx in ctxs]
    clocks = sampler.stop()
    ms_step = max_over_ranks(dev_ms)
    launches = int(sum(b_["kernel_launches"] - a_["kernel_launches"] for a_, b_ in zip(s0, s1)))
    stage_keys = ("ms_call_alleles", "ms_build_edges", "ms_read_correction", "ms_wall_call_alleles", "ms_wall_build_edges", "ms_wall_solve",
                  "ms_host_filters", "ms_host_sweep", "ms_kernel_fold_edges", "ms_sweep")
    stage = {k[3:]: float(np.mean([st[k] for st in s1])) for k in stage_keys}
    stage["sweep_fallbacks"] = int(sum(b_["sweep_fallbacks"] - a_["sweep_fallbacks"] for a_, b_ in zip(s0, s1)))
    stage["slow_path_contigs"] = int(sum(b_["slow_path_contigs"] - a_["slow_path_contigs"] for a_, b_ in zip(s0, s1)))

    # ---- the dominant kernel timed alone (one context, nothing else on the GPU): roofline ----
    ctx0, contig0 = ctxs[0], contigs[0]
    calls = ctx0.call_alleles(params, want_host=True)
    n_calls0 = calls["n_calls"]
    b1 = algorithmic_bytes_k1(contig0, calls["read_status"], n_calls0)
    k1_ms = []
    for _ in range(max(args.steps, 5)):
        ctx0.call_alleles(params, want_host=False)
        k1_ms.append(ctx0.stats()["ms_kernel_call_alleles"])
    k1 = float(np.mean(k1_ms))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy; the kernel is timed alone)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = b1 / (k1 * 1e-3) / 1e9 if k1 > 0 else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_k1_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("contig_mb") == args.contig_mb and tj.get("reads") == contig0.n_reads:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]

    # ---- end-to-end leg: pinned host buffers through the C ABI ----
    for _ in range(min(args.warmup, 2)):
        run_all(step_e2e)
    s2 = [ctx.stats() for ctx in ctxs]
    e2e_dev_ms, e2e_wall_ms, res_e2e = timed(step_e2e, args.steps, 2)
    e2e_ms = max_over_ranks(e2e_dev_ms)
    s3 = [ctx.stats() for ctx in ctxs]
    d2h_step = int(sum(b_["d2h_bytes"] - a_["d2h_bytes"] for a_, b_ in zip(s2, s3)) / args.steps)
    h2d_step = int(sum(b_["h2d_bytes"] - a_["h2d_bytes"] for a_, b_ in zip(s2, s3)) / args.steps)
    for ra, rb in zip(res, res_e2e):
        for k in ("ps", "hap_ref", "read_hp"):
            assert np.array_equal(ra[k], rb[k]), "resident and end-to-end legs disagree"
    # allele calls of every contig (one extra pass, outside the timed regions)
    calls_gpu = 0
    for i in range(C_):
        ctxs[i].submit_device(dev_batches[i])
        calls_gpu += int(ctxs[i].call_alleles(params, want_host=False)["n_calls"])

    total_reads = sum_over_ranks(float(n_reads_gpu))
    total_calls = sum_over_ranks(float(calls_gpu))
    line = {
        "metric": "phase_hot_path_reads_per_s", "value": total_reads / (ms_step * 1e-3), "unit": "reads/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32/f32", "data": "synthetic",
        "allele_calls_per_s": total_calls / (ms_step * 1e-3),
        "config": {"workload": WORKLOAD.format(n=C_, mb=args.contig_mb, tot=C_ * args.contig_mb, depth=args.depth, kb=args.mean_len / 1e3, sp=args.variant_spacing), "contigs_per_gpu": C_,
                   "reads_per_gpu": n_reads_gpu, "variants_per_gpu": int(sum(c.n_var for c in contigs)), "allele_calls_per_gpu": calls_gpu,
                   "cigar_ops_per_read": float(np.mean([c.n_cigar.mean() for c in contigs])), "input_bytes_per_gpu": input_bytes,
                   "l2": "inputs (%.1f GB per GPU) are far larger than the 126 MB L2; no flush needed" % (input_bytes / 1e9),
                   "host_sync": "blocking" if blocking else ("yield" if args.sync == "yield" else "spin"), "host_cores": ncores,
                   "parallelism": f"contig-sharded x{world} GPUs, {C_} contigs in flight per GPU (one lps_ctx + host thread each), no collective",
                   "timing": "CUDA events on every context's stream (the streams the kernels run on), max over contexts and ranks",
                   "wall_ms_per_step_rank0": wall_ms, "synth_seconds": t_gen},
        "e2e": {"value": total_reads / (e2e_ms * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                "ms_per_step": e2e_ms, "host_buffer_bytes": host_bytes,
                "note": "pinned SEQ/QUAL stay on the host; the kernel gathers the sectors it needs over PCIe (zero-copy), "
                        "CIGAR and per-read records are copied; h2d bytes are the library's own count",
                "cigar_wire_format": "uint32 (BAM)" if args.cigar32 else "16-bit compact stream (lps_pack_cigar16), widened on the device"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_call_alleles", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": b1, "kernel_ms": k1,
                     "peak_source": peak_src, "launch": f"one {args.contig_mb} Mb contig ({contig0.n_reads} reads), timed alone"},
        "stage_ms": dict(k_call_alleles_alone=k1, **stage),
    }
    if rank == 0 and world == 1 and not args.no_other_paths:
        # ---- the other dialects of the hot path (BASELINE configs C3 / C4), kernel-resident device time of one call each ----
        other = {}
        try:
            tp = ffi.default_tag_params()
            phased = contig0.phased(res[0]["ps"], res[0]["hap_ref"] == 1)
            pv = phased.variants_struct()
            ctx0.set_variants(pv, 0)
            ctx0.submit_device(dev_batches[0])
            ms = []
            for _ in range(args.warmup + args.steps):
                r_tag = ctx0.tag_reads(tp, want_calls=False)
                ms.append(ctx0.stats()["ms_tag_reads"])
            ms = float(np.mean(ms[args.warmup:]))
            other["haplotag"] = {"config": "C3: germline haplotag of one %.0f Mb contig with the phase set of this run" % args.contig_mb,
                                 "reads_per_s": contig0.n_reads / (ms * 1e-3), "device_ms": ms, "alignments": contig0.n_reads,
                                 "tagged": int((r_tag["hp"] != 0).sum())}
            kw = synth_kwargs(args, 900)
            kw.update(contig_len=int(min(args.contig_mb, 32.0) * 1_000_000), somatic_rate=3000.0 / 64e6, indel_frac=0.1)
            kn, kt = dict(kw), dict(kw)
            kn.update(depth=25.0, purity=0.0, read_seed=901)
            kt.update(depth=50.0, purity=0.6, read_seed=902)
            cn, ct = synth_mod.Contig(**kn), synth_mod.Contig(**kt)
            un = cn.somatic_union(seed=9)
            ut = un.with_reads_of(ct)
            sp = ffi.LpsTagParams(mapping_quality=20, mapq_filter=0, tag_supplementary=1, have_reference=1, percentage_threshold=0.6)
            for name, c, cls in (("extract_normal", un, host.ExtractNorDataChrProcessor), ("extract_tumor", ut, host.ExtractTumDataChrProcessor),
                                 ("somatic_tag", ut, host.SomaticHaplotagChrProcessor)):
                proc = cls(ctx0, c, sp)
                ms, wd = [], []
                for _ in range(args.warmup + args.steps):
                    r_s = proc.processSingleChrom(c)
                    st = ctx0.stats()
                    ms.append(st["ms_tag_reads"]); wd.append(st["ms_kernel_window_diff"])
                ms = float(np.mean(ms[args.warmup:]))
                other[name] = {"reads_per_s": c.n_reads / (ms * 1e-3), "device_ms": ms, "alignments": c.n_reads, "tumor_positions": int(r_s["n_tum"])}
                if name == "extract_tumor":
                    wdm = float(np.mean(wd[args.warmup:]))
                    # 250 B of SEQ / reference / CIGAR + 8 B of histogram update per (tumor position, alignment) pair
                    other[name].update(window_items=r_s["n_window_items"], k_window_diff_ms=wdm,
                                       k_window_diff_gbs=258.0 * r_s["n_window_items"] / (wdm * 1e-3) / 1e9 if wdm > 0 else None)
            other["somatic_config"] = "C4 shard: tumor 50x (purity 0.6) / normal 25x pair of one %.0f Mb contig, ~%d somatic SNV+indel, union map of %d positions" % (
                kw["contig_len"] / 1e6, int(un.var_is_somatic.sum()), un.n_var)
        except Exception as e:  # secondary numbers: never fail the headline line
            other["error"] = repr(e)
        try:
            other["bgzf_inflate"] = bench_bgzf(ctx0, host, ffi, torch, ncores, args)
        except Exception as e:
            other["bgzf_inflate"] = {"error": repr(e)}
        line["other_paths"] = other
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = run_reference_cpu(args, min(ncores, 64), args.cpu_sample_mb)
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
