"""Contig sharding across the GPUs of one box: contigs are independent units in the reference
(`#pragma omp parallel for` over chrName, src/phase/PhasingProcess.cpp:113; src/haplotag/HaplotagParsingBam.cpp:277), so each
rank owns whole contigs and there is NO collective on the data path.  The per-contig results are merged on the host exactly
like mergeAllChrPhasingResult (src/shared/Util.cpp:7-12): a plain map union keyed by "<chr>_<pos>".

No torch / CUDA imports here: the partition and the merge are host logic and are tested on CPU (gloo, world_size 2).
"""
import heapq

# GRCh38 primary contigs (Mb), the shape of a whole-genome job: the largest is 8 % of the total, so 8 ranks balance well
GRCH38_MB = {"chr1": 248.96, "chr2": 242.19, "chr3": 198.30, "chr4": 190.21, "chr5": 181.54, "chr6": 170.81, "chr7": 159.35,
             "chr8": 145.14, "chr9": 138.39, "chr10": 133.80, "chr11": 135.09, "chr12": 133.28, "chr13": 114.36, "chr14": 107.04,
             "chr15": 101.99, "chr16": 90.34, "chr17": 83.26, "chr18": 80.37, "chr19": 58.62, "chr20": 64.44, "chr21": 46.71,
             "chr22": 50.82, "chrX": 156.04, "chrY": 57.23}


def lpt_partition(weights, n_bins):
    """Greedy longest-processing-time partition.  weights: {name: weight} or a sequence; returns n_bins lists of keys,
    heaviest first inside a bin.  Deterministic: ties go to the lower bin index, equal weights keep their input order."""
    items = list(weights.items()) if isinstance(weights, dict) else list(enumerate(weights))
    order = sorted(range(len(items)), key=lambda i: (-items[i][1], i))
    bins = [[] for _ in range(n_bins)]
    heap = [(0.0, b) for b in range(n_bins)]
    heapq.heapify(heap)
    for i in order:
        load, b = heapq.heappop(heap)
        bins[b].append(items[i][0])
        heapq.heappush(heap, (load + float(items[i][1]), b))
    return bins


def bin_loads(weights, bins):
    w = weights if isinstance(weights, dict) else dict(enumerate(weights))
    return [sum(w[k] for k in b) for b in bins]


def contigs_of_rank(weights, world, rank):
    """The contigs rank `rank` of `world` processes; every rank computes the same partition from the same weights
    (read counts estimated from the BAM index in a real run), so nothing has to be communicated."""
    return lpt_partition(weights, world)[rank]


def export_phasing_result(chr_name, var_pos, ps, hap_ref):
    """PhasingResult of one contig (VairiantGraph::exportResult, src/phase/PhasingGraph.cpp:1049-1077):
    {"<chr>_<pos0>": ("a|b", PS)} for every variant that ended up in a phase set."""
    out = {}
    for i, p in enumerate(ps):
        if p:
            h = int(hap_ref[i])
            out[f"{chr_name}_{int(var_pos[i])}"] = (f"{h}|{1 - h}", int(p))
    return out


def merge_phasing_results(per_contig):
    """mergeAllChrPhasingResult (src/shared/Util.cpp:7-12): map union; std::map::insert keeps the first value of a key."""
    merged = {}
    for res in per_contig:
        for k, v in res.items():
            merged.setdefault(k, v)
    return merged


def merge_read_statistics(per_contig):
    """ReadStatistics are plain counters summed over contigs (mergeReadStatistics, src/haplotag/HaplotagProcess.cpp)."""
    total = {}
    for st in per_contig:
        for k, v in st.items():
            total[k] = total.get(k, 0) + int(v)
    return total
