#!/usr/bin/env python
"""bench.py -- aligned bases counted per second on the pileup counting hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

One "step" = one pass of the hot path over one batch: zero the accumulators, K1 (CIGAR
walk + count, with its sparse corrections), K2+K3 (--summarise reductions) and the
read-back of the per-sample summary scalars.

Default workload (`cfg2x12`): BASELINE.json configs[1] -- the SARS-CoV-2 29,903 bp
reference, ~124k synthetic 400 bp amplicon reads per sample, `--summarise` -- batched as
12 independent samples per GPU (configs[3]: 96 samples over 8 GPUs = 12 per GPU), so the
per-step input (~180 MB) exceeds the 126 MB L2 and N GPUs weak-scale to 12*N samples with
no data-path collective.  Two different resident batches alternate between steps.

The JSON line carries `value` (inputs resident in HBM, CUDA-event timed), `e2e` (the same
step through the C ABI from pinned host buffers, H2D and D2H inside the timed region),
`roofline` for the counting kernel against MEASURED_PEAKS.json and `cpu_baseline`
(the compiled reference operator + the oracle's port of get_stats on one host core).
`--impl reference` times the reference's CPU path with every host core (one process per
sample); it is the only arm that executes oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE line, the JSON record: everything else a process writes to file descriptor 1 while the
# bench runs (NCCL prints "NCCL version ..." there at the first communicator, make, child processes) goes to stderr.
_JSON_FD = None


def claim_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


METRIC = "aligned_bases_per_sec"
UNIT = "aligned bases/s"
SAMPLES_PER_GPU = 12
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback
# DRAM bytes per K1 launch from the committed ncu --set full captures of the default shapes:
# dram__bytes_read.sum + dram__bytes_write.sum of k1_count_fast, one launch, ncu --set full (profiles/r4_a_k1_summary.md)
NCU_TRAFFIC_BYTES = {"cfg2x12": 197_192_960, "cfg3": 118_753_536, "cfg5": 2_832_785_000}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["cfg2x12", "cfg3", "cfg5"],
                    help="one workload only; default: cfg2x12 as the line, with cfg3 / cfg5 sub-records at one GPU and the "
                         "region-sharded cfg5 sub-record at several")
    ap.add_argument("--samples-per-gpu", type=int, default=SAMPLES_PER_GPU)
    ap.add_argument("--reads-per-sample", type=int, default=124_000)
    ap.add_argument("--cfg5-reads", type=int, default=12_888_833)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gen-samples", default=None, help="internal: synthesise these sample seeds into /tmp and exit")
    ap.add_argument("--variant", type=int, default=0, help="0 k1_count_fast + general walker behind it, 1 per-base atomics K1, 2 the general walker alone")
    return ap.parse_args()


# ----------------------------------------------------------------------------- workloads
def _sample_path(seed, n_reads):
    return f"/tmp/bc_bench_sample_{seed}_{n_reads}.npz"


def _generate_sample(job):
    """Worker: synthesise one config-1/2 sample into the /tmp cache (numpy only, no CUDA)."""
    seed, n_reads = job
    from basecount_b200 import synth
    from basecount_b200.records import select_reads
    b = select_reads(synth.amplicon_sample(seed=seed, n_reads=n_reads), 0, 0)
    path = _sample_path(seed, n_reads)
    tmp = f"{path}.{os.getpid()}.tmp.npz"
    np.savez(tmp, starts=b.starts, cigar=b.cigar, cigar_off=b.cigar_off, seq=b.seq, qual=b.qual, seq_off=b.seq_off)
    os.replace(tmp, path)
    return seed


def make_samples(seeds, n_reads):
    """Synthetic config-1/2 samples, filtered and trimmed as get_basecounts would.  Cached in /tmp so
    repeated invocations on one box (plain run, then ncu; N = 1, 2, 4, 8 back to back) do not regenerate;
    missing ones are synthesised by child processes of this script sharing the host cores between the ranks."""
    from basecount_b200.records import ReadBatch
    seeds = list(seeds)
    missing = [(s, n_reads) for s in seeds if not os.path.exists(_sample_path(s, n_reads))]
    if missing:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        workers = max(1, min(len(missing), (os.cpu_count() or 1) // max(world, 1), 12))
        if workers > 1:                                  # plain child processes of this script (numpy only, no CUDA)
            procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--gen-samples",
                                       ",".join(str(j[0]) for j in missing[w::workers]), "--reads-per-sample", str(n_reads)])
                     for w in range(workers)]
            for pr in procs:
                if pr.wait() != 0:
                    raise RuntimeError("sample generation failed")
        else:
            for job in missing:
                _generate_sample(job)
    out = []
    for s in seeds:
        z = np.load(_sample_path(s, n_reads))
        out.append(ReadBatch(z["starts"], z["cigar"], z["cigar_off"], z["seq"], z["qual"], z["seq_off"]))
    return out


def build_workload(args, rank, workload):
    """Returns (list of ReadBatch lists [one per alternating batch], ref_lens, label)."""
    from basecount_b200 import synth
    from basecount_b200.records import select_reads
    if workload == "cfg2x12":
        s = args.samples_per_gpu
        base = 100 + rank * 2 * s
        sets = [make_samples(range(base, base + s), args.reads_per_sample),
                make_samples(range(base + s, base + 2 * s), args.reads_per_sample)]
        return sets, [synth.SARS2_LEN] * s, f"cfg2_summarise_x{s}_samples_per_gpu"
    if workload == "cfg3":
        sets = [[select_reads(synth.deep_short_read_sample(seed=3 + 10 * rank + i), 0, 0)] for i in range(2)]
        return sets, [synth.SARS2_LEN], "cfg3_2M_reads_150bp"
    n = args.cfg5_reads
    L = int(synth.CHR20_LEN * (n / 12_888_833))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # region sharding: ONE reference, rank r owns [bounds[r], bounds[r+1]) and generates
        # only the reads that start there (they may run past the right edge -> halo columns)
        from basecount_b200 import dist as bdist
        bounds = bdist.region_bounds(L, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        rec = synth.uniform_short_read_sample(seed=5 + rank, ref_len=L, n_reads=n // world, start_lo=lo,
                                              start_hi=min(hi, L - 150))
        local = bdist.select_region(select_reads(rec, 0, 0), lo, hi)
        return [[local]], [L], f"cfg5_uniform_150bp_L{L}_region_sharded"
    sets = [[select_reads(synth.uniform_short_read_sample(seed=5 + rank, ref_len=L, n_reads=n), 0, 0)]]
    return sets, [L], f"cfg5_uniform_150bp_L{L}"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def wait_first(self, timeout=3.0):
        """Block until nvidia-smi has delivered its first sample: its start-up (process creation, NVML initialisation,
        the first query) takes driver locks and tens of milliseconds, which must not fall into a timed region that
        is two milliseconds long."""
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.01)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU legs (execute oracle/)
def _cpu_sample_worker(payload):
    """One sample through the reference's CPU path: compiled count.cpp bcount (oracle/_ref) when
    present, else the C port; then the oracle's port of get_stats + the summarise block."""
    from oracle import bcount as obc
    from oracle import stats as ost
    lists, ref_len, use_ref = payload
    t0 = time.perf_counter()
    if use_ref:
        counts = obc.load_ref_bcount()(ref_len, 0, *lists)
    else:
        from basecount_b200.records import ReadBatch
        counts = obc.bcount_flat(ref_len, 0, ReadBatch.from_lists(*lists)).tolist()
    t1 = time.perf_counter()
    cov, ent, _ = ost.per_position_vectors(counts)
    ost.summary(cov, ent, ref_len)
    return t1 - t0, time.perf_counter() - t1


def cpu_baseline(samples, ref_len, n_samples=3):
    """Single core (the reference is single-threaded): bcount + get_stats + summary on a bounded sample."""
    from oracle import bcount as obc
    use_ref = obc.load_ref_bcount() is not None
    chosen = samples[:n_samples]
    bases = sum(b.aligned_bases() for b in chosen)
    payloads = [(b.to_lists(), ref_len, use_ref) for b in chosen]      # list building is not timed (pysam's job)
    t_count = t_stats = 0.0
    for p in payloads:
        a, b = _cpu_sample_worker(p)
        t_count += a
        t_stats += b
    return {"value": bases / (t_count + t_stats), "unit": UNIT, "cores": 1,
            "kind": "reference+port" if use_ref else "port",
            "sample": f"{len(chosen)} of the step's samples ({bases} aligned bases): "
                      f"{'compiled count.cpp bcount (oracle/_ref)' if use_ref else 'C port of bcount'} {t_count:.2f}s "
                      f"+ python port of get_stats/summarise {t_stats:.2f}s; Python-list inputs prebuilt"}


def _reference_worker(conn, seed, n_reads, ref_len, use_ref):
    """One host process owning one sample: builds the reference's Python-list inputs once (that is
    pysam's job in the reference and is not timed), then runs the CPU path every time it is told to."""
    sample = make_samples([seed], n_reads)[0]
    payload = (sample.to_lists(), ref_len, use_ref)
    conn.send(("ready", sample.aligned_bases()))
    while True:
        msg = conn.recv()
        if msg != "go":
            break
        conn.send(_cpu_sample_worker(payload))
    conn.close()


def run_reference(args):
    """--impl reference: the reference's CPU path (compiled count.cpp bcount + get_stats/summarise
    port) on every host core: one persistent process per sample, all samples of a step in parallel."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from basecount_b200 import synth
    from oracle import bcount as obc
    use_ref = obc.load_ref_bcount() is not None
    cores = os.cpu_count() or 1
    # our arm's step at N GPUs is 12 x N samples (weak scaling); the CPU arm keeps every host core busy with one
    # sample each, up to that many -- a bounded sample of the step: the RATE it measures does not depend on how
    # many samples a step holds once all cores are busy
    per_step = max(1, min(args.samples_per_gpu * max(args.gpus, 1), cores))
    make_samples(range(100, 100 + per_step), args.reads_per_sample)          # fill the /tmp cache once, serially
    ctx = mp.get_context("fork")
    workers = []
    for i in range(per_step):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=_reference_worker, args=(b, 100 + i, args.reads_per_sample, synth.SARS2_LEN, use_ref),
                         daemon=True)
        pr.start()
        workers.append((pr, a))
    bases = sum(a.recv()[1] for _, a in workers)
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))

    def one_step():
        for _, a in workers:
            a.send("go")
        return [a.recv() for _, a in workers]

    for _ in range(warm):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    for pr, a in workers:
        a.send("stop")
        pr.join(timeout=10)
    value = bases * steps / dt
    kind = "reference+port" if use_ref else "port"
    sample = (f"bounded sample of the {args.samples_per_gpu * max(args.gpus, 1)}-sample step: {per_step} samples x "
              f"{args.reads_per_sample} reads per timed step, one persistent process per sample on {cores} host cores "
              f"(all busy); compiled count.cpp bcount (oracle/_ref) + python port of get_stats / summarise "
              f"(the reference's main.py is not on the GPU box); Python-list inputs prebuilt (pysam's job); "
              f"steps capped at {steps}, warmup {warm}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"cfg2_summarise_x{args.samples_per_gpu}_samples_per_gpu", "samples_per_timed_step": per_step,
                       "reads_per_sample": args.reads_per_sample, "ref_len": synth.SARS2_LEN},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": per_step, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------- BAM file -> summary
def bam_end_to_end(eng, with_reference):
    """wall clock from a BAM path to the --summarise numbers of one config-1 sample (124,000
    reads x 400 bp) through the product's own API: native block-parallel BGZF/BAM decode, filter + soft-clip
    trimming, 2-bit packing, H2D, K1, K2/K3, D2H.  Beside it (optional) the reference's path on one host
    core: alignment decode by the in-repo numpy reader (pysam is not installable in this image), Python
    lists as pysam would hand them over, the compiled count.cpp bcount and the get_stats/summarise port."""
    from basecount_b200 import bamio, synth
    from basecount_b200.main import count_alignments
    from basecount_b200.records import select_reads
    path = "/tmp/bc_bench_cfg1_seed100.bam"
    if not os.path.exists(path):
        bamio.write_bam(path + ".tmp", synth.amplicon_sample(seed=100))
        os.replace(path + ".tmp", path)
    count_alignments(path, engine=eng)                     # warm-up: page cache, allocations
    eng.summary(False)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        pile = count_alignments(path, engine=eng)
        nz, cs, es = eng.summary(False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    b = select_reads(bamio.read_bam(path), 0, 0)
    bases = b.aligned_bases()
    out = {"seconds": best, "aligned_bases": bases, "reads": int(pile.num_reads[0]), "value": bases / best, "unit": UNIT,
           "decoder": "native (csrc/bam_decode.h: block-parallel inflate, one pass from records to the packed batch), best of 3",
           "bam_mb": os.path.getsize(path) / 1e6}
    if with_reference:
        from oracle import bcount as obc
        from oracle import stats as ost
        use_ref = obc.load_ref_bcount() is not None
        t0 = time.perf_counter()
        rb = select_reads(bamio.read_bam(path), 0, 0)
        lists = rb.to_lists()
        t1 = time.perf_counter()
        counts = obc.load_ref_bcount()(synth.SARS2_LEN, 0, *lists) if use_ref else obc.bcount_flat(synth.SARS2_LEN, 0, rb).tolist()
        cov, ent, _ = ost.per_position_vectors(counts)
        ost.summary(cov, ent, synth.SARS2_LEN)
        t2 = time.perf_counter()
        out["reference"] = {"seconds": t2 - t0, "decode_seconds": t1 - t0, "count_and_stats_seconds": t2 - t1, "cores": 1,
                            "value": bases / (t2 - t0), "unit": UNIT,
                            "decoder": "in-repo numpy BAM reader standing in for pysam (absent in this image)",
                            "kind": "reference" if use_ref else "port"}
    return out


def bam_batch_end_to_end(local, n_files, barrier, max_over_ranks, sum_over_ranks):
    """BASELINE configs[3] through the BAM path: every rank takes `n_files` config-1 BAMs (96 over 8 GPUs = 12 per
    rank) from file path to --summarise numbers through the product API -- native decode, one-pass pack, H2D, K1,
    K2 -- opening the next file on a second host thread while the current one is packed and counted.  All ranks run
    at once and share the box's host cores, which is what bounds this path (the GPU part is ~2 ms per file)."""
    from concurrent.futures import ThreadPoolExecutor
    from basecount_b200 import bamio, synth
    from basecount_b200.engine import Engine
    from basecount_b200.main import count_alignments
    path = "/tmp/bc_bench_cfg1_seed100.bam"
    if not os.path.exists(path):
        tmp = f"{path}.{os.getpid()}.tmp"
        bamio.write_bam(tmp, synth.amplicon_sample(seed=100))
        os.replace(tmp, path)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = max(1, (os.cpu_count() or 1) // max(world, 1))
    eng = Engine(local)
    count_alignments(path, engine=eng)                     # warm-up: page cache, allocations
    eng.summary(False)
    bases_per_file = int(eng.counts(0).sum())              # every aligned base is in exactly one cell
    barrier()
    t0 = time.perf_counter()
    reads = 0
    with ThreadPoolExecutor(max_workers=1) as pool:
        fut = pool.submit(bamio.NativeBam, path, threads)
        for k in range(n_files):
            nb = fut.result()
            if k + 1 < n_files:
                fut = pool.submit(bamio.NativeBam, path, threads)
            pile = count_alignments(nb, engine=eng)
            eng.summary(False)
            reads += int(pile.num_reads[0])
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    eng.close()
    total = sum_over_ranks(float(bases_per_file * n_files))
    return {"seconds": dt, "files_per_rank": n_files, "files_total": int(n_files * world), "reads_per_rank": reads,
            "value": total / dt, "unit": UNIT,
            "host_threads_per_rank": threads, "host_cores": os.cpu_count(),
            "decoder": "native (csrc/bam_decode.h); the same 39.7 MB config-1 BAM read n times per rank (page-cache hot); "
                       "next file opened on a second thread while the current one is packed and counted"}


def tsv_stream_end_to_end(with_reference, ref_len=4_000_000):
    """`basecount BAM_FILE` (BASELINE configs[0]'s mode) at a whole-genome shape: a 4 Mb reference at 30x (800 k reads x
    150 bp -- configs[4]'s shape at 1/16 of its length) from the BAM path to the per-position TSV on disk, streamed
    window by window (BaseCount.write_tsv: bc_rows_window + the native emitter, main.py:454-466).  rows/s from file
    path to the last byte written; the first 50,000 rows are compared byte for byte with the oracle's text, and the
    reference's own Python path (get_stats + str(round())) is timed on those rows beside it."""
    from basecount_b200 import BaseCount, bamio, synth
    path = f"/tmp/bc_bench_region_{ref_len}.bam"
    if not os.path.exists(path):
        rec = synth.uniform_short_read_sample(seed=5, ref_len=ref_len, n_reads=ref_len * 30 // 150, read_len=150, ref_name="chr20s")
        bamio.write_bam(path, rec)
    out_path = "/tmp/bc_bench_rows.tsv"
    best, rows, size = None, 0, 0
    for _ in range(2):
        t0 = time.perf_counter()
        with BaseCount(path) as bc, open(out_path, "w") as f:
            rows = bc.write_tsv(f)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        size = os.path.getsize(out_path)
    rec = {"seconds": best, "rows": rows, "value": rows / best, "unit": "rows/s", "tsv_mb": size / 1e6, "ref_len": ref_len,
           "bam_mb": os.path.getsize(path) / 1e6,
           "path": "native BAM decode -> K1 -> K2 rows per 524,288-position window -> native emitter (tsv_format.h) -> file"}
    if with_reference:
        from oracle import bcount as obc
        from oracle import stats as ost
        from basecount_b200.records import select_reads
        n_check = 50_000
        b = select_reads(synth.uniform_short_read_sample(seed=5, ref_len=ref_len, n_reads=ref_len * 30 // 150, read_len=150,
                                                         ref_name="chr20s"), 0, 0)            # what the BAM was written from
        part = synth.take_batch(b, np.flatnonzero(b.starts < n_check))
        counts = obc.bcount_flat(n_check + 4096, 0, part)[:n_check].tolist()
        t0 = time.perf_counter()
        want = ost.format_tsv(ost.columns(), ost.rows(counts, "chr20s"))
        dt_ref = time.perf_counter() - t0
        with open(out_path) as f:
            got = "".join(next(f) for _ in range(n_check + 1))
        assert got == want[:len(got)] and len(got) > n_check, "streamed TSV differs from the oracle's text"
        rec["oracle_guard"] = f"header + first {n_check} rows equal the oracle's text byte for byte"
        rec["reference"] = {"value": n_check / dt_ref, "unit": "rows/s", "cores": 1, "kind": "port",
                            "sample": f"python port of get_stats + str(round(x, 3)) (main.py:14-79,454-466) on the first {n_check} "
                                      "positions; counting and BAM decode not included"}
    return rec


def region_bam_end_to_end(rank, world, local, barrier, max_over_ranks, sum_over_ranks, ref_len=4_000_000):
    """BASELINE configs[4] through the BAM path, scaled to a 4 Mb reference (800 k reads x 150 bp; writing a
    64 Mb BAM with the in-repo writer would take minutes): every rank opens ITS region of one coordinate-sorted,
    BAI-indexed BAM (csrc/bam_index.h: only that region's BGZF blocks are read and inflated), counts it, merges the
    halos (bc_halo_merge) and gets the all-reduced --summarise numbers.  Wall clock from file path to numbers."""
    import torch.distributed as dist
    from basecount_b200 import bamio, synth
    from basecount_b200 import dist as bdist
    from basecount_b200.engine import Engine
    path = f"/tmp/bc_bench_region_{ref_len}.bam"
    if rank == 0 and not os.path.exists(path + ".bai"):
        rec = synth.uniform_short_read_sample(seed=5, ref_len=ref_len, n_reads=ref_len * 30 // 150, read_len=150, ref_name="chr20s")
        bamio.write_bam(path, rec)
        bamio.write_bai(path)
    barrier()
    threads = max(1, (os.cpu_count() or 1) // max(world, 1))
    eng = Engine(local)
    bdist.engine_comm(eng, dist, rank, world)
    be = bdist.GpuBackend(eng, None, native_comm=True)
    best, n_total = None, 0
    for it in range(3):
        barrier()
        t0 = time.perf_counter()
        _, n = bdist.count_region_sharded_bam(be, dist, rank, world, path, 0, ref_len, threads=threads)
        pc, depth, ent = bdist.summary_region_sharded(be, dist, world, ref_len)
        dt = max_over_ranks(time.perf_counter() - t0)
        best = dt if best is None else min(best, dt)
        n_total = int(sum_over_ranks(float(n)))
    eng.comm_destroy()
    eng.close()
    return {"seconds": best, "reads": n_total, "aligned_bases": int(round(float(depth) * ref_len)), "ref_len": ref_len,
            "value": float(depth) * ref_len / best, "unit": UNIT, "bam_mb": os.path.getsize(path) / 1e6,
            "host_threads_per_rank": threads,
            "decoder": "native region fetch through the BAI index, one region per rank; best of 3"}


# ----------------------------------------------------------------------------- measuring one workload
def _peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def _pinned_out(n_slots):
    from basecount_b200 import _lib as bclib
    return (bclib.pinned_empty(n_slots, np.int64), bclib.pinned_empty(n_slots, np.int64),
            bclib.pinned_empty(n_slots, np.float64))


def oracle_guard(eng, workload, sets, ref_lens):
    """The timed configuration must produce the ORACLE's counts (oracle/ is the checker here, never what is timed):
    one sample of the step through oracle.bcount_flat, cell for cell.  For config 5 the first megabase -- reads are
    taken by start, and a read that starts at or beyond a column never touches it, so the prefix is exact."""
    from oracle import bcount as obc
    from basecount_b200 import synth
    b = sets[0][0]
    got = eng.counts(0)
    if workload == "cfg5":
        cut = min(1_000_000, ref_lens[0])
        idx = np.flatnonzero(b.starts < cut)
        part = synth.take_batch(b, idx)
        want = obc.bcount_flat(min(ref_lens[0], cut + 4096), 0, part).astype(np.int64)[:cut]
        assert np.array_equal(got[:cut], want), "config 5: the first megabase differs from the oracle"
        return f"first {cut} columns of the matrix ({idx.size} reads) equal oracle.bcount_flat"
    want = obc.bcount_flat(ref_lens[0], 0, b).astype(np.int64)
    assert np.array_equal(got, want), f"{workload}: sample 0 differs from the oracle"
    return f"sample 0 of the step ({b.n} reads, {ref_lens[0]} x 6 cells) equals oracle.bcount_flat"


def measure_workload(args, workload, rank, local, world, barrier, max_over_ranks, sum_over_ranks, steps, warmup,
                     want_clocks=True):
    """value / e2e / roofline of one workload on this rank's GPU (sample-sharded over ranks: no data-path collective)."""
    from basecount_b200.engine import Engine
    from basecount_b200.pack import pack_batches
    sets, ref_lens, label = build_workload(args, rank, workload)
    eng = Engine(local)
    eng.set_count_variant(args.variant)
    eng.begin(ref_lens)
    packed = [pack_batches(s, 0, pinned=True) for s in sets]
    resident = [eng.upload(p) for p in packed]
    bases_per_step = [p.aligned_bases for p in packed]
    alg_bytes = [p.algorithmic_bytes(ref_lens) for p in packed]
    n_slots = len(ref_lens)
    d2h_bytes = n_slots * 24
    outs = [_pinned_out(n_slots) for _ in range(max(steps, warmup, 1))]

    tiles = None
    if workload == "cfg3":
        # BASELINE configs[2] is --summarise-with-bed: the 98-amplicon scheme's windows go through K3 every step
        from basecount_b200 import synth as _synth
        from basecount_b200.scheme import load_scheme
        bed = f"/tmp/bc_bench_artic_like_{rank}.bed"
        _synth.artic_like_bed(bed)
        sch = load_scheme(bed)
        tiles = ([t[2]["inside_start"] for t in sch], [t[2]["inside_end"] for t in sch])
        amp_outs = [(np.empty((6, len(sch)), np.float64), np.zeros(len(sch), np.uint8)) for _ in outs]
        d2h_bytes += len(sch) * 49                # six float64 vectors + the empty-window flags

    def step(i, out, src):
        """Queue one step; results land in the pinned `out` arrays (valid after eng.sync())."""
        eng.reset()
        eng.push(src[i % len(src)])               # resident batch, or pinned host SoA -> H2D -> K1
        eng.summary_async(out, False)             # K2 + K3 on the summary stream, D2H of the per-sample scalars
        if tiles is not None:                     # K2 rows + K3 segmented mean / median, results at the next sync
            eng.amplicons_async(0, tiles[0], tiles[1], *amp_outs[i % len(amp_outs)])

    step(0, outs[0], resident)
    eng.sync()
    guard = oracle_guard(eng, workload, sets, ref_lens)
    nz, cs = outs[0][0].copy(), outs[0][1].copy()
    c0 = eng.counts(0)
    assert int(cs[0]) == int(c0[:, :5].sum()) and int(nz[0]) == int((c0[:, :5].sum(axis=1) != 0).sum())

    # ---- value: inputs resident in HBM; the K steps are queued back to back (each step's
    #      summary is copied to its own pinned slot) and the stream is drained at the end
    clocks = ClockSampler(local) if want_clocks and not os.environ.get("BENCH_NO_CLOCKS") else None
    if clocks:
        clocks.start()
    for i in range(warmup):
        step(i, outs[i], resident)
    eng.sync()
    if clocks:
        clocks.wait_first()
    barrier()
    launches0 = eng.kernel_launches()
    eng.timer_start()
    t_wall0 = time.perf_counter()
    for i in range(steps):
        step(i, outs[i], resident)
    t_queue = time.perf_counter() - t_wall0       # what the host needed to queue the steps (nothing waits inside)
    ms_dev = eng.timer_stop()
    eng.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.kernel_launches() - launches0
    hist = eng.count_kernel_ms_history(min(steps, 256))          # most recent first
    k1_ms = [(m, (steps - 1 - k) % len(resident)) for k, m in enumerate(hist)]
    for i in range(steps):                                        # every step must have produced its summary
        assert int(outs[i][1].sum()) > 0
    ms_dev = max_over_ranks(ms_dev)
    total_bases = sum_over_ranks(float(sum(bases_per_step[i % len(resident)] for i in range(steps))))
    value = total_bases / (ms_dev * 1e-3)
    clk = None
    if clocks:
        # the timed region is only milliseconds long, so keep the same steps running until nvidia-smi
        # (100 ms period) has seen ~1.5 s of this load
        t_end = time.perf_counter() + 1.5
        i = 0
        while time.perf_counter() < t_end:
            step(i, outs[i % len(outs)], resident)
            i += 1
            if i % 64 == 0:
                eng.sync()
        eng.sync()
        clk = clocks.stop()
        clk["window"] = "timed region plus the same steps repeated for 1.5 s"

    # ---- e2e: pinned host buffers through the C ABI, copies inside the timed region
    for i in range(min(warmup, 3)):
        step(i, outs[i], packed)
    eng.sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i, outs[i], packed)
        eng.sync()                                # the caller reads this step's result before the next
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_bases / e2e_s
    h2d = int(np.mean([packed[i % len(packed)].h2d_bytes() for i in range(steps)]))

    # ---- roofline of the counting kernel (K1): algorithmic bytes / its own device time.  In the pipeline of steps the
    #      summary of step i runs on its own stream beside the K1 of step i + 1 (and takes SM time from it), so the
    #      kernel is also timed with nothing beside it -- zero, count, zero, count ... -- and THAT is the roofline figure;
    #      its time inside the pipeline is reported next to it.
    peak, peak_src = _peak()
    k1_pipeline_ms = float(np.mean([m for m, _ in k1_ms]))
    n_iso = min(max(steps, 4), 20)
    for i in range(n_iso):
        eng.reset()
        eng.push(resident[i % len(resident)])
    eng.sync()
    hist = eng.count_kernel_ms_history(n_iso)
    k1_ms = [(m, (n_iso - 1 - k) % len(resident)) for k, m in enumerate(hist)]
    k1_avg_ms = float(np.mean([m for m, _ in k1_ms]))
    k1_bytes = float(np.mean([alg_bytes[j] for _, j in k1_ms]))
    achieved = k1_bytes / (k1_avg_ms * 1e-3) / 1e9
    k1_bases = float(np.mean([bases_per_step[j] for _, j in k1_ms]))
    traffic = NCU_TRAFFIC_BYTES.get(workload) if (args.variant == 0 and _is_default_shape(args, workload)) else None
    rec = {
        "value": value, "unit": UNIT, "ms_per_step": ms_dev / steps, "steps": steps, "warmup": warmup,
        "config": {"workload": label, "samples_per_gpu": n_slots, "reads_per_step_per_gpu": packed[0].n_reads,
                   "aligned_bases_per_step_per_gpu": bases_per_step[0], "ref_len": ref_lens[0],
                   "l2": f"inputs {packed[0].h2d_bytes() / 1e6:.0f} MB per step"
                         + (" > 126 MB L2; " if packed[0].h2d_bytes() > 126e6 else "; ")
                         + (f"{len(resident)} resident batches alternate" if len(resident) > 1 else
                            "count planes + inputs exceed L2" if ref_lens[0] * 24 + packed[0].h2d_bytes() > 126e6 else
                            "the accumulators are re-zeroed (both sets alternate) every step"),
                   "timing": "CUDA events on the engine's compute stream; max over ranks", "k1_variant": args.variant,
                   "oracle_guard": guard},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_s / steps, "timing": "host wall clock, device-synchronised both sides",
                "h2d_gbs_step": h2d / (e2e_s / steps) / 1e9},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this "
                                       "kernel on this workload (profiles/r4_a_k1_summary.md); null if not captured",
                     "kernel": {0: "k1_count_fast (+ k1_count_tiled over what it defers: nothing on this workload)",
                                1: "k1_count_per_base", 2: "k1_count_tiled"}[args.variant],
                     "kernel_ms": k1_avg_ms, "kernel_ms_in_pipeline": k1_pipeline_ms,
                     "kernel_timing": "CUDA events around the kernel on its stream (library event ring), mean of "
                                      f"{n_iso} launches with only the accumulator reset between them; in_pipeline = the same "
                                      "inside the timed steps, where the previous step's summary kernel shares the SMs",
                     "algorithmic_bytes_per_launch": k1_bytes,
                     "bytes_per_aligned_base": k1_bytes / k1_bases, "kernel_aligned_bases_per_s": k1_bases / (k1_avg_ms * 1e-3),
                     "step_over_kernel": (value / max(world, 1)) / (k1_bases / (k1_avg_ms * 1e-3)),
                     "peak_source": peak_src},
        "gpu_launches": int(launches), "clocks": clk, "wall_s_timed_region": t_wall,
        "host_queue_us_per_step": 1e6 * t_queue / steps,
    }
    state = {"eng": eng, "resident": resident, "sets": sets, "ref_lens": ref_lens, "packed": packed, "h2d": h2d}
    return rec, state


def _is_default_shape(args, workload):
    if workload == "cfg2x12":
        return args.samples_per_gpu == SAMPLES_PER_GPU and args.reads_per_sample == 124_000
    if workload == "cfg5":
        return args.cfg5_reads == 12_888_833
    return True


def release(state):
    for r in state["resident"]:
        r.free()
    state["eng"].close()


def h2d_control(eng, h2d_bytes, barrier, max_over_ranks, reps=8):
    """Plain pinned cudaMemcpyAsync of one step's input bytes on every rank at once (bc_h2d_probe: cudaHostAlloc
    memory, the engine's copy stream, CUDA events): what the platform gives the end-to-end path to work with
    (PCIe / host memory shared by the ranks of one box).  Minimum over ranks."""
    barrier()
    g = eng.h2d_probe(h2d_bytes, reps)
    return -max_over_ranks(-g)


def cpu_baseline_cfg3(batch, ref_len):
    """Reference CPU path of --summarise-with-bed on a quarter of the step's reads, one core: compiled bcount
    (oracle/_ref) + the oracle's port of get_stats, the summarise block and the amplicon block (main.py:501-595)."""
    from oracle import bcount as obc
    from oracle import stats as ost
    from basecount_b200 import synth
    use_ref = obc.load_ref_bcount() is not None
    part = synth.take_batch(batch, np.arange(0, batch.n, 4))
    lists = part.to_lists()
    bed = "/tmp/bc_bench_artic_like_cpu.bed"
    synth.artic_like_bed(bed)
    t0 = time.perf_counter()
    counts = obc.load_ref_bcount()(ref_len, 0, *lists) if use_ref else obc.bcount_flat(ref_len, 0, part).tolist()
    t1 = time.perf_counter()
    cov, ent, sec = ost.per_position_vectors(counts)
    ost.summary(cov, ent, ref_len)
    win = [(d["inside_start"], d["inside_end"]) for _, _, d in ost.scheme_windows(bed)]
    ost.amplicon_vectors(cov, ent, sec, win)
    t2 = time.perf_counter()
    bases = part.aligned_bases()
    return {"value": bases / (t2 - t0), "unit": UNIT, "cores": 1, "kind": "reference+port" if use_ref else "port",
            "sample": f"every 4th read of one step ({part.n} reads, {bases} aligned bases): bcount {t1 - t0:.2f}s + python "
                      f"port of get_stats / summarise / amplicon block {t2 - t1:.2f}s; Python-list inputs prebuilt"}


def cpu_baseline_cfg5(batch, ref_len):
    """Reference CPU path of --summarise on the first megabase of the reference, one core (the full 64.4 Mb
    would take ~12 minutes in get_stats alone): rate = aligned bases of that prefix / time."""
    from oracle import bcount as obc
    from oracle import stats as ost
    from basecount_b200 import synth
    use_ref = obc.load_ref_bcount() is not None
    cut = min(1_000_000, ref_len)
    part = synth.take_batch(batch, np.flatnonzero(batch.starts < cut - 200))
    lists = part.to_lists()
    t0 = time.perf_counter()
    counts = obc.load_ref_bcount()(cut, 0, *lists) if use_ref else obc.bcount_flat(cut, 0, part).tolist()
    t1 = time.perf_counter()
    cov, ent, _ = ost.per_position_vectors(counts)
    ost.summary(cov, ent, cut)
    t2 = time.perf_counter()
    bases = part.aligned_bases()
    return {"value": bases / (t2 - t0), "unit": UNIT, "cores": 1, "kind": "reference+port" if use_ref else "port",
            "sample": f"the first {cut} positions ({part.n} reads, {bases} aligned bases; the rate is what the full "
                      f"reference would run at, extrapolated): bcount {t1 - t0:.2f}s + python port of get_stats / "
                      f"summarise {t2 - t1:.2f}s; Python-list inputs prebuilt"}


# ----------------------------------------------------------------------------- config 5 over N GPUs
def measure_region_sharded(args, rank, world, local, barrier, max_over_ranks, sum_over_ranks, steps, warmup):
    """One reference cut into `world` regions (strong scaling).  A step = restore the halo
    columns, zero the accumulators, count the rank's reads (resident in HBM), bc_halo_merge (NCCL send/recv of the
    halo plane segments to the ranks that own them, add what arrives, cut the slot to the owned columns), the
    --summarise reductions and the all-reduce of the three scalars -- all enqueued on the engine's compute
    stream by the C-ABI library: no host synchronisation inside a step, the K steps are queued back to back."""
    import torch.distributed as dist
    from basecount_b200 import dist as bdist
    from basecount_b200 import synth
    from basecount_b200.engine import Engine
    from basecount_b200.pack import pack_batches
    sets, ref_lens, label = build_workload(args, rank, "cfg5")
    local_reads, ref_len = sets[0][0], ref_lens[0]
    bounds = bdist.region_bounds(ref_len, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    own = hi - lo
    h = bdist.halo_columns(local_reads, own, ref_len - hi)
    eng = Engine(local)
    bdist.engine_comm(eng, dist, rank, world)                 # the process group only carries the 128-byte id
    halos = eng.allgather_u32(h)                              # known once the reads are packed
    eng.begin([own + h])
    packed = pack_batches([local_reads], 0, pinned=True)
    resident = eng.upload(packed)
    bases = packed.aligned_bases
    outs = [_pinned_out(1) for _ in range(max(steps, warmup, 1))]

    def step(i, src):
        eng.set_length(0, own + h)                            # the halo columns are back (device-side write)
        eng.reset()
        eng.push(src)
        eng.halo_merge(0, bounds, halos)
        eng.summary_allreduce_async(outs[i % len(outs)], False)

    step(0, resident)
    eng.sync()
    # checks: every aligned base of every rank is in exactly one cell of the merged matrix; every rank holds the
    # same all-reduced summary; rank 0's columns equal the oracle's (nothing reaches it from the left)
    total_bases = sum_over_ranks(float(bases))
    got = eng.counts(0)
    cells = sum_over_ranks(float(got.sum()))
    assert cells == total_bases, (cells, total_bases)
    mine = [float(outs[0][0][0]), float(outs[0][1][0]), float(outs[0][2][0])]
    for v in mine:
        assert max_over_ranks(v) == v and -max_over_ranks(-v) == v, "ranks disagree on the all-reduced summary"
    assert int(mine[1]) == int(total_bases - sum_over_ranks(float(got[:, 5].sum())))
    guard = None
    if rank == 0:
        from oracle import bcount as obc
        cut = min(1_000_000, own)
        part = synth.take_batch(local_reads, np.flatnonzero(local_reads.starts < cut))
        want = obc.bcount_flat(cut + 4096, 0, part).astype(np.int64)[:cut]
        assert np.array_equal(got[:cut], want), "region-sharded: rank 0's first megabase differs from the oracle"
        guard = f"rank 0: first {cut} owned columns equal oracle.bcount_flat; cells over ranks == aligned bases"
    clocks = ClockSampler(local)
    clocks.start()
    for i in range(warmup):
        step(i, resident)
    eng.sync()
    clocks.wait_first()
    barrier()
    launches0 = eng.kernel_launches()
    eng.timer_start()
    for i in range(steps):
        step(i, resident)
    ms = eng.timer_stop()
    eng.sync()
    barrier()
    ms = max_over_ranks(ms)
    launches = eng.kernel_launches() - launches0
    t_end = time.perf_counter() + 1.0                          # let nvidia-smi see the load
    i = 0
    while time.perf_counter() < t_end:
        step(i, resident)
        i += 1
        if i % 32 == 0:
            eng.sync()
    eng.sync()
    clk = clocks.stop()
    hist = eng.count_kernel_ms_history(min(steps, 256))
    k1_ms = max_over_ranks(float(np.mean(hist)))
    # e2e: the same step from pinned host buffers (H2D inside), result read back every step
    step(0, packed)
    eng.sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i, packed)
        eng.sync()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    peak, peak_src = _peak()
    alg = packed.algorithmic_bytes([own + h])
    achieved = alg / (k1_ms * 1e-3) / 1e9
    rec = {"value": total_bases * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms / steps, "scaling": "strong",
           "config": {"workload": label, "ref_len": ref_len, "reads_total": int(sum_over_ranks(float(packed.n_reads))),
                      "aligned_bases_total": total_bases, "halo_columns": halos,
                      "collective": "bc_halo_merge: ncclSend/ncclRecv of the halo plane segments + k_halo_add, then "
                                    "ncclAllReduce of 3 scalars per step, all on the compute stream (library-owned communicator)",
                      "l2": f"inputs {packed.h2d_bytes() / 1e6:.0f} MB per rank per step; count planes {own * 24 / 1e6:.0f} MB per rank",
                      "timing": "CUDA events on the engine's compute stream around K steps queued back to back; max over ranks",
                      "oracle_guard": guard},
           "e2e": {"value": total_bases * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": packed.h2d_bytes(),
                   "d2h_bytes_per_step": 24, "ms_per_step": 1e3 * e2e_s / steps,
                   "timing": "host wall clock, device-synchronised both sides"},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": None, "kernel": "k1_count_fast", "kernel_ms": k1_ms, "algorithmic_bytes_per_launch": alg,
                        "note": "per rank (max over ranks of the mean K1 time)", "peak_source": peak_src},
           "gpu_launches": int(launches), "clocks": clk}
    resident.free()
    eng.comm_destroy()
    eng.close()
    return rec


# ----------------------------------------------------------------------------- our arm
def main():
    args = parse()
    claim_stdout()
    if args.gen_samples:
        for seed in args.gen_samples.split(","):
            _generate_sample((int(seed), args.reads_per_sample))
        return
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: basecount_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    import __graft_entry__ as ge
    ge.build()
    from basecount_b200.hostbind import bind_to_device_node

    numa = bind_to_device_node(local)              # pinned buffers on the GPU's NUMA node (matters for e2e at N > 1)
    comm = (barrier, max_over_ranks, sum_over_ranks)
    default_line = args.workload is None
    workload = args.workload or "cfg2x12"

    if workload == "cfg5" and world > 1:           # explicit: the region-sharded run as the line itself
        rec = measure_region_sharded(args, rank, world, local, *comm, args.steps, args.warmup)
        line = {"metric": METRIC, "higher_is_better": True, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "cpu_baseline": None, **rec}
        if rank == 0:
            emit(line)
        dist.destroy_process_group()
        return

    rec, state = measure_workload(args, workload, rank, local, world, *comm, args.steps, args.warmup)
    rec["config"]["host_numa"] = numa
    line = {"metric": METRIC, "value": rec.pop("value"), "unit": rec.pop("unit"), "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": rec.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic"}
    rec.pop("steps"), rec.pop("warmup")
    line.update(rec)
    # what the platform gives the end-to-end path: plain pinned copies of the same bytes, all ranks at once
    ctl = h2d_control(state["eng"], state["h2d"], barrier, max_over_ranks)
    line["e2e"]["h2d_gbs_control"] = ctl
    line["e2e"]["note"] = ("end to end is bound by the host-to-device copy: h2d_gbs_step is what a step achieves per GPU, "
                           "h2d_gbs_control a plain pinned cudaMemcpyAsync of the same bytes on every rank at once")
    if rank == 0 and world == 1 and not args.no_cpu_baseline and workload == "cfg2x12":
        line["cpu_baseline"] = cpu_baseline(state["sets"][0], state["ref_lens"][0])
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0 and world == 1 and workload == "cfg2x12":
        for r in state["resident"]:
            r.free()
        state["resident"] = []
        line["bam_e2e"] = bam_end_to_end(state["eng"], not args.no_cpu_baseline)
    release(state)
    if rank == 0 and world == 1 and default_line:
        line["tsv_stream"] = tsv_stream_end_to_end(not args.no_cpu_baseline)

    if default_line and world == 1:
        # the other single-GPU shapes of BASELINE.json (configs[2] and configs[4]) as sub-records of the same line
        sub = {}
        for w, cpu in (("cfg3", cpu_baseline_cfg3), ("cfg5", cpu_baseline_cfg5)):
            r, st = measure_workload(args, w, rank, local, world, *comm, min(args.steps, 10), min(args.warmup, 3),
                                     want_clocks=False)
            r["cpu_baseline"] = None if args.no_cpu_baseline else cpu(st["sets"][0][0], st["ref_lens"][0])
            release(st)
            sub[w] = r
        line["configs"] = sub
    if default_line and world > 1:
        # the split with a data-path collective (configs[4], strong scaling) beside the sample-sharded line
        line["region_sharded"] = measure_region_sharded(args, rank, world, local, *comm, min(args.steps, 10),
                                                        min(args.warmup, 3))
        line["region_sharded"]["bam_e2e"] = region_bam_end_to_end(rank, world, local, *comm)
    if default_line:
        line["bam_e2e_batch"] = bam_batch_end_to_end(local, SAMPLES_PER_GPU, *comm)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
