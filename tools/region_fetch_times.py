#!/usr/bin/env python
"""Host-side decode time of the region-sharded path: whole-file decode against index-aware region fetch
 on a config-5-shaped BAM scaled down to `--ref-len` (uniform 150 bp reads at 30x).
CPU only (the decoder is host code of the C-ABI library):
    python tools/region_fetch_times.py [--ref-len 8000000] [--world 8]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-len", type=int, default=8_000_000)
    ap.add_argument("--world", type=int, default=8)
    args = ap.parse_args()
    from basecount_b200 import bamio, synth
    from basecount_b200.build import build
    build()
    n_reads = args.ref_len * 30 // 150
    path = f"/tmp/bc_region_demo_{args.ref_len}.bam"
    if not os.path.exists(path + ".bai"):
        rec = synth.uniform_short_read_sample(seed=5, ref_len=args.ref_len, n_reads=n_reads, read_len=150, ref_name="chr20s")
        bamio.write_bam(path, rec)
        t = time.perf_counter()
        bamio.write_bai(path)
        t_index = time.perf_counter() - t
    else:
        t_index = None

    def best(fn, reps=3):
        out = []
        for _ in range(reps):
            t = time.perf_counter()
            r = fn()
            out.append(time.perf_counter() - t)
        return min(out), r

    def whole():
        nb = bamio.NativeBam(path)
        b = nb.select(0, 0)
        nb.close()
        return b.n

    bounds = np.linspace(0, args.ref_len, args.world + 1, dtype=np.int64)

    def region(r):
        nb = bamio.NativeBam(path, region=(0, int(bounds[r]), int(bounds[r + 1])))
        b = nb.select(0, 0)
        nb.close()
        return b.n

    t_whole, n_whole = best(whole)
    per_rank = [best(lambda r=r: region(r)) for r in range(args.world)]
    assert sum(n for _, n in per_rank) == n_whole
    print(json.dumps({"bam_mb": os.path.getsize(path) / 1e6, "reads_kept": n_whole, "host_cores": os.cpu_count(),
                      "index_build_s": t_index, "whole_file_decode_s": t_whole, "world": args.world,
                      "region_fetch_s_max_over_ranks": max(t for t, _ in per_rank),
                      "region_fetch_s_per_rank": [round(t, 4) for t, _ in per_rank],
                      "speedup_vs_whole_file": t_whole / max(t for t, _ in per_rank)}))


if __name__ == "__main__":
    main()
