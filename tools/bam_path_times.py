#!/usr/bin/env python
"""Where the BAM-path-to-summary wall clock goes for one config-1 sample (124,000 reads x 400 bp):
inflate + record index (bc_bam_open), selection + 2-bit packing (bc_bam_pack_*), push + summarise.
    python tools/bam_path_times.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from basecount_b200 import bamio, synth
    from basecount_b200.build import build
    build()
    path = "/tmp/bc_bench_cfg1_seed100.bam"
    if not os.path.exists(path):
        bamio.write_bam(path + ".tmp", synth.amplicon_sample(seed=100))
        os.replace(path + ".tmp", path)
    eng = None
    try:
        from basecount_b200.engine import Engine
        eng = Engine(0)
    except Exception as e:                                 # no GPU: host phases only
        print("no engine:", e)
    print(f"host cores {os.cpu_count()}, BAM {os.path.getsize(path) / 1e6:.1f} MB")
    for it in range(4):
        t0 = time.perf_counter()
        nb = bamio.NativeBam(path)
        t1 = time.perf_counter()
        nb.core()
        t2 = time.perf_counter()
        p = nb.pack(0, 0, 0)                               # selection + trimming + 2-bit packing, one native pass
        nb.close()
        t3 = time.perf_counter()
        t4 = t3
        if eng is not None:
            eng.begin([nb.ref_lengths[0]])
            eng.push(p)
            eng.sync()
            eng.summary(False)
            t4 = time.perf_counter()
        print(f"open {1e3 * (t1 - t0):6.1f}  core {1e3 * (t2 - t1):6.1f}  select+pack {1e3 * (t3 - t2):6.1f}  "
              f"count+summarise {1e3 * (t4 - t3):6.1f}  total {1e3 * (t4 - t0):6.1f} ms")


if __name__ == "__main__":
    main()
