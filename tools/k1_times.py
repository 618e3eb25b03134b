#!/usr/bin/env python
"""K1's own device time (the library's event ring around the counting kernel) on the bench workloads, resident
inputs, reset + count queued back to back.  Cheap same-box A/B of library builds:
    BASECOUNT_B200_LIB=/path/to/variant.so python tools/k1_times.py [--workloads cfg2x12,cfg3,cfg5] [--reps 30]
Every run also checks that the cells of sample 0 add up to its aligned bases."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="cfg2x12,cfg3,cfg5")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--cfg5-reads", type=int, default=0, help="a shorter config 5 (the reference length scales with it)")
    a = ap.parse_args()
    import bench
    from basecount_b200 import _lib as bclib
    from basecount_b200.engine import Engine
    from basecount_b200.pack import pack_batches
    sys.argv = sys.argv[:1]
    args = bench.parse()
    if a.cfg5_reads:
        args.cfg5_reads = a.cfg5_reads
    name = os.path.basename(bclib.LIB_PATH)
    for wl in a.workloads.split(","):
        sets, ref_lens, label = bench.build_workload(args, 0, wl)
        eng = Engine(0)
        eng.begin(ref_lens)
        packed = [pack_batches(s, 0, pinned=True) for s in sets]
        resident = [eng.upload(p) for p in packed]
        for i in range(3):
            eng.reset()
            eng.push(resident[i % len(resident)])
        eng.sync()
        for i in range(a.reps):
            eng.reset()
            eng.push(resident[i % len(resident)])
        eng.sync()
        hist = np.asarray(eng.count_kernel_ms_history(a.reps))
        got = eng.counts(0)
        last = (a.reps - 1) % len(resident)
        ok = int(got.sum()) == int(sets[last][0].aligned_bases())           # sample 0 of the last batch pushed
        alg = packed[0].algorithmic_bytes(ref_lens)
        peak, _ = bench._peak()
        print(f"{name:28s} {wl:8s} K1 mean {1e3 * hist.mean():8.2f} us  min {1e3 * hist.min():8.2f} us  frac {alg / (hist.mean() * 1e-3) / 1e9 / peak:.3f}  cells==bases {ok}", flush=True)
        for r in resident:
            r.free()
        eng.close()


if __name__ == "__main__":
    main()
