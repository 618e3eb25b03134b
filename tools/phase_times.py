#!/usr/bin/env python
"""Device time of each phase of a bench step (reset / count / summarise), each queued back to back
on the engine's compute stream and timed with CUDA events.  Diagnostic only:
    python tools/phase_times.py [--reps 50]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--samples", type=int, default=12)
    args = ap.parse_args()
    import bench
    from basecount_b200 import _lib as bclib
    from basecount_b200.engine import Engine
    from basecount_b200.pack import pack_batches
    sets = [bench.make_samples(range(100 + k * args.samples, 100 + (k + 1) * args.samples), 124_000) for k in range(2)]
    ref_lens = [29903] * args.samples
    eng = Engine(0)
    eng.begin(ref_lens)
    resident = [eng.upload(pack_batches(s, 0, pinned=True)) for s in sets]
    outs = [(bclib.pinned_empty(args.samples, np.int64), bclib.pinned_empty(args.samples, np.int64),
             bclib.pinned_empty(args.samples, np.float64)) for _ in range(args.reps)]

    def timed(name, fn):
        for i in range(3):
            fn(i)
        eng.sync()
        eng.timer_start()
        t0 = time.perf_counter()
        for i in range(args.reps):
            fn(i)
        t_host = time.perf_counter() - t0             # what the host needs to QUEUE the iterations (no synchronisation inside)
        ms = eng.timer_stop()
        eng.sync()
        print(f"{name:34s} {1e3 * ms / args.reps:8.2f} us per iteration   (host queues one in {1e6 * t_host / args.reps:6.2f} us)")

    timed("reset", lambda i: eng.reset())
    timed("count (K1 + corrections + check)", lambda i: eng.push(resident[i % 2]))
    eng.reset()
    eng.push(resident[0])
    timed("summarise (K2)", lambda i: eng.summary_async(outs[i], False))
    timed("reset + count", lambda i: (eng.reset(), eng.push(resident[i % 2])))
    timed("reset + count + summarise (step)", lambda i: (eng.reset(), eng.push(resident[i % 2]), eng.summary_async(outs[i], False)))
    hist = eng.count_kernel_ms_history(args.reps)
    print(f"{'K1 alone (event ring)':34s} {1e3 * float(np.mean(hist)):8.2f} us")


if __name__ == "__main__":
    main()
