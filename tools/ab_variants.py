#!/usr/bin/env python
"""Build experimental variants of the library next to the product build and print / run the A/B command.

A variant is NAME:DEFINE[,DEFINE...] (the -D macros csrc/*.cuh read: BC_K1_*, BC_K1F_*, BC_K2_PER_CTA, ...), built
into basecount_b200/csrc/variants/libNAME.so (git-ignored, travels to the GPU box).  On the box every variant
runs the counting parity tests and tools/phase_times.py through BASECOUNT_B200_LIB, then the product build runs
phase_times.py on the same box for comparison (box-to-box spread is +-3 %, so only same-box numbers compare).

    python tools/ab_variants.py mask0:BC_K1F_MASK3=0 regs160:BC_K1F_REGS=160 k2_512:BC_K2_PER_CTA=512
    python tools/ab_variants.py --env BASECOUNT_B200_K1=walker --run seq416:BC_K1_SEQCAP=416,BC_K1_CIGCAP=96

--run executes the command here (on a GPU box); without it the script prints the shell text to hand to gpurun.
"""
import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="+", help="NAME:DEFINE[,DEFINE...]")
    ap.add_argument("--env", action="append", default=[], help="KEY=VALUE exported for the variant runs (e.g. BASECOUNT_B200_K1=walker)")
    ap.add_argument("--tests", default="tests/test_gpu_counts.py", help="pytest target run for every variant")
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--run", action="store_true")
    args = ap.parse_args()
    from basecount_b200 import build as b
    vdir = os.path.join(b.CSRC, "variants")
    os.makedirs(vdir, exist_ok=True)
    jobs = []
    for v in args.variants:
        name, _, defs = v.partition(":")
        jobs.append((name, [d for d in defs.split(",") if d]))

    def one(job):
        name, defs = job
        out = os.path.join(vdir, f"lib{name}.so")
        b.build(force=True, out=out, defines=defs)
        return out

    with ThreadPoolExecutor(max_workers=min(4, len(jobs))) as ex:
        for out in ex.map(one, jobs):
            print("built", os.path.relpath(out, ROOT), file=sys.stderr)
    b.build()                                        # the product build the variants are compared with
    env = " ".join(f"{e}" for e in args.env)
    lines = ["mkdir -p gpurun_out; rm -f gpurun_out/ab.log"]
    for name, _ in jobs:
        lines += [f'echo "=== {name}" >> gpurun_out/ab.log',
                  f"(env {env} BASECOUNT_B200_LIB=$PWD/basecount_b200/csrc/variants/lib{name}.so timeout -k 5 120 "
                  f"python -m pytest {args.tests} -m gpu -x -q) 2>&1 | tail -2 >> gpurun_out/ab.log",
                  f"(env {env} BASECOUNT_B200_LIB=$PWD/basecount_b200/csrc/variants/lib{name}.so timeout -k 5 120 "
                  f"python tools/phase_times.py --reps {args.reps}) 2>&1 >> gpurun_out/ab.log"]
    lines += ['echo "=== product build" >> gpurun_out/ab.log',
              f"(timeout -k 5 120 python tools/phase_times.py --reps {args.reps}) 2>&1 >> gpurun_out/ab.log",
              "cat gpurun_out/ab.log"]
    script = "\n".join(lines)
    if args.run:
        sys.exit(subprocess.run(["bash", "-c", script], cwd=ROOT).returncode)
    print(script)


if __name__ == "__main__":
    main()
