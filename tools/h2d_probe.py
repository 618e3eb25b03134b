#!/usr/bin/env python
"""Plain pinned host-to-device copy rate of this box (bc_h2d_probe: cudaHostAlloc memory, the engine's copy stream,
CUDA events) -- the ceiling of the end-to-end path, which is bound by that copy.
    python tools/h2d_probe.py [MB ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from basecount_b200.engine import Engine
    eng = Engine(0)
    for mb in [int(a) for a in sys.argv[1:]] or [184, 1024]:
        print(f"pinned H2D {mb} MB: {eng.h2d_probe(mb << 20, 8):.2f} GB/s", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
