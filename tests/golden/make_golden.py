#!/usr/bin/env python
"""Generate tests/golden/*.json|npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    make -C oracle ref && PYTHONHASHSEED=0 python tests/golden/make_golden.py

What runs here is the reference's own code, imported from where it lies:
  * count.bcount            -> oracle/_ref/count*.so, compiled from basecount/count.cpp
  * get_stats, run(), ...   -> /root/reference/basecount/main.py
  * load_scheme             -> /root/reference/basecount/scheme.py
pysam is not installable in this image, so `import pysam` (main.py:2) is satisfied by
the small in-memory stand-in below, which serves the attributes main.py:97-99,122,127,
165-173,204 touch from a `Records` object.  Nothing from the reference is copied.
"""
import array
import contextlib
import copy
import gzip
import io
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))     # module `count`
sys.path.insert(0, "/root/reference")

from basecount_b200 import synth                              # noqa: E402
from basecount_b200.records import FLAG_UNMAPPED, Records, _leading_trailing_clips   # noqa: E402

# ---------------------------------------------------------------- pysam stand-in
_REGISTRY = {}


class _Read:
    __slots__ = ("is_unmapped", "mapping_quality", "reference_name", "reference_start",
                 "query_alignment_sequence", "query_alignment_qualities", "cigartuples")


class _AlignmentFile:
    def __init__(self, path, mode="rb"):
        self._rec = _REGISTRY[path]
        self.references = tuple(self._rec.ref_names)
        self.lengths = tuple(self._rec.ref_lengths)

    def fetch(self, until_eof=False):
        rec = self._rec
        lead, trail = _leading_trailing_clips(rec)
        seq_b = rec.seq.tobytes()
        for i in range(rec.n):
            r = _Read()
            r.is_unmapped = bool(rec.flag[i] & FLAG_UNMAPPED)
            r.mapping_quality = int(rec.mapq[i])
            r.reference_name = rec.ref_names[rec.ref_id[i]] if rec.ref_id[i] >= 0 else None
            r.reference_start = int(rec.pos[i])
            a = int(rec.seq_off[i] + lead[i])
            b = int(rec.seq_off[i + 1] - trail[i])
            r.query_alignment_sequence = seq_b[a:b].decode("ascii")
            r.query_alignment_qualities = array.array("B", rec.qual[a:b].tobytes())
            c0, c1 = int(rec.cigar_off[i]), int(rec.cigar_off[i + 1])
            r.cigartuples = [(int(w & 0xF), int(w >> 4)) for w in rec.cigar[c0:c1]]
            yield r

    def close(self):
        pass


_pysam = types.ModuleType("pysam")
_pysam.AlignmentFile = _AlignmentFile
_pysam.set_verbosity = lambda v: 0
sys.modules["pysam"] = _pysam

import basecount.main as refmain                              # noqa: E402  (the reference)
from basecount.scheme import load_scheme as ref_load_scheme  # noqa: E402
from count import bcount as ref_bcount                       # noqa: E402


def dump_gz(name, obj):
    """Deterministic gzip'd JSON (mtime 0) to keep the fixtures small."""
    with open(os.path.join(HERE, name), "wb") as raw:
        with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as gz:
            gz.write(json.dumps(obj).encode())


def run_cli(argv):
    """Reference run() with argv; returns stdout text."""
    old = sys.argv
    sys.argv = ["basecount"] + argv
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            refmain.run()
    finally:
        sys.argv = old
    return buf.getvalue()


def records_to_npz(path, rec):
    np.savez_compressed(path, ref_names=np.array(rec.ref_names), ref_lengths=np.array(rec.ref_lengths),
                        ref_id=rec.ref_id, pos=rec.pos, mapq=rec.mapq, flag=rec.flag, cigar=rec.cigar,
                        cigar_off=rec.cigar_off, seq=rec.seq, qual=rec.qual, seq_off=rec.seq_off)


# ---------------------------------------------------------------- stats goldens
def stats_goldens():
    rng = np.random.default_rng(2024)
    readme = [[0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0], [0, 1, 2, 0, 0, 0], [0, 0, 0, 1334, 0, 0],
              [0, 1427, 0, 2, 0, 0], [1453, 0, 1, 0, 11, 0], [1471, 0, 1, 0, 0, 0], [1464, 1, 13, 0, 0, 0],
              [0, 1479, 0, 2, 1, 0], [1, 1479, 0, 3, 0, 0], [0, 1480, 0, 2, 2, 0], [1483, 0, 2, 0, 0, 0]]
    special = [[5, 5, 0, 0, 0, 0], [1, 1, 1, 1, 1, 1], [0, 0, 0, 0, 0, 7], [0, 0, 0, 0, 3, 0], [2, 0, 0, 0, 0, 9],
               [7, 7, 7, 7, 7, 0], [1, 2, 3, 4, 5, 6], [4000000000, 1, 0, 0, 0, 0], [3, 3, 3, 0, 0, 0],
               [0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 1], [123456, 654321, 111, 7, 0, 12]]
    rnd = []
    for _ in range(400):
        kind = rng.integers(0, 4)
        if kind == 0:
            row = rng.integers(0, 4, size=6)
        elif kind == 1:
            row = rng.integers(0, 3000, size=6) * (rng.random(6) < 0.5)
        elif kind == 2:
            row = np.zeros(6, dtype=np.int64)
            row[rng.integers(0, 6)] = rng.integers(1, 100000)
            row[rng.integers(0, 6)] += rng.integers(0, 30)
        else:
            row = rng.integers(0, 2 ** 31, size=6)
        rnd.append([int(x) for x in row])
    counts = readme + special + rnd
    out = {"counts": counts, "cases": []}
    for show_n in (False, True):
        for long_format in (False, True):
            rows = refmain.get_stats(copy.deepcopy(counts), "REF", show_n_bases=show_n, long_format=long_format)
            out["cases"].append({"show_n_bases": show_n, "long_format": long_format, "rows": rows})
    # README.md:20-59 -- the literal printed cells for those rows (3 d.p.), as a format golden
    out["readme_text"] = {
        "0,1,2,0,0": "3\t0\t1\t2\t0\t0\t0.0\t33.333\t66.667\t0.0\t0.0\t0.395\t0.0",
        "0,1427,0,2,0": "1429\t0\t1427\t0\t2\t0\t0.0\t99.86\t0.0\t0.14\t0.0\t0.007\t0.0",
        "1453,0,1,0,11": "1465\t1453\t0\t1\t0\t11\t99.181\t0.0\t0.068\t0.0\t0.751\t0.031\t0.207",
        "1464,1,13,0,0": "1478\t1464\t1\t13\t0\t0\t99.053\t0.068\t0.88\t0.0\t0.0\t0.035\t0.186",
        "0,1479,0,2,1": "1482\t0\t1479\t0\t2\t1\t0.0\t99.798\t0.0\t0.135\t0.067\t0.01\t0.459",
        "1,1479,0,3,0": "1483\t1\t1479\t0\t3\t0\t0.067\t99.73\t0.0\t0.202\t0.0\t0.013\t0.406",
        "0,1480,0,2,2": "1484\t0\t1480\t0\t2\t2\t0.0\t99.73\t0.0\t0.135\t0.135\t0.013\t0.5",
        "0,0,0,1334,0": "1334\t0\t0\t0\t1334\t0\t0.0\t0.0\t0.0\t100.0\t0.0\t0.0\t1",
        "0,0,0,0,0": "0\t0\t0\t0\t0\t0\t-1\t-1\t-1\t-1\t-1\t1\t1",
    }
    # cross-check the README cells through the reference's own formatting
    for key, text in out["readme_text"].items():
        c = [int(x) for x in key.split(",")] + [0]
        row = refmain.get_stats([c], "X")[0]
        got = "\t".join(str(round(x, 3)) for x in row[2:])
        assert got == text, (key, got, text)
    dump_gz("stats.json.gz", out)


# ---------------------------------------------------------------- bcount goldens
def bcount_goldens():
    kats = []

    def kat(ref_len, mbq, reads, quals, starts, ctuples):
        case = {"ref_len": ref_len, "min_base_quality": mbq, "reads": reads, "qualities": quals, "starts": starts,
                "ctuples": [[list(t) for t in c] for c in ctuples]}
        try:
            case["counts"] = ref_bcount(ref_len, mbq, reads, quals, starts, ctuples)
        except IndexError:
            case["error"] = "IndexError"
        kats.append(case)

    q5 = [[30] * 5]
    kat(10, 0, ["ACGTN"], q5, [2], [[(0, 2), (1, 1), (2, 1), (0, 2)]])          # SURVEY 8(c) KATs
    kat(10, 31, ["ACGTN"], q5, [2], [[(0, 2), (1, 1), (2, 1), (0, 2)]])
    kat(3, 0, ["ACGTN"], q5, [0], [[(0, 5)]])                                   # IndexError
    kat(4, 0, ["acRA"], [[30] * 4], [0], [[(0, 4)]])
    kat(12, 0, ["NNAC"], [[30] * 4], [1], [[(4, 3), (0, 2), (3, 2), (0, 2), (5, 1)]])
    kat(8, 0, [], [], [], [])
    kat(0, 0, [], [], [], [])
    kat(6, 0, [""], [[]], [3], [[]])
    kat(6, 0, ["AC"], [[1, 1]], [5], [[(0, 1)]])                                 # last column ok
    kat(6, 0, ["AC"], [[1, 1]], [5], [[(0, 2)]])                                 # one past -> IndexError
    kat(6, 5, ["AC"], [[1, 1]], [5], [[(0, 2)]])                                 # filtered past the end: no error
    kat(6, 0, ["AR"], [[9, 9]], [5], [[(0, 2)]])                                 # uncounted letter past the end
    kat(6, 0, ["A"], [[9]], [4], [[(0, 1), (2, 2)]])                             # deletion past the end
    kat(40, 20, ["ACGTACGTAC", "TTTTTTTTTT"], [[19, 20, 21, 0, 40, 5, 60, 20, 19, 93], [20] * 10], [0, 30],
        [[(7, 3), (8, 2), (1, 2), (0, 3)], [(0, 4), (3, 2), (0, 4), (9, 3), (6, 2)]])
    for seed in range(12):
        b = synth.fuzz_batch(seed, n_reads=60, ref_len=300, allow_overflow=(seed % 4 == 3))
        reads, quals, starts, ctuples = b.to_lists()
        kat(300, [0, 0, 20, 13][seed % 4], reads, quals, starts, ctuples)
    dump_gz("bcount_kats.json.gz", kats)


# ---------------------------------------------------------------- scheme goldens
BEDS = {
    "plain": "chr\t10\t30\tS_1_LEFT\t1\t+\nchr\t380\t400\tS_1_RIGHT\t1\t-\nchr\t300\t322\tS_2_LEFT\t2\t+\n"
             "chr\t690\t710\tS_2_RIGHT\t2\t-\nchr\t600\t625\tS_3_LEFT\t1\t+\nchr\t980\t1000\tS_3_RIGHT\t1\t-\n",
    "alts_unsorted": "c 600 625 P_3_LEFT 1 +\nc 980 1000 P_3_RIGHT 1 -\nc 10 30 P_1_LEFT 1 +\n"
                     "c 6 28 P_1_LEFT_alt1 1 +\nc 12 33 P_1_left_alt2 1 +\nc 380 400 P_1_RIGHT 1 -\n"
                     "c 377 405 P_1_RIGHT_alt7 1 -\nc 300 322 P_2_LEFT 2 +\nc 690 710 P_2_RIGHT 2 -\n"
                     "c 1200 1220 P_10_LEFT 2 +\nc 1500 1520 P_10_RIGHT 2 -\n",
    "missing_side": "c 10 30 Q_1_LEFT 1 +\nc 380 400 Q_1_RIGHT 1 -\nc 300 322 Q_2_LEFT 2 +\n"
                    "c 600 625 Q_3_LEFT 1 +\nc 980 1000 Q_3_RIGHT 1 -\n",
    "single": "c 10 30 R_1_LEFT 1 +\nc 380 400 R_1_RIGHT 1 -\n",
    "empty": "",
    "inverted": "c 10 30 T_1_LEFT 1 +\nc 80 100 T_1_RIGHT 1 -\nc 20 40 T_2_LEFT 1 +\nc 90 110 T_2_RIGHT 1 -\n"
                "c 30 50 T_3_LEFT 1 +\nc 100 120 T_3_RIGHT 1 -\n",
}


def scheme_goldens(tmpdir):
    out = {}
    for name, text in BEDS.items():
        p = os.path.join(tmpdir, name + ".bed")
        with open(p, "w") as fh:
            fh.write(text)
        out[name] = {"bed": text, "scheme": [[s, t, d] for s, t, d in ref_load_scheme(p)]}
    synth.artic_like_bed(os.path.join(tmpdir, "artic.bed"))
    out["artic_like"] = {"bed": open(os.path.join(tmpdir, "artic.bed")).read(),
                         "scheme": [[s, t, d] for s, t, d in ref_load_scheme(os.path.join(tmpdir, "artic.bed"))]}
    with open(os.path.join(HERE, "scheme.json"), "w") as fh:
        json.dump(out, fh)


# ---------------------------------------------------------------- CLI goldens
def cli_goldens(tmpdir):
    """Whole-path text goldens: reference run() over small synthetic alignments."""
    cases = {}
    # a 3 kb toy genome with 10 amplicons, shaped like config 1-3 (gaps of zero coverage at the ends)
    rec = synth.amplicon_sample(seed=21, n_reads=900, ref_len=3000, ref_name="toy")
    st = synth.amplicon_starts(3000)
    bed = os.path.join(tmpdir, "toy.bed")
    lines = []
    for i in range(0, st.size, 10):
        a = int(st[i])
        lines.append(f"toy\t{a}\t{a + 24}\ttoy_{i // 10 + 1}_LEFT\t1\t+")
        lines.append(f"toy\t{a + 376}\t{a + 400}\ttoy_{i // 10 + 1}_RIGHT\t1\t-")
    bed_text = "\n".join(lines) + "\n"
    with open(bed, "w") as fh:
        fh.write(bed_text)
    _REGISTRY["toy.bam"] = rec
    records_to_npz(os.path.join(HERE, "cli_toy.npz"), rec)
    argsets = {
        "tsv": [], "tsv_n": ["--show-n-bases"], "long": ["--long-format"], "long_n_dp5": ["--long-format", "--show-n-bases", "--decimal-places", "5"],
        "tsv_q20_m30": ["--min-base-quality", "20", "--min-mapping-quality", "30"],
        "tsv_q40_m60_chunk7": ["--min-base-quality", "40", "--min-mapping-quality", "60", "--chunk-size", "7"],
        "summarise": ["--summarise"], "summarise_q20": ["--summarise", "--min-base-quality", "20", "--decimal-places", "6"],
        "bed": ["--summarise-with-bed", bed], "bed_n_q20": ["--summarise-with-bed", bed, "--show-n-bases", "--min-base-quality", "20"],
        "tsv_dp0": ["--decimal-places", "0"],
    }
    for name, argv in argsets.items():
        text = run_cli(["toy.bam"] + argv)
        cases[name] = {"records": "cli_toy.npz", "argv": [a if a != bed else "@BED" for a in argv], "stdout": text}
    cases["_bed_text"] = bed_text
    # short reads, deep, config-3 shape, reduced
    rec3 = synth.deep_short_read_sample(seed=23, n_reads=4000, ref_len=3000, ref_name="toy")
    _REGISTRY["toy3.bam"] = rec3
    records_to_npz(os.path.join(HERE, "cli_toy3.npz"), rec3)
    for name, argv in {"deep_bed": ["--summarise-with-bed", bed], "deep_tsv": []}.items():
        cases[name] = {"records": "cli_toy3.npz", "argv": [a if a != bed else "@BED" for a in argv],
                       "stdout": run_cli(["toy3.bam"] + argv)}
    dump_gz("cli.json.gz", cases)

    # BaseCount API goldens (main.py:208-359) on the same toy input
    bc = refmain.BaseCount("toy.bam", min_base_quality=10)
    api = {"columns": bc.columns, "references": bc.references, "reference_lengths": bc.reference_lengths,
           "num_reads": bc.num_reads(), "num_reads_ref": bc.num_reads("toy"),
           "mean_coverage": float(bc.mean_coverage()), "mean_entropy": float(bc.mean_entropy()),
           "mean_entropy_min50": float(bc.mean_entropy(min_coverage=50)),
           "rows_head": list(bc.rows())[:40], "records_head": list(bc.records("toy"))[100:103]}
    with open(os.path.join(HERE, "api.json"), "w") as fh:
        json.dump(api, fh)


if __name__ == "__main__":
    import tempfile
    assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0 (reference order is set order)"
    stats_goldens()
    bcount_goldens()
    with tempfile.TemporaryDirectory() as td:
        scheme_goldens(td)
        cli_goldens(td)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
