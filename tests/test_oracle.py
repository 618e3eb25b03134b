"""The oracle (oracle/) pinned against vectors produced by the unmodified reference
(tests/golden/make_golden.py) and, when oracle/_ref is present, against the compiled
reference live.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, typed_equal
from basecount_b200 import synth
from basecount_b200.records import ReadBatch, select_reads
from oracle import bcount as obc
from oracle import stats as ost


def _run_oracle(case):
    b = ReadBatch.from_lists(case["reads"], case["qualities"], case["starts"],
                             [[tuple(t) for t in c] for c in case["ctuples"]])
    return obc.bcount_flat(case["ref_len"], case["min_base_quality"], b)


def test_bcount_oracle_matches_reference_kats():
    kats = load_golden("bcount_kats.json.gz")
    assert len(kats) >= 20
    n_err = 0
    for case in kats:
        if "error" in case:
            n_err += 1
            with pytest.raises(IndexError):
                _run_oracle(case)
        else:
            got = _run_oracle(case)
            want = np.asarray(case["counts"], dtype=np.uint32).reshape(case["ref_len"], 6)
            assert np.array_equal(got, want)
    assert n_err >= 3


def test_bcount_oracle_matches_compiled_reference_live():
    ref = obc.load_ref_bcount()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for seed in range(100, 110):
        b = synth.fuzz_batch(seed, n_reads=150, ref_len=400)
        for mbq in (0, 20, 40):
            want = np.asarray(ref(400, mbq, *b.to_lists()), dtype=np.uint32)
            assert np.array_equal(obc.bcount_flat(400, mbq, b), want)
    rec = synth.amplicon_sample(seed=5, n_reads=3000, ref_len=5000)
    b = select_reads(rec, 0, 30)
    want = np.asarray(ref(5000, 20, *b.to_lists()), dtype=np.uint32)
    assert np.array_equal(obc.bcount_flat(5000, 20, b), want)


def test_aligned_bases_definition():
    b = synth.fuzz_batch(3, n_reads=100)
    assert obc.aligned_bases(b.cigar) == b.aligned_bases()
    op, ln = b.cigar & 0xF, b.cigar >> 4
    assert b.aligned_bases() == int(sum(int(l) for o, l in zip(op, ln) if o in (0, 2, 3, 7, 8)))


def test_stats_oracle_matches_reference_rows():
    g = load_golden("stats.json.gz")
    for case in g["cases"]:
        got = ost.rows(g["counts"], "REF", case["show_n_bases"], case["long_format"])
        assert len(got) == len(case["rows"])
        for a, b in zip(got, case["rows"]):
            assert typed_equal(a, b), (a, b)


def test_stats_oracle_readme_text():
    g = load_golden("stats.json.gz")
    for key, text in g["readme_text"].items():
        c = [int(x) for x in key.split(",")] + [0]
        row = ost.rows([c], "X")[0]
        assert "\t".join(str(round(x, 3)) for x in row[2:]) == text


def test_scheme_oracle_matches_reference(tmp_path):
    g = load_golden("scheme.json")
    for name, case in g.items():
        p = tmp_path / (name + ".bed")
        p.write_text(case["bed"])
        got = [[s, t, d] for s, t, d in ost.scheme_windows(str(p))]
        assert got == case["scheme"], name


def test_cli_text_oracle_matches_reference(tmp_path):
    """End-to-end restatement (oracle counts -> oracle stats -> oracle formatting) against the
    text the reference CLI printed for the same alignments."""
    from conftest import load_records
    g = load_golden("cli.json.gz")
    bed = tmp_path / "toy.bed"
    bed.write_text(g["_bed_text"])
    for name, case in g.items():
        if name.startswith("_"):
            continue
        rec = load_records(case["records"])
        argv = case["argv"]
        def opt(flag, default):
            return int(argv[argv.index(flag) + 1]) if flag in argv else default
        mbq, mmq, dp = opt("--min-base-quality", 0), opt("--min-mapping-quality", 0), opt("--decimal-places", 3)
        show_n, long_format = "--show-n-bases" in argv, "--long-format" in argv
        b = select_reads(rec, 0, mmq)
        counts = obc.bcount_flat(rec.ref_lengths[0], mbq, b).astype(np.int64).tolist()
        if "--summarise" in argv or "--summarise-with-bed" in argv:
            cov, ent, sec = ost.per_position_vectors(counts, show_n)
            text = ost.format_summary("toy", rec.ref_lengths[0], b.n, *ost.summary(cov, ent, rec.ref_lengths[0]), dp=dp)
            if "--summarise-with-bed" in argv:
                win = [(d["inside_start"], d["inside_end"]) for _, _, d in ost.scheme_windows(str(bed))]
                text += ost.format_amplicons(ost.amplicon_vectors(cov, ent, sec, win), dp=dp)
        else:
            text = ost.format_tsv(ost.columns(show_n, long_format), ost.rows(counts, "toy", show_n, long_format), dp)
        assert text == case["stdout"], name
