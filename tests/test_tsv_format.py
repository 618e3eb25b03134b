"""The native TSV emitter (csrc/tsv_format.h) against the reference's own formatting expression
`str(round(x, decimal_places))` (basecount/main.py:461) on adversarial values.  CPU only."""
import numpy as np
import pytest

from basecount_b200.main import build_rows, format_rows_text


def _python_text(ref, counts, st, show_n, long_format, dp):
    rows = build_rows(ref, counts, st, show_n, long_format)
    return "\n".join("\t".join(x if isinstance(x, str) else str(round(x, dp)) for x in row) for row in rows)


def _case(seed, L, show_n):
    rng = np.random.default_rng(seed)
    k = 6 if show_n else 5
    counts = rng.integers(0, 3000, size=(L, 6)).astype(np.int64)
    counts[rng.random(L) < 0.1] = 0
    cov = counts[:, :k].sum(axis=1)
    pc = np.zeros((k, L))
    nz = cov > 0
    pc[:, nz] = 100 * (counts[nz, :k].T / cov[nz])
    ent = rng.random(L)
    sec = rng.random(L)
    # ties and near-ties of the 3rd / 4th decimal, values that round up across a digit boundary,
    # tiny values, exact binary fractions, the extremes the columns can hold
    special = np.array([0.0005, 0.0015, 0.0025, 0.00049999999999999994, 0.0005000000000000001, 2.6745, 2.675, 1.0005,
                        0.125, 0.375, 0.0625, 99.9995, 99.99949999999999, 100.0, 0.0, 1.0, 1e-300, 4.9e-324, 0.00005,
                        0.9999999999999999, 33.333333333333336, 66.66666666666667, 0.1 + 0.2, 1 / 3, 2 / 3, 0.30000000000000004,
                        12.3456789012345, 0.5, 1.5, 2.5, 1e-4, 1.00000000000001e-4, 9.99e-5, 50.0, 25.00005])
    m = min(special.size, L)
    ent[:m] = special[:m] % 1.0000001
    sec[L - m:] = special[:m] % 1.0000001
    for j in range(k):
        pc[j, (np.arange(m) * 3 + j) % L] = special[:m]
    flags = np.where(cov == 0, 3, 0).astype(np.uint8)
    flags[(rng.random(L) < 0.05) & (cov > 0)] = 2
    pc[:, cov == 0] = -1.0
    ent[cov == 0] = 1.0
    sec[flags != 0] = 1.0
    return counts, {"coverage": cov, "pc": pc, "entropy": ent, "secondary": sec, "flags": flags}


@pytest.mark.parametrize("show_n", [False, True])
@pytest.mark.parametrize("long_format", [False, True])
def test_native_rows_text_equals_python_formatting(show_n, long_format):
    counts, st = _case(3, 700, show_n)
    for dp in (0, 1, 2, 3, 4):
        got = format_rows_text("MN908947.3", counts, st, show_n, long_format, dp, threads=3)
        assert got is not None
        assert got == _python_text("MN908947.3", counts, st, show_n, long_format, dp), dp


def test_native_emitter_refuses_what_it_cannot_do_exactly():
    counts, st = _case(5, 50, False)
    assert format_rows_text("x", counts, st, False, False, 5) is None          # repr switches to 1e-05 notation
    assert format_rows_text("x", counts, st, False, False, -1) is None
    st["entropy"][7] = float("nan")
    assert format_rows_text("x", counts, st, False, False, 3) is None
    st["entropy"][7] = 0.5
    st["pc"][2, 9] = 1e300
    assert format_rows_text("x", counts, st, False, False, 3) is None


def test_many_random_doubles():
    """2 M random doubles (uniform, and scaled near ties) through both formatters."""
    rng = np.random.default_rng(99)
    L = 400_000
    counts = np.ones((L, 6), dtype=np.int64)
    base = rng.integers(0, 100_000, size=(5, L)) / 1000.0
    jitter = rng.choice([0.0, 0.0005, 0.00049999999, 0.00050000001, 1e-17, -1e-17], size=(5, L))
    st = {"coverage": np.full(L, 5, dtype=np.int64), "pc": np.abs(base + jitter), "entropy": rng.random(L),
          "secondary": rng.random(L) * 1e-3, "flags": np.zeros(L, dtype=np.uint8)}
    got = format_rows_text("r", counts, st, False, False, 3)
    assert got == _python_text("r", counts, st, False, False, 3)
