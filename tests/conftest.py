import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    p = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        with gzip.open(p, "rt") as fh:
            return json.load(fh)
    with open(p) as fh:
        return json.load(fh)


def load_records(name):
    from basecount_b200.records import Records
    z = np.load(os.path.join(GOLDEN, name))
    return Records([str(x) for x in z["ref_names"]], [int(x) for x in z["ref_lengths"]], z["ref_id"], z["pos"],
                   z["mapq"], z["flag"], z["cigar"], z["cigar_off"], z["seq"], z["qual"], z["seq_off"])


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def typed_equal(a, b):
    """Equality that also distinguishes int from float (the reference's sentinels are ints)."""
    if isinstance(a, (list, tuple)):
        return isinstance(b, (list, tuple)) and len(a) == len(b) and all(typed_equal(x, y) for x, y in zip(a, b))
    if isinstance(a, bool) or isinstance(b, bool):
        return a == b
    if isinstance(a, (int, np.integer)):
        return isinstance(b, (int, np.integer)) and int(a) == int(b)
    if isinstance(a, (float, np.floating)):
        return isinstance(b, (float, np.floating)) and float(a) == float(b)
    return a == b
