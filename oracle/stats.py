"""oracle/stats.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (pure Python / numpy, float64) of the reference's per-position
statistics, summary, amplicon vectors, BED scheme and text emission:

    position_stats / rows          <- basecount/main.py:10-79   (get_entropy, get_stats)
    summary                        <- basecount/main.py:469-499
    amplicon_vectors               <- basecount/main.py:501-551
    format_*                       <- basecount/main.py:456-466, 488-499, 554-595
    scheme_windows                 <- basecount/scheme.py:3-78  (load_scheme, clip=True)

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
Parity status: PINNED against tests/golden/*.json, which tests/golden/make_golden.py
produced by importing the unmodified reference modules from /root/reference
(and against the README example rows, README.md:20-59).
"""
from __future__ import annotations

import math

import numpy as np

BASES6 = ("A", "C", "G", "T", "DS", "N")


def _norm_entropy(counts, total, scale):
    """scale * sum(-p log2 p) with p = c / total.  Uses the builtin sum() over a list that
    holds int 0 for empty classes, exactly as main.py:11 does: on CPython >= 3.12 that is a
    Neumaier-compensated float sum starting from int 0 (so never -0.0)."""
    probs = [c / total for c in counts]
    return scale * sum([-(p * math.log2(p)) if p != 0 else 0 for p in probs])


def position_stats(count_row, show_n_bases=False):
    """One reference position: (coverage, counts, percentages, entropy, secondary).

    Sentinels keep the reference's Python types: coverage 0 -> percentages are int -1,
    entropy and secondary are int 1 (main.py:34-36); secondary coverage 0 -> secondary
    stays int 1 (main.py:47).
    """
    c = [int(x) for x in count_row[:6]]
    if not show_n_bases:
        c = c[:5]                                    # drop N (main.py:30-31)
    k = len(c)
    cov = sum(c)                                     # main.py:37
    if cov == 0:
        return cov, c, [-1] * k, 1, 1
    pcs = [100 * (x / cov) for x in c]               # divide, then scale (main.py:40-41)
    ent = _norm_entropy(c, cov, 1 / math.log2(k))    # main.py:24,42
    top = int(np.argmax(c))                          # first maximum on ties (main.py:45)
    rest = c[:top] + c[top + 1:]
    rest_cov = sum(rest)
    sec = 1
    if rest_cov != 0:
        sec = _norm_entropy(rest, rest_cov, 1 / math.log2(k - 1))   # main.py:25,51
    return cov, c, pcs, ent, sec


def rows(counts, ref, show_n_bases=False, long_format=False):
    """Row lists exactly as get_stats builds them (main.py:57-78)."""
    names = BASES6 if show_n_bases else BASES6[:5]
    out = []
    for i, row in enumerate(counts):
        cov, c, pcs, ent, sec = position_stats(row, show_n_bases)
        if long_format:
            for b, x, pc in zip(names, c, pcs):
                out.append([ref, i + 1, cov, b, x, pc, ent, sec])
        else:
            out.append([ref, i + 1, cov, *c, *pcs, ent, sec])
    return out


def columns(show_n_bases=False, long_format=False):
    """Header names (main.py:232-265)."""
    if long_format:
        return ["reference", "position", "coverage", "base", "count", "percentage", "entropy", "secondary_entropy"]
    b = ["a", "c", "g", "t", "ds"] + (["n"] if show_n_bases else [])
    return ["reference", "position", "coverage"] + [f"num_{x}" for x in b] + [f"pc_{x}" for x in b] + \
           ["entropy", "secondary_entropy"]


def per_position_vectors(counts, show_n_bases=False):
    """coverage / entropy / secondary_entropy lists as run() collects them (main.py:474-477)."""
    cov, ent, sec = [], [], []
    for row in counts:
        a, _, _, e, s = position_stats(row, show_n_bases)
        cov.append(a)
        ent.append(e)
        sec.append(s)
    return cov, ent, sec


def summary(cov, ent, ref_len):
    """(pc_reference_coverage, avg_depth, avg_entropy), main.py:479-485."""
    avg_cov = np.mean(cov)
    avg_ent = np.mean(ent)
    pc = 100 * (len([x for x in cov if x != 0]) / ref_len)
    return pc, avg_cov, avg_ent


def amplicon_vectors(cov, ent, sec, windows):
    """Six vectors in print order (main.py:506-551).  windows = [(inside_start, inside_end)],
    both ends inclusive, 0-based indices (main.py:523)."""
    out = [[] for _ in range(6)]
    n = len(cov)
    for lo, hi in windows:
        a = max(lo, 0)
        b = min(hi, n - 1)
        for k, vec in enumerate((cov, ent, sec)):
            vals = vec[a:b + 1] if a <= b else []
            if len(vals):
                out[2 * k].append(np.mean(vals))
                out[2 * k + 1].append(np.median(vals))
            else:
                out[2 * k].append(-1)
                out[2 * k + 1].append(-1)
    return out


def _cell(x, dp):
    return x if isinstance(x, str) else str(round(x, dp))


def format_tsv(cols, all_rows, dp=3):
    """Per-position output text (main.py:456-466)."""
    lines = ["\t".join(cols)]
    for r in all_rows:
        lines.append("\t".join(_cell(x, dp) for x in r))
    return "\n".join(lines) + "\n"


def format_summary(ref, ref_len, num_reads, pc, avg_cov, avg_ent, dp=3):
    """Six `name<TAB>value` lines (main.py:488-499)."""
    items = [("reference_name", ref), ("reference_length", round(ref_len, dp)), ("num_reads", round(num_reads, dp)),
             ("pc_reference_coverage", round(pc, dp)), ("avg_depth", round(avg_cov, dp)),
             ("avg_entropy", round(avg_ent, dp))]
    return "".join(f"{k}\t{v}\n" for k, v in items)


AMPLICON_NAMES = ("mean_coverage_amplicon_vector", "median_coverage_amplicon_vector",
                  "mean_entropy_amplicon_vector", "median_entropy_amplicon_vector",
                  "mean_secondary_entropy_amplicon_vector", "median_secondary_entropy_amplicon_vector")


def format_amplicons(vectors, dp=3):
    """Six `name<TAB>v1, v2, ...` lines, "-" for an empty scheme (main.py:554-595)."""
    s = ""
    for name, vec in zip(AMPLICON_NAMES, vectors):
        s += name + "\t" + (", ".join(str(round(x, dp)) for x in vec) if vec else "-") + "\n"
    return s


def scheme_windows(bed_path):
    """Restatement of load_scheme(bed, clip=True) (scheme.py:3-78).

    Returns [(scheme, tile, {"start","inside_start","inside_end","end"})] sorted by
    int(tile), inner windows clipped to the neighbours' OUTER bounds (scheme.py:66-72).
    """
    with open(bed_path) as fh:
        recs = []
        for line in fh:
            f = line.strip().split()
            scheme, tile, side = f[3].split("_", 2)           # scheme.py:9
            recs.append((int(f[1]), int(f[2]), scheme, tile, side.upper()))
    tiles = {}
    order = []
    for s, e, scheme, tile, side in recs:
        t = tiles.get(tile)
        if t is None:
            t = tiles[tile] = {"start": -1, "inside_start": -1, "inside_end": -1, "end": -1}
        if "LEFT" in side:                                    # scheme.py:19-29
            if t["start"] == -1:
                t["start"], t["inside_start"] = s, e
            t["start"] = min(t["start"], s)
            t["inside_start"] = max(t["inside_start"], e)
        elif "RIGHT" in side:                                 # scheme.py:31-41
            if t["end"] == -1:
                t["end"], t["inside_end"] = e, s
            t["end"] = max(t["end"], e)
            t["inside_end"] = min(t["inside_end"], s)
    seen = set()
    for _, _, scheme, tile, _ in recs:                        # scheme.py:46-57 (first-seen order)
        t = tiles[tile]
        if t["inside_start"] != -1 and t["inside_end"] != -1 and tile not in seen:
            seen.add(tile)
            order.append((scheme, tile, t))
    order.sort(key=lambda x: int(x[1]))                       # stable, scheme.py:59
    out = []
    for i, (scheme, tile, t) in enumerate(order):             # scheme.py:60-74
        d = dict(t)
        if i > 0:
            d["inside_start"] = order[i - 1][2]["end"]
        if i < len(order) - 1:
            d["inside_end"] = order[i + 1][2]["start"]
        out.append((scheme, tile, d))
    return out
