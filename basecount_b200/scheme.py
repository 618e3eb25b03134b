"""Primer-scheme BED -> per-amplicon windows, for --summarise-with-bed.

Same observable behaviour as the reference's load_scheme(bed, clip=True)
(basecount/scheme.py:3-78): whitespace-split lines, column 4 is SCHEME_TILE_SIDE[...],
LEFT / RIGHT matched case-insensitively as substrings so `_alt` primers widen the outer
bounds and tighten the inner ones, tiles need both sides, sorted by int(tile), and the
inner window of each tile is clipped to the neighbours' OUTER bounds.  Column 1 (chrom)
is ignored, exactly like the reference; a blank or short line raises, as it does there.

The windows feed K3 (bc_amplicons) as [inside_start, inside_end] inclusive 0-based ranges.
"""
from __future__ import annotations


class _Tile:
    __slots__ = ("scheme", "name", "start", "inside_start", "inside_end", "end")

    def __init__(self, scheme, name):
        self.scheme, self.name = scheme, name
        self.start = self.inside_start = self.inside_end = self.end = -1

    def add_left(self, s, e):
        if self.start == -1:
            self.start, self.inside_start = s, e
        self.start = min(self.start, s)                  # leftmost LEFT start
        self.inside_start = max(self.inside_start, e)    # rightmost LEFT end

    def add_right(self, s, e):
        if self.end == -1:
            self.end, self.inside_end = e, s
        self.end = max(self.end, e)                      # rightmost RIGHT end
        self.inside_end = min(self.inside_end, s)        # leftmost RIGHT start

    def complete(self):
        return self.inside_start != -1 and self.inside_end != -1

    def as_dict(self):
        return {"start": self.start, "inside_start": self.inside_start, "inside_end": self.inside_end, "end": self.end}


def load_scheme(bed, clip=True):
    tiles = {}
    first_seen = []
    with open(bed) as fh:
        for line in fh:
            f = line.strip().split()
            start, end = int(f[1]), int(f[2])
            scheme, name, side = f[3].split("_", 2)
            t = tiles.get(name)
            if t is None:
                t = tiles[name] = _Tile(scheme, name)
                first_seen.append((scheme, name))
            side = side.upper()
            if "LEFT" in side:
                t.add_left(start, end)
            elif "RIGHT" in side:
                t.add_right(start, end)
    # the reference reports the scheme name of the FIRST line of each tile
    kept = [(scheme, name, tiles[name]) for scheme, name in first_seen if tiles[name].complete()]
    kept.sort(key=lambda x: int(x[1]))
    out = []
    for i, (scheme, name, t) in enumerate(kept):
        d = t.as_dict()
        if clip:
            if i > 0:
                d["inside_start"] = kept[i - 1][2].end
            if i < len(kept) - 1:
                d["inside_end"] = kept[i + 1][2].start
        out.append((scheme, name, d))
    return out


def windows(scheme):
    """(inside_start[], inside_end[]) as run() extracts them (basecount/main.py:503-504)."""
    return [t[2]["inside_start"] for t in scheme], [t[2]["inside_end"] for t in scheme]
