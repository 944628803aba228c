#!/usr/bin/env python
"""Pin the interval half of the count-matrix oracle against the REAL pyranges (TEST INFRASTRUCTURE; run
it on any machine where pyranges can be installed -- it is not in this image):

    pip install pyranges
    python oracle/pin_against_pyranges.py   # exit 0 = pinned, 3 = pyranges absent (the join stays UNPINNED), 1 = mismatch

What is compared, on seeded random fragments / regions (overlapping regions of very different lengths, touching
intervals, chromosomes present on one side only):

* ``regions.join(fragments)`` (cistopic_class.py:868)     == ``fragment_oracle.join_regions_fragments`` as a multiset of
  (regionID, Name) pairs, and the count matrix the reference builds from it (:880-889) == ``fragment_oracle.fragment_matrix``
* ``regions.overlap(blacklist, invert=True)`` (:575-579)   == ``fragment_matrix.regions_outside_blacklist`` (the set of
  regions kept; pyranges returns them grouped by chromosome, the mirror in input order)
* ``pyranges.read_bed`` (:835, :578)                       == ``fragment_matrix.read_bed`` (Chromosome, Start, End)

On success it writes ``tests/golden/pyranges_pin.json`` -- commit it and drop the "parity unpinned" note of
oracle/fragment_oracle.py and DESIGN.md section 3.9.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EXIT_PINNED, EXIT_MISMATCH, EXIT_ABSENT = 0, 1, 3


def cases():
    for seed, overlapping in ((1, False), (2, True), (3, True)):
        rng = np.random.default_rng(seed)
        chroms = ["chr1", "chr2", "chr3", "chrX"]
        n_regions, n_fragments, length = 300, 20000, 200000
        rs = rng.integers(0, length, n_regions) if overlapping else rng.permutation(length // 600)[:n_regions] * 600
        rl = rng.integers(1, 3000, n_regions) if overlapping else rng.integers(200, 600, n_regions)
        regions = pd.DataFrame({"Chromosome": rng.choice(chroms[:3], n_regions), "Start": rs, "End": rs + rl}).drop_duplicates()
        fs = rng.integers(0, length, n_fragments)
        fragments = pd.DataFrame({"Chromosome": rng.choice(chroms, n_fragments), "Start": fs, "End": fs + rng.integers(1, 700, n_fragments),
                                  "Name": [f"BC{j:03d}" for j in rng.integers(0, 60, n_fragments)], "Score": 1})
        # touching intervals: a fragment that ends where a region starts, and one that starts where it ends
        r0 = regions.iloc[0]
        edge = pd.DataFrame({"Chromosome": [r0.Chromosome] * 2, "Start": [max(0, r0.Start - 10), r0.End], "End": [r0.Start, r0.End + 10],
                             "Name": ["EDGE1", "EDGE2"], "Score": 1})
        yield seed, regions.reset_index(drop=True), pd.concat([fragments, edge], ignore_index=True)


def run(verbose=True):
    try:
        import pyranges as pr
    except Exception as e:
        if verbose:
            print(f"pyranges not importable ({e!r}): the interval join of the count-matrix oracle stays PARITY UNPINNED")
        return EXIT_ABSENT, {}
    from oracle import fragment_oracle as fo
    from pycistopic_b200 import fragment_matrix as fm
    report, ok = {"pyranges": getattr(pr, "__version__", "?"), "cases": {}}, True
    for seed, regions, fragments in cases():
        region_id = regions.Chromosome.astype(str) + ":" + regions.Start.astype(str) + "-" + regions.End.astype(str)
        pr_regions = pr.PyRanges(regions.assign(regionID=region_id))
        joined = pr_regions.join(pr.PyRanges(fragments)).df
        mine = fo.join_regions_fragments(regions, fragments)
        same_pairs = sorted(zip(joined.regionID.astype(str), joined.Name.astype(str))) == sorted(zip(mine.regionID.astype(str), mine.Name.astype(str)))
        counts_df = pd.concat([joined.regionID.astype("category"), joined.Name.astype("category")], axis=1, sort=False)
        ref_matrix = (counts_df.groupby(["Name", "regionID"], sort=False, observed=True).size().unstack(level="Name", fill_value=0)
                      .astype(np.int32))
        want = fo.fragment_matrix(regions, fragments)
        same_matrix = (sorted(ref_matrix.index) == sorted(want.index) and sorted(ref_matrix.columns) == sorted(want.columns)
                       and np.array_equal(ref_matrix.loc[want.index, want.columns].to_numpy(), want.to_numpy()))
        rng = np.random.default_rng(100 + seed)
        bs = rng.integers(0, 200000, 25)
        black = pd.DataFrame({"Chromosome": rng.choice(["chr1", "chr2", "chrM"], 25), "Start": bs, "End": bs + rng.integers(1, 5000, 25)})
        kept = pr_regions.overlap(pr.PyRanges(black), invert=True).df
        kept_ids = set(kept.Chromosome.astype(str) + ":" + kept.Start.astype(str) + "-" + kept.End.astype(str))
        same_blacklist = kept_ids == set(region_id[fm.regions_outside_blacklist(regions, black)])
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "regions.bed")
            regions.to_csv(path, sep="\t", header=False, index=False)
            a, b = pr.read_bed(path).df, fm.read_bed(path)
            key = ["Chromosome", "Start", "End"]
            same_bed = (a[key].astype({"Chromosome": str}).sort_values(key).reset_index(drop=True)
                        .equals(b[key].astype({"Chromosome": str, "Start": a.Start.dtype, "End": a.End.dtype}).sort_values(key).reset_index(drop=True)))
        report["cases"][str(seed)] = {"join_pairs": same_pairs, "count_matrix": same_matrix, "blacklist": same_blacklist, "read_bed": same_bed,
                                      "pairs": int(len(joined))}
        ok &= same_pairs and same_matrix and same_blacklist and same_bed
        if verbose:
            print(seed, report["cases"][str(seed)])
    if ok:
        with open(os.path.join(ROOT, "tests", "golden", "pyranges_pin.json"), "w") as f:
            json.dump(report, f, indent=1)
    return (EXIT_PINNED if ok else EXIT_MISMATCH), report


if __name__ == "__main__":
    sys.exit(run()[0])
