"""CPU oracle of the fragments-in-regions count matrix.  TEST INFRASTRUCTURE ONLY.

Restates cistopic_class.py:868-889 of the reference: ``regions.join(fragments)`` -- pyranges' overlap join on
half-open intervals, one output row per (region, fragment) pair with region.Start < fragment.End and
fragment.Start < region.End on the same chromosome -- followed, verbatim, by the reference's pandas counting
(``groupby(["Name", "regionID"], sort=False, observed=True).size().unstack(level="Name", fill_value=0)``).
pyranges is not installed here, so the join itself is a plain NumPy all-pairs test per chromosome
(**parity unpinned** for the join -- oracle/pin_against_pyranges.py pins it wherever pyranges installs; the pandas
half is the reference's own expression).
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def join_regions_fragments(regions: pd.DataFrame, fragments: pd.DataFrame) -> pd.DataFrame:
    """All (region, fragment) overlaps: columns regionID, Name (and the coordinates)."""
    out = []
    region_id = (regions["Chromosome"].astype(str) + ":" + regions["Start"].astype(str) + "-" + regions["End"].astype(str))
    for chrom, reg in regions.assign(regionID=region_id).groupby("Chromosome", sort=False):
        frag = fragments[fragments["Chromosome"] == chrom]
        if len(frag) == 0:
            continue
        rs, re = reg["Start"].to_numpy()[:, None], reg["End"].to_numpy()[:, None]
        fs, fe = frag["Start"].to_numpy()[None, :], frag["End"].to_numpy()[None, :]
        ri, fi = np.nonzero((rs < fe) & (fs < re))
        out.append(pd.DataFrame({"regionID": reg["regionID"].to_numpy()[ri], "Name": frag["Name"].to_numpy()[fi]}))
    if not out:
        return pd.DataFrame({"regionID": [], "Name": []})
    return pd.concat(out, ignore_index=True)


def fragment_matrix(regions: pd.DataFrame, fragments: pd.DataFrame, valid_bc=None) -> pd.DataFrame:
    """The dense regions x barcodes count DataFrame of cistopic_class.py:880-889."""
    if valid_bc is not None:
        fragments = fragments[fragments.Name.isin(set(valid_bc))]
    fragments_in_regions = join_regions_fragments(regions, fragments)
    counts_df = pd.concat(
        [fragments_in_regions.regionID.astype("category"), fragments_in_regions.Name.astype("category")], axis=1, sort=False)
    m = (counts_df.groupby(["Name", "regionID"], sort=False, observed=True).size()
         .unstack(level="Name", fill_value=0).astype(np.int32))
    m.columns.names = [None]
    return m
