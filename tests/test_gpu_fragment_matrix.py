"""GPU parity tests of the fragments-in-regions count matrix (reference
cistopic_class.py:868-889, :508-655) against oracle/fragment_oracle.py: bit-exact counts."""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

from oracle import fragment_oracle as fo

pytestmark = pytest.mark.gpu


def _random_case(seed, n_chrom=4, n_regions=400, n_fragments=20000, n_cells=60, overlapping=False, chrom_len=300000):
    rng = np.random.default_rng(seed)
    chroms = [f"chr{i + 1}" for i in range(n_chrom)]
    rc = rng.integers(0, n_chrom, n_regions)
    if overlapping:
        rs = rng.integers(0, chrom_len, n_regions)
        rl = rng.integers(1, 3000, n_regions)  # overlapping regions of very different lengths
    else:
        rs = rng.permutation(chrom_len // 600)[:n_regions] * 600
        rl = rng.integers(200, 600, n_regions)
    regions = pd.DataFrame({"Chromosome": [chroms[c] for c in rc], "Start": rs.astype(np.int64), "End": (rs + rl).astype(np.int64)})
    regions = regions.drop_duplicates().reset_index(drop=True)
    fc = rng.integers(0, n_chrom + 1, n_fragments)  # one extra chromosome without any region
    fs = rng.integers(0, chrom_len, n_fragments)
    fl = rng.integers(1, 700, n_fragments)
    names = [f"BC{j:03d}" for j in range(n_cells)]
    fragments = pd.DataFrame({"Chromosome": [(chroms + ["chrUn"])[c] for c in fc], "Start": fs, "End": fs + fl,
                              "Name": [names[j] for j in rng.integers(0, n_cells, n_fragments)],
                              "Score": rng.integers(1, 4, n_fragments)})
    return regions, fragments, names


def _assert_same(m, region_names, cell_names, ref: pd.DataFrame):
    assert sorted(region_names) == sorted(ref.index) and sorted(cell_names) == sorted(ref.columns)
    got = pd.DataFrame(m.toarray(), index=region_names, columns=cell_names).loc[ref.index, ref.columns]
    np.testing.assert_array_equal(got.to_numpy(), ref.to_numpy())


@pytest.mark.parametrize("overlapping", [False, True])
@pytest.mark.parametrize("seed", [1, 2])
def test_counts_match_oracle(seed, overlapping):
    from pycistopic_b200.fragment_matrix import count_fragments_in_regions
    regions, fragments, names = _random_case(seed, overlapping=overlapping)
    m, rn, cn = count_fragments_in_regions(fragments, regions)
    assert m.dtype == np.int32 and m.has_sorted_indices
    _assert_same(m, rn, cn, fo.fragment_matrix(regions, fragments))
    # region order = order of `regions`, barcodes sorted
    all_names = (regions.Chromosome + ":" + regions.Start.astype(str) + "-" + regions.End.astype(str)).tolist()
    assert rn == [n for n in all_names if n in set(rn)] and cn == sorted(cn)
    # a barcode whitelist in a given order (cistopic_class.py:855-857)
    valid = names[10:40][::-1]
    m2, rn2, cn2 = count_fragments_in_regions(fragments, regions, valid_bc=valid)
    _assert_same(m2, rn2, cn2, fo.fragment_matrix(regions, fragments, valid_bc=valid))
    assert cn2 == [c for c in valid if c in set(cn2)]


def test_many_cells_window_and_touching_intervals():
    """More cells than one shared-memory window (57 k counters), and the half-open boundary cases: a fragment that
    ends where a region starts (or starts where it ends) does not overlap."""
    from pycistopic_b200.fragment_matrix import count_fragments_in_regions
    regions = pd.DataFrame({"Chromosome": ["chr1", "chr1", "chr2"], "Start": [100, 500, 100], "End": [200, 900, 200]})
    n_cells = 130000
    rng = np.random.default_rng(0)
    cells = rng.integers(0, n_cells, 40000)
    frag = pd.DataFrame({"Chromosome": "chr1", "Start": 150, "End": 600, "Name": [f"c{j:06d}" for j in cells]})
    edge = pd.DataFrame({"Chromosome": ["chr1", "chr1", "chr1", "chr2", "chr2"], "Start": [50, 200, 199, 200, 0],
                         "End": [100, 300, 201, 300, 101], "Name": ["edgeA", "edgeB", "edgeC", "edgeD", "edgeE"]})
    fragments = pd.concat([frag, edge], ignore_index=True)
    m, rn, cn = count_fragments_in_regions(fragments, regions)
    ref = fo.fragment_matrix(regions, fragments)
    _assert_same(m, rn, cn, ref)
    assert "edgeA" not in cn and "edgeB" not in cn and "edgeD" not in cn and "edgeC" in cn and "edgeE" in cn
    assert m.sum() == 2 * len(cells) + 2


def test_object_from_fragments_feeds_the_sampler():
    """create_cistopic_object_from_fragments -> run_cgs_models: the binary matrix built from fragments is what the
    path consumes (cistopic_class.py:592, lda_models.py:151)."""
    from pycistopic_b200 import run_cgs_models
    from pycistopic_b200.fragment_matrix import create_cistopic_object_from_fragments
    regions, fragments, names = _random_case(5, n_regions=300, n_fragments=60000, n_cells=80)
    obj = create_cistopic_object_from_fragments(fragments, regions, is_acc=1, min_frag=1, project="p")
    ref = fo.fragment_matrix(regions, fragments)
    assert obj.cell_names == [c + "___p" for c in sorted(ref.columns)]
    got = pd.DataFrame(obj.fragment_matrix.toarray(), index=obj.region_names, columns=[c[:-4] for c in obj.cell_names])
    np.testing.assert_array_equal(got.loc[ref.index, ref.columns].to_numpy(), ref.to_numpy())
    assert obj.binary_matrix.max() == 1 and obj.binary_matrix.nnz == (ref.to_numpy() > 0).sum()
    np.testing.assert_array_equal(obj.cell_data["cisTopic_nr_frag"].to_numpy().astype(np.int64),
                                  got.to_numpy().sum(axis=0))
    np.testing.assert_array_equal(obj.region_data["cisTopic_nr_acc"].to_numpy(), (got.to_numpy() > 0).sum(axis=1))
    assert list(obj.region_data.columns[:4]) == ["Chromosome", "Start", "End", "Width"]
    # is_acc = 2: only pairs with at least two fragments are accessible
    obj2 = create_cistopic_object_from_fragments(fragments, regions, is_acc=2, project="p")
    assert obj2.binary_matrix.nnz == (ref.to_numpy() >= 2).sum()
    models = run_cgs_models(obj, n_topics=[3], n_iter=20, random_state=1)
    obj.add_LDA_model(models[0])
    assert obj.selected_model.cell_topic.shape == (3, len(obj.cell_names))
    assert list(obj.selected_model.topic_region.index) == obj.region_names
