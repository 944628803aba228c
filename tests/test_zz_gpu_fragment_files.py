"""GPU test of the whole front end: fragments file (BGZF) + regions BED + blacklist BED ->
``create_cistopic_object_from_fragments`` (cistopic_class.py:738-935) -> ``run_cgs_models``.  The reader, the host
logic and the device counting each have their own parity tests (test_cpu_fragment_files.py,
test_gpu_fragment_matrix.py); this one runs them together.  (Named to run after the parity suites.)"""
import numpy as np
import pandas as pd
import pytest

from oracle import fragment_oracle as fo
from test_cpu_fragment_files import _fragments, _write

pytestmark = pytest.mark.gpu


def test_fragment_files_to_models(tmp_path):
    from pycistopic_b200 import run_cgs_models
    from pycistopic_b200.fragment_matrix import create_cistopic_object_from_fragments
    frags = _fragments(21, n=40000, n_cells=50, score=False, duplicates=True)
    rng = np.random.default_rng(21)
    starts = np.sort(rng.choice(400, 250, replace=False)) * 500
    regions = pd.DataFrame({"Chromosome": rng.choice(["chr1", "chr2", "chr3", "chr4", "chrX"], 250), "Start": starts,
                            "End": starts + rng.integers(200, 700, 250)})
    black = pd.DataFrame({"Chromosome": ["chr1"], "Start": [0], "End": [30000]})
    f_path = _write(frags, str(tmp_path / "fragments.tsv.gz"), header="# id=s\n", compress="bgzip")
    r_path = _write(regions, str(tmp_path / "regions.bed"))
    b_path = _write(black, str(tmp_path / "blacklist.bed"))
    obj = create_cistopic_object_from_fragments(f_path, r_path, path_to_blacklist=b_path, project="s")
    want = fo.fragment_matrix(regions, frags.drop_duplicates())
    keep = [r for r in want.index if not (r.startswith("chr1:") and int(r.split(":")[1].split("-")[0]) < 30000)]
    want = want.loc[keep]
    bare = [c.split("___")[0] for c in obj.cell_names]
    got = pd.DataFrame(obj.fragment_matrix.toarray(), index=obj.region_names, columns=bare)
    assert sorted(obj.region_names) == sorted(want.index) and set(bare) <= set(want.columns)
    np.testing.assert_array_equal(got.loc[want.index].to_numpy(), want.loc[:, bare].to_numpy())
    assert obj.binary_matrix.max() == 1 and "Unique_nr_frag" in obj.cell_data.columns
    models = run_cgs_models(obj, n_topics=[4], n_iter=20, random_state=3)
    obj.add_LDA_model(models[0])
    assert obj.selected_model.cell_topic.shape == (4, len(obj.cell_names))
    assert int(models[0].topic_ass["Assignments"].sum()) == obj.binary_matrix.nnz


def test_c_host_example_on_the_device(tmp_path):
    """examples/host_example.c: a C host drives corpus -> model -> sweeps -> exports through the ABI and checks the
    bookkeeping invariants itself; its initial log-likelihood is the oracle's for the same matrix (z = i mod K)."""
    import re
    import subprocess

    from test_cpu_oracle_and_host import build_c_host_example
    exe = build_c_host_example(tmp_path)
    if exe is None:
        pytest.skip("no gcc")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and res.stdout.rstrip().endswith("OK"), res.stdout + res.stderr
    m = re.search(r"(\d+) tokens, K = 4: log-likelihood (-?[\d.]+) -> (-?[\d.]+) after 150 sweeps", res.stdout)
    assert m and int(m.group(1)) == 22432
    # oracle/lda_oracle.OracleLDA on the matrix the program generates (xorshift64, seed in the source): -213918.78 at
    # z = i mod K; after 150 sweeps eight seeds of the sequential chain end between -171 758 and -171 983
    assert float(m.group(2)) == pytest.approx(-213918.78, abs=0.2)
    assert float(m.group(3)) == pytest.approx(-171870.0, rel=0.01)  # the 1 % of north_star
