"""CPU suite: the NMF / PPC oracle (oracle/nmf_oracle.c) against the reference's own golden vectors
(T/testdata/PPCTestData.java, NMFTestData.java, ClusteringTestData.java via tests/golden/make_golden_nmf.py).
The reference's bar is accuracy = 1e-4 (T/util/HadoopIntegrationTest.java:53); the restatement lands at
1e-10 or better, which also pins the Jacobi structure (W is updated from the OLD H) and the fact that the
PPC row normalisation is a no-op in the reference."""
import json
import os

import numpy as np
import pytest

from oracle import nmf_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def ppc():
    return _load("ppc_test_data.json")


@pytest.fixture(scope="module")
def nmf():
    return _load("nmf_test_data.json")


@pytest.mark.parametrize("which,mode", [("ppc", orc.PPC), ("nmf", orc.NMF)])
@pytest.mark.parametrize("combine_len,split_rows", [(0, 0), (8, 16)])
def test_one_and_ten_iterations_match_the_reference_goldens(which, mode, combine_len, split_rows, ppc, nmf):
    g = ppc if which == "ppc" else nmf
    u, i, s = orc.coo_from_dense(g["A"])
    # T/nmf/ppc/hcomputation/TestHDFSPPCHComputation.java:66, T/nmf/hcomputation/TestHDFSHComputation.java:66,
    # T/nmf/wcomputation/TestHDFSWComputation.java:66
    H1, W1 = orc.run(mode, u, i, s, g["H_init"], g["W_init"], 1, combine_len=combine_len, split_rows=split_rows)
    assert np.max(np.abs(H1 - np.array(g["H_one"]))) < 1e-10
    assert np.max(np.abs(W1 - np.array(g["W_one"]))) < 1e-10
    # T/nmf/ppc/PPCHDFSDriverTest.java:62-63, T/nmf/NMFHDFSDriverTest.java:62-63
    H10, W10 = orc.run(mode, u, i, s, g["H_init"], g["W_init"], 10, combine_len=combine_len, split_rows=split_rows)
    assert np.max(np.abs(H10 - np.array(g["H_ten"]))) < 1e-10
    assert np.max(np.abs(W10 - np.array(g["W_ten"]))) < 1e-10


def test_the_reference_never_normalises_ppc_rows(ppc):
    """PPCHComputationReducer.java:88-90 drops the vector `normalize(1)` returns: the golden H_ten rows do not
    sum to 1, and applying the normalisation moves H_ten beyond the reference's own 1e-4 bar."""
    g = ppc
    assert np.max(np.abs(np.array(g["H_ten"]).sum(1) - 1.0)) > 1e-3
    u, i, s = orc.coo_from_dense(g["A"])
    Hn, _ = orc.run(orc.PPC, u, i, s, g["H_init"], g["W_init"], 10, apply_normalization=True, normalization_frequency=-1)
    assert np.allclose(Hn.sum(1), 1.0, atol=1e-12)
    assert np.max(np.abs(Hn - np.array(g["H_ten"]))) > 1e-4


def test_ppc_hand_example(ppc):
    """the 5 x 7 example of PPCTestData.java:29-57 (unused by the reference's tests): h1p is one PPC H step
    from (h0p, w0p).  Its ratings are not float-representable, hence 1e-7 (the ratings file holds FloatWritable)."""
    g = ppc
    u, i, s = orc.coo_from_dense(g["Ap"])
    H1, _ = orc.run(orc.PPC, u, i, s, g["h0p"], g["w0p"], 1)
    assert np.max(np.abs(H1 - np.array(g["h1p"]))) < 1e-7


def test_cluster_assignment_and_counts():
    g = _load("clustering_test_data.json")       # T/nmf/clustering/TestClusterAssignment.java
    cl, cnt = orc.cluster_assign(g["H"])
    assert cl.tolist() == g["clustering"]
    assert cnt.tolist() == g["clusteringCount"]


def test_max_value_index_rules():
    H = np.array([[0.2, 0.7, 0.7], [0.0, 0.0, 0.0], [-1.0, 0.0, -2.0], [-1.0, -0.5, -2.0]])
    cl, cnt = orc.cluster_assign(H)
    assert cl.tolist() == [1, 0, 1, 1]           # first max; all-zero -> first zero; negative max with a zero -> the zero
    assert cnt.tolist() == [1, 3, 0]


def test_missing_rows_fail_like_the_reference(ppc):
    g = ppc
    u, i, s = orc.coo_from_dense(g["A"])
    keep = u != 7                                  # user 7 has no rating: HComputationReducer.java:52-55
    with pytest.raises(orc.OracleError) as e:
        orc.run(orc.PPC, u[keep], i[keep], s[keep], g["H_init"], g["W_init"], 1)
    assert e.value.code == -2 and e.value.bad_id == 7
    keep = i != 42                                 # item 42 unrated: WComputationMapper.java:95-98
    with pytest.raises(orc.OracleError) as e:
        orc.run(orc.NMF, u[keep], i[keep], s[keep], g["H_init"], g["W_init"], 1)
    assert e.value.code == -10 and e.value.bad_id == 42


def test_summation_structure_only_moves_the_last_bits(ppc):
    """combiner group sizes change the summation tree, never the result beyond rounding (the reference's own order is
    shuffle-defined); relabelling users and items permutes the factor rows accordingly."""
    g = ppc
    u, i, s = orc.coo_from_dense(g["A"])
    base = orc.run(orc.PPC, u, i, s, g["H_init"], g["W_init"], 5)
    for cl, sr in ((1, 1), (3, 7), (1024, 256)):
        H, W = orc.run(orc.PPC, u, i, s, g["H_init"], g["W_init"], 5, combine_len=cl, split_rows=sr)
        assert np.max(np.abs(H - base[0]) / np.abs(base[0])) < 1e-12 and np.max(np.abs(W - base[1]) / np.abs(base[1])) < 1e-12
    rng = np.random.default_rng(0)
    pu, pi = rng.permutation(30), rng.permutation(100)          # new id of old id k is p[k] + 1
    H0 = np.empty((30, 10)); H0[pu] = np.array(g["H_init"])
    W0 = np.empty((100, 10)); W0[pi] = np.array(g["W_init"])
    Hp, Wp = orc.run(orc.PPC, pu[u - 1] + 1, pi[i - 1] + 1, s, H0, W0, 5)
    assert np.max(np.abs(Hp[pu] - base[0]) / np.abs(base[0])) < 1e-12
    assert np.max(np.abs(Wp[pi] - base[1]) / np.abs(base[1])) < 1e-12
    # input order of the ratings is irrelevant (they are sorted by (row, column) first)
    sh = rng.permutation(len(u))
    Hs, Ws = orc.run(orc.PPC, u[sh], i[sh], s[sh], g["H_init"], g["W_init"], 5)
    assert np.array_equal(Hs, base[0]) and np.array_equal(Ws, base[1])
