"""GPU parity of the NMF / PPC clustering step through the C ABI of include/filmyou_nmf.h:
  * against the reference's golden H / W after 1 and 10 iterations (the reference's bar is 1e-4,
    T/util/HadoopIntegrationTest.java:53; asserted here at 1e-10),
  * BIT-EXACT against the CPU oracle (oracle/nmf_oracle.c) run with the same combiner structure,
  * ClusterAssignmentJob / CountClustersJob against T/testdata/ClusteringTestData.java,
  * chained into the RM2 engine: clustering produced on the GPU feeds fy_rm2_set_clustering."""
import json
import os

import numpy as np
import pytest

import filmyou_core_b200 as fy
from filmyou_core_b200 import datagen
from filmyou_core_b200.nmf import NMF, PPC, NmfEngine, cluster_users
from oracle import nmf_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        return json.load(f)


def _gpu(mode, u, i, s, H, W, n_iter, **kw):
    H = np.asarray(H, np.float64); W = np.asarray(W, np.float64)
    with NmfEngine(mode, H.shape[0], W.shape[0], H.shape[1], n_iter, **kw) as eng:
        eng.set_ratings(u, i, s)
        eng.set_factors(H, W)
        eng.run()
        H2, W2 = eng.factors()
        return H2, W2, eng.profile()


@pytest.mark.parametrize("which,mode", [("ppc", PPC), ("nmf", NMF)])
@pytest.mark.parametrize("combine_len,split_rows", [(1024, 256), (0, 0), (8, 16)])
def test_golden_one_and_ten_iterations(which, mode, combine_len, split_rows):
    g = _load("%s_test_data.json" % which)
    u, i, s = orc.coo_from_dense(g["A"])
    for n_iter, hk, wk in ((1, "H_one", "W_one"), (10, "H_ten", "W_ten")):
        H, W, prof = _gpu(mode, u, i, s, g["H_init"], g["W_init"], n_iter, combine_len=combine_len, split_rows=split_rows)
        assert np.max(np.abs(H - np.array(g[hk]))) < 1e-10
        assert np.max(np.abs(W - np.array(g[wk]))) < 1e-10
        Ho, Wo = orc.run(mode, u, i, s, g["H_init"], g["W_init"], n_iter, combine_len=combine_len, split_rows=split_rows)
        assert np.array_equal(H, Ho) and np.array_equal(W, Wo)          # bit-exact
        assert prof["kernel_launches"] > 0 and prof["iterations"] == n_iter
        assert prof["graph_replays"] == (n_iter if n_iter >= 4 else 0)


def test_cluster_assignment_golden():
    g = _load("clustering_test_data.json")
    H = np.array(g["H"])
    with NmfEngine(PPC, H.shape[0], 3, H.shape[1], 0) as eng:
        eng.set_factors(H, np.ones((3, H.shape[1])))
        cl, cnt = eng.cluster_assignment()
    assert cl.tolist() == g["clustering"] and cnt.tolist() == g["clusteringCount"]
    H = np.array([[0.2, 0.7, 0.7], [0.0, 0.0, 0.0], [-1.0, 0.0, -2.0], [-1.0, -0.5, -2.0]])
    with NmfEngine(PPC, 4, 3, 3, 0) as eng:
        eng.set_factors(H, np.ones((3, 3)))
        cl, cnt = eng.cluster_assignment()
    assert cl.tolist() == [1, 0, 1, 1] and cnt.tolist() == [1, 3, 0]


def _dense_items(r):
    """drop never-rated items (the reference throws on them) by renumbering the rated ones 1..M'"""
    ids, inv = np.unique(r.item, return_inverse=True)
    return (inv + 1).astype(np.int32), len(ids)


def _random_factors(rng, n, k):
    F = rng.random((n, k)) + 1e-12
    return F / F.sum(1, keepdims=True)


@pytest.mark.parametrize("shape,k,n_iter,mode", [("tiny", 4, 5, PPC), ("small", 7, 6, NMF), ("ml-100k", 10, 5, PPC),
                                                 ("ml-100k", 33, 4, PPC), ("small", 1, 3, PPC), ("small", 300, 2, PPC)])
def test_synthetic_bit_exact_vs_oracle(shape, k, n_iter, mode):
    r = datagen.generate(shape)
    item, M = _dense_items(r)
    rng = np.random.default_rng(5)
    H0, W0 = _random_factors(rng, r.n_users, k), _random_factors(rng, M, k)
    H, W, prof = _gpu(mode, r.user, item, r.score, H0, W0, n_iter)
    Ho, Wo = orc.run(mode, r.user, item, r.score, H0, W0, n_iter, combine_len=1024, split_rows=256)
    assert np.array_equal(H, Ho) and np.array_equal(W, Wo)
    assert np.array_equal(orc.cluster_assign(Ho)[0], orc.cluster_assign(H)[0])


def test_heavy_rows_use_several_combiner_groups_and_duplicates_are_summed():
    r = datagen.generate("small")
    item, M = _dense_items(r)
    u = np.concatenate([r.user, r.user[:50]]); it = np.concatenate([item, item[:50]])
    s = np.concatenate([r.score, r.score[:50] + 0.5])                # the reference sums duplicate records
    s[7] = 0.0; s[11] = -1.0                                         # dropped (score <= 0)
    rng = np.random.default_rng(9)
    H0, W0 = _random_factors(rng, r.n_users, 5), _random_factors(rng, M, 5)
    H, W, _ = _gpu(PPC, u, it, s, H0, W0, 3, combine_len=4, split_rows=7)
    Ho, Wo = orc.run(PPC, u, it, s, H0, W0, 3, combine_len=4, split_rows=7)
    assert np.array_equal(H, Ho) and np.array_equal(W, Wo)


def test_intended_normalisation_is_opt_in():
    g = _load("ppc_test_data.json")
    u, i, s = orc.coo_from_dense(g["A"])
    for nf, n_iter in ((-1, 10), (3, 7)):
        H, W, _ = _gpu(PPC, u, i, s, g["H_init"], g["W_init"], n_iter, apply_normalization=True, normalization_frequency=nf)
        Ho, Wo = orc.run(PPC, u, i, s, g["H_init"], g["W_init"], n_iter, apply_normalization=True, normalization_frequency=nf,
                         combine_len=1024, split_rows=256)
        assert np.array_equal(H, Ho) and np.array_equal(W, Wo)
    assert np.allclose(H.sum(1), 1.0, atol=1e-3)


def test_errors_match_the_reference_exceptions():
    g = _load("ppc_test_data.json")
    u, i, s = orc.coo_from_dense(g["A"])
    for keep, code, text in ((u != 7, -2, "User 7 has not rated any item"), (i != 42, -10, "Item 42 has not been rated by anybody")):
        with NmfEngine(PPC, 30, 100, 10, 1) as eng:
            eng.set_ratings(u[keep], i[keep], s[keep])
            eng.set_factors(g["H_init"], g["W_init"])
            with pytest.raises(fy.Rm2Error) as e:
                eng.run()
            assert e.value.code == code and text in str(e.value)
    with NmfEngine(PPC, 30, 100, 10, 1) as eng:
        eng.set_ratings(u + 1, i, s)                                  # user 31 is out of range
        eng.set_factors(g["H_init"], g["W_init"])
        with pytest.raises(fy.Rm2Error) as e:
            eng.run()
        assert e.value.code == -1
        with pytest.raises(fy.Rm2Error):
            eng.set_factors(np.ones((3, 3)), np.ones((3, 3)))
    with NmfEngine(PPC, 30, 100, 10, 1) as eng:
        with pytest.raises(fy.Rm2Error) as e:
            eng.run()
        assert e.value.code == -8
    with pytest.raises(fy.Rm2Error) as e:
        NmfEngine(PPC, 30, 100, 513, 1)
    assert e.value.code == -9


def test_random_start_is_seeded_and_row_normalised():
    with NmfEngine(PPC, 200, 300, 12, 0) as eng:
        eng.init_random(3)
        H1, W1 = eng.factors()
        eng.init_random(3)
        H2, W2 = eng.factors()
        eng.init_random(4)
        H3, _ = eng.factors()
    assert np.array_equal(H1, H2) and np.array_equal(W1, W2) and not np.array_equal(H1, H3)
    assert np.allclose(H1.sum(1), 1.0, atol=1e-12) and np.allclose(W1.sum(1), 1.0, atol=1e-12)
    assert H1.min() > 0 and abs(H1.mean() - 1.0 / 12) < 1e-3


def test_clustering_chain_feeds_the_rm2_engine():
    """RMRecommenderDriver.run order: PPC -> cluster assignment -> count -> RM2 (RMRecommenderDriver.java:164-206)."""
    from oracle import rm2_oracle as rm2
    r = datagen.generate("small")
    item, M = _dense_items(r)
    k = 4
    rng = np.random.default_rng(21)
    H0, W0 = _random_factors(rng, r.n_users, k), _random_factors(rng, M, k)
    ids, cl, cnt = cluster_users(r.user, item, r.score, r.n_users, M, k, 8, mode=PPC, H=H0, W=W0)
    Ho, _ = orc.run(PPC, r.user, item, r.score, H0, W0, 8, combine_len=1024, split_rows=256)
    clo, cnto = orc.cluster_assign(Ho)
    assert np.array_equal(cl, clo) and np.array_equal(cnt, cnto)
    if cnt.min() < 2:
        pytest.skip("degenerate clustering for this seed")
    with fy.Rm2Engine(lam=0.1, number_of_items=M, top_n=10) as eng:
        eng.set_ratings(r.user, item, r.score)
        eng.set_clustering(ids, cl, cnt)
        eng.run()
        got = eng.results()
    want = rm2.run(r.user, item, r.score, ids, clo, cnto, 0.1, M, 10)
    assert np.array_equal(got["user"], want["user"]) and np.array_equal(got["item"], want["item"])


def test_file_level_driver_reads_and_writes_sequence_files(tmp_path):
    """PPCHDFSDriverTest at the file level (T/nmf/ppc/PPCHDFSDriverTest.java:40-66): H, W and A SequenceFiles in, ten
    iterations, H and W SequenceFiles out + ClusterAssignmentJob / CountClustersJob outputs."""
    from filmyou_core_b200 import seqfile
    g = _load("ppc_test_data.json")
    u, i, s = orc.coo_from_dense(g["A"])
    d = str(tmp_path)
    os.makedirs(d + "/A")
    seqfile.write_intpair_float(d + "/A/data", u, i, s)                                   # createIntPairFloatFile
    seqfile.write_int_vector(d + "/H", np.arange(1, 31), np.array(g["H_init"]))          # createDoubleMatrix(..., "H", 1)
    seqfile.write_int_vector(d + "/W", np.arange(1, 101), np.array(g["W_init"]))
    with NmfEngine(PPC, 30, 100, 10, 10) as eng:
        eng.run_files(d + "/A", d + "/H", d + "/W", h_out=d + "/H2", w_out=d + "/W2", clustering_out=d + "/clustering",
                      clustering_count_out=d + "/clusteringCount")
    hk, H = seqfile.read_int_vector(d + "/H2")
    wk, W = seqfile.read_int_vector(d + "/W2")
    assert hk.tolist() == list(range(1, 31)) and wk.tolist() == list(range(1, 101))
    assert np.max(np.abs(H - np.array(g["H_ten"]))) < 1e-10 and np.max(np.abs(W - np.array(g["W_ten"]))) < 1e-10   # compareIntVectorData
    ck, cv = seqfile.read_int_int(d + "/clustering")
    cl, cnt = orc.cluster_assign(H)
    assert ck.tolist() == list(range(1, 31)) and np.array_equal(cv, cl)
    nk, nv = seqfile.read_int_int(d + "/clusteringCount")
    assert nk.tolist() == np.flatnonzero(cnt).tolist() and nv.tolist() == cnt[cnt > 0].tolist()
    with NmfEngine(PPC, 30, 100, 10, 3) as eng:                                          # random start, seeded
        eng.run_files(d + "/A", seed=5, h_out=d + "/H3")
        eng2_H = seqfile.read_int_vector(d + "/H3")[1]
    with NmfEngine(PPC, 30, 100, 10, 3) as eng:
        eng.set_ratings(u, i, s); eng.init_random(5); eng.run()
        assert np.array_equal(eng.factors()[0], eng2_H)
    with NmfEngine(PPC, 31, 100, 10, 1) as eng:                                          # H has 30 rows, not numberOfUsers
        with pytest.raises(fy.Rm2Error):
            eng.run_files(d + "/A", d + "/H", d + "/W")


def test_sub_cluster_assignment_golden():
    """TestClusterAssignment.subClusteringTest (T/nmf/clustering/TestClusterAssignment.java:70-101)."""
    from filmyou_core_b200.nmf import sub_cluster_ids
    g = _load("clustering_test_data.json")["subClustering"]
    got = []
    for c, H in enumerate((g["H0"], g["H1"])):
        H = np.array(H)
        with NmfEngine(PPC, H.shape[0], 3, H.shape[1], 0) as eng:
            eng.set_factors(H, np.ones((3, H.shape[1])))
            arg_max, _ = eng.cluster_assignment()
        got += sub_cluster_ids(c, arg_max, g["numberOfUsers"], g["numberOfClusters"]).tolist()
    assert got == g["clustering"]


def test_cluster_refinement_feeds_the_rm2_engine():
    from filmyou_core_b200.nmf import refine_clusters
    from oracle import rm2_oracle as rm2
    r = datagen.generate("small")
    item, M = _dense_items(r)
    cl2, cnt2, made = refine_clusters(r.user, item, r.score, r.cl_user, r.cl_cluster, r.n_clusters, 20, 6, seed=3)
    stride = -(-r.n_users // r.n_clusters)
    assert np.array_equal(cl2 // stride, r.cl_cluster)                 # users stay inside their first-level cluster
    assert made == sum(-(-int(n) // 20) for n in r.cluster_size) and cnt2.sum() == r.n_users
    assert len(np.unique(cl2)) > r.n_clusters
    cl3, _, _ = refine_clusters(r.user, item, r.score, r.cl_user, r.cl_cluster, r.n_clusters, 20, 6, seed=3)
    assert np.array_equal(cl2, cl3)                                    # seeded, deterministic
    if cnt2[cnt2 > 0].min() < 2:
        return
    with fy.Rm2Engine(lam=0.1, number_of_items=M, top_n=5) as eng:
        eng.set_ratings(r.user, item, r.score)
        eng.set_clustering(r.cl_user, cl2, cnt2)
        eng.run()
        got = eng.results()
    want = rm2.run(r.user, item, r.score, r.cl_user, cl2, cnt2, 0.1, M, 5)
    assert np.array_equal(got["user"], want["user"]) and np.array_equal(got["item"], want["item"])
