#!/usr/bin/env python3
"""Extract the reference's golden vectors for the NMF / PPC clustering step.

Runs ONLY in the build container (needs /root/reference).  Parses the Java array literals of the
reference's test data -- numeric vectors only, no reference source is copied:

  T/testdata/PPCTestData.java:25-27     numberOfUsers=30 / numberOfItems=100 / numberOfClusters=10
  T/testdata/PPCTestData.java:60-263    A[item][user] ratings (0 = none)
  T/testdata/PPCTestData.java:265,570   W_init, H_init
  T/testdata/PPCTestData.java:665,969   W_one, H_one   (H_one asserted by T/nmf/ppc/hcomputation/TestHDFSPPCHComputation.java:66)
  T/testdata/PPCTestData.java:1064,1369 W_ten, H_ten   (asserted by T/nmf/ppc/PPCHDFSDriverTest.java:62-63)
  T/testdata/NMFTestData.java           same fields for plain NMF
                                        (T/nmf/hcomputation/TestHDFSHComputation.java:66, T/nmf/wcomputation/TestHDFSWComputation.java:66,
                                         T/nmf/NMFHDFSDriverTest.java:62-63)
  T/testdata/ClusteringTestData.java:28-93  H (30x5) -> clustering (arg-max) -> clusteringCount
                                            (T/nmf/clustering/TestClusterAssignment.java)
  T/util/HadoopIntegrationTest.java:53  accuracy = 1e-4; normalizationFrequency is NOT set by the tests, so
                                        `iteration % -1 == 0` holds and PPC L1-normalises every iteration

(T/ = /root/reference/src/test/java/es/udc/fi/dc/irlab/)

Usage:  python tests/golden/make_golden_nmf.py   -> writes tests/golden/{ppc,nmf,clustering}_test_data.json
"""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import T, HERE, java_array, java_scalar  # noqa: E402


def factor_fixture(path, label):
    src = open(path).read()
    out = {
        "source": "filmyou-core T/testdata/%s (numeric vectors only)" % label,
        "numberOfUsers": int(java_scalar(src, "numberOfUsers")),
        "numberOfItems": int(java_scalar(src, "numberOfItems")),
        "numberOfClusters": int(java_scalar(src, "numberOfClusters")),
        "accuracy": 1e-4,
        "eps": 1e-12,                      # M/nmf/MatrixComputationJob.java:41
        "normalizationFrequency": -1,      # unset in the tests -> conf.getInt(..., -1)
    }
    for name in ("A", "W_init", "H_init", "W_one", "H_one", "W_ten", "H_ten"):
        out[name] = java_array(src, name)
    assert len(out["A"]) == out["numberOfItems"] and len(out["A"][0]) == out["numberOfUsers"]
    for name in ("W_init", "W_one", "W_ten"):
        assert len(out[name]) == out["numberOfItems"] and len(out[name][0]) == out["numberOfClusters"], name
    for name in ("H_init", "H_one", "H_ten"):
        assert len(out[name]) == out["numberOfUsers"] and len(out[name][0]) == out["numberOfClusters"], name
    return out


def main():
    if not os.path.isdir(T):
        sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
    for label, fname in (("PPCTestData.java", "ppc_test_data.json"), ("NMFTestData.java", "nmf_test_data.json")):
        out = factor_fixture(T + "testdata/" + label, label)
        if label.startswith("PPC"):
            src = re.sub(r"(\d)d\b", r"\1", open(T + "testdata/" + label).read())   # "2.585d" -> "2.585"
            for name in ("Ap", "h0p", "w0p", "w1p", "h1p"):   # the 5x7 hand example of the unit tests
                out[name] = java_array(src, name)
        with open(os.path.join(HERE, fname), "w") as f:
            json.dump(out, f, separators=(",", ":"))
        print("wrote", fname, os.path.getsize(os.path.join(HERE, fname)), "bytes")
    cl = open(T + "testdata/ClusteringTestData.java").read()
    out = {
        "source": "filmyou-core T/testdata/ClusteringTestData.java (numeric vectors only)",
        "numberOfUsers": int(java_scalar(cl, "numberOfUsers")),
        "numberOfClusters": int(java_scalar(cl, "numberOfClusters")),
        "H": java_array(cl, "H"),
        "clustering": java_array(cl, "clustering"),
        "clusteringCount": java_array(cl, "clusteringCount"),
    }
    # T/testdata/SubClusteringTestData.java:25-96: two clusters of 17 and 13 users, their H files (keys 1..17 and
    # 18..30, T/nmf/clustering/TestClusterAssignment.java:79-82) and the expected sub-cluster ids
    # cluster * ceil(numberOfUsers / numberOfClusters) + arg-max (M/nmf/clustering/FindSubClusterMapper.java:52-77)
    sc = open(T + "testdata/SubClusteringTestData.java").read()
    sc = sc.replace("double[][] H1 = {", "double[][] H1 = new double[][] {")
    out["subClustering"] = {
        "numberOfUsers": int(java_scalar(sc, "numberOfUsers")),
        "numberOfClusters": int(java_scalar(sc, "numberOfClusters")),
        "numberOfSubClusters": int(java_scalar(sc, "numberOfSubClusters")),
        "H0": java_array(sc, "H0"),
        "H1": java_array(sc, "H1"),
        "clustering": java_array(sc, "clustering"),
    }
    with open(os.path.join(HERE, "clustering_test_data.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote clustering_test_data.json")


if __name__ == "__main__":
    main()
