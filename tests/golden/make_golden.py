#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the RM2 path into JSON fixtures.

Runs ONLY in the build container (needs /root/reference, which does not exist on
the GPU box).  It parses the Java array literals of the reference's test data;
no reference source is copied, only the numeric vectors the reference's tests
assert against:

  T/testdata/RMTestData.java:25-27    numberOfUsers / numberOfItems / numberOfClusters
  T/testdata/RMTestData.java:32-232   A[item][user] ratings (0 = no rating)
  T/testdata/RMTestData.java:234-403  recommendations: 507 (user, item, score) triples
  T/testdata/RMTestData.java:408-410  userSum
  T/testdata/RMTestData.java:415-421  itemSum
  T/testdata/RMTestData.java:426      totalSum
  T/testdata/RMTestData.java:431-464  itemColl = p(i|C)
  T/testdata/ClusteringTestData.java:90-93  clustering (user -> cluster), clusteringCount
  T/testdata/RMTestData2.java:25-60   5x3 toy with hand-checkable sums
  T/util/HadoopIntegrationTest.java:53,81-100  accuracy=1e-4, lambda=0.5, N=1000,
                                               clusterSplit=5, splitSize=3

(T/ = /root/reference/src/test/java/es/udc/fi/dc/irlab/)

Usage:  python tests/golden/make_golden.py   -> writes tests/golden/*.json
"""
import json
import os
import re
import sys

T = "/root/reference/src/test/java/es/udc/fi/dc/irlab/"
HERE = os.path.dirname(os.path.abspath(__file__))


def java_array(src, name):
    """Return the nested python list of the Java array literal assigned to `name`."""
    m = re.search(r"\b" + re.escape(name) + r"\s*=\s*new\s+\w+(\[\])+\s*\{", src)
    if not m:
        raise KeyError(name)
    i = m.end() - 1
    depth = 0
    j = i
    while True:
        c = src[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    lit = src[i:j + 1]
    lit = re.sub(r"//[^\n]*", "", lit)
    lit = lit.replace("{", "[").replace("}", "]")
    lit = re.sub(r"(\d)\.(?=\s*[,\]])", r"\1.0", lit)  # "238." -> "238.0"
    lit = re.sub(r",\s*\]", "]", lit)
    return json.loads(lit)


def java_scalar(src, name):
    m = re.search(r"\b" + re.escape(name) + r"\s*=\s*([-0-9.eE]+)\s*;", src)
    return float(m.group(1))


def main():
    if not os.path.isdir(T):
        sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
    rm = open(T + "testdata/RMTestData.java").read()
    cl = open(T + "testdata/ClusteringTestData.java").read()
    rm2 = open(T + "testdata/RMTestData2.java").read()

    A = java_array(rm, "A")
    recs = java_array(rm, "recommendations")
    out = {
        "source": "filmyou-core T/testdata/RMTestData.java + ClusteringTestData.java (numeric vectors only)",
        "numberOfUsers": int(java_scalar(rm, "numberOfUsers")),
        "numberOfItems": int(java_scalar(rm, "numberOfItems")),
        "numberOfClusters": int(java_scalar(rm, "numberOfClusters")),
        "lambda": 0.5,                 # T/util/HadoopIntegrationTest.java:96
        "numberOfRecommendations": 1000,  # :84
        "clusterSplit": 5,             # :97
        "splitSize": 3,                # :98
        "accuracy": 1e-4,              # :53
        "A_item_by_user": A,
        "recommendations": recs,
        "userSum": java_array(rm, "userSum"),
        "itemSum": java_array(rm, "itemSum"),
        "totalSum": java_scalar(rm, "totalSum"),
        "itemColl": java_array(rm, "itemColl"),
        "clustering": java_array(cl, "clustering"),
        "clusteringCount": java_array(cl, "clusteringCount"),
    }
    assert len(A) == 100 and all(len(r) == 30 for r in A)
    assert len(recs) == 507
    with open(os.path.join(HERE, "rm_test_data.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))

    out2 = {
        "source": "filmyou-core T/testdata/RMTestData2.java (numeric vectors only)",
        "numberOfUsers": int(java_scalar(rm2, "numberOfUsers")),
        "numberOfItems": int(java_scalar(rm2, "numberOfItems")),
        "numberOfClusters": int(java_scalar(rm2, "numberOfClusters")),
        "A_item_by_user": java_array(rm2, "A"),
        "userSum": java_array(rm2, "userSum"),
        "itemSum": java_array(rm2, "itemSum"),
        "totalSum": java_scalar(rm2, "totalSum"),
        "itemColl": java_array(rm2, "itemColl"),
        "clustering": java_array(rm2, "clustering"),
        "clusteringCount": java_array(rm2, "clusteringCount"),
    }
    with open(os.path.join(HERE, "rm_test_data2.json"), "w") as f:
        json.dump(out2, f, separators=(",", ":"))
    print("wrote rm_test_data.json (%d ratings, %d golden triples), rm_test_data2.json"
          % (sum(1 for r in A for x in r if x > 0), len(recs)))


if __name__ == "__main__":
    main()
