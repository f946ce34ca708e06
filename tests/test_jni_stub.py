"""f4: the JNI stub integration/jni/filmyou_rm2_jni.c.  No JDK exists in the build image, so the stub
is compiled against tests/mock_jni/jni.h (the few JNI declarations it uses) and driven through a fake JNIEnv by
tests/jni_mock_harness.c: direct buffers in, one native call per Java method, results read back exactly as
RM2GpuJob / a PPC caller would.  CPU: it compiles warning-free and defines every `native` method the Java classes
declare.  GPU: the harness output equals the oracles'."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import filmyou_core_b200 as fy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "integration", "jni", "filmyou_rm2_jni.c")
JAVA = os.path.join(ROOT, "integration", "java", "es", "udc", "fi", "dc", "irlab")


def _compile_stub(tmp_path):
    obj = str(tmp_path / "jni_stub.o")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fPIC", "-I", os.path.join(ROOT, "tests", "mock_jni"),
                           "-I", os.path.join(ROOT, "include"), "-c", STUB, "-o", obj])
    return obj


def test_stub_compiles_and_defines_every_native_method(tmp_path):
    obj = _compile_stub(tmp_path)
    syms = subprocess.check_output(["nm", "--defined-only", obj], text=True)
    defined = set(re.findall(r"\bT (Java_\w+)", syms))
    for cls, path in (("es_udc_fi_dc_irlab_rm_RM2Native", os.path.join(JAVA, "rm", "RM2Native.java")),
                      ("es_udc_fi_dc_irlab_nmf_ppc_NmfNative", os.path.join(JAVA, "nmf", "ppc", "NmfNative.java"))):
        src = open(path).read()
        natives = re.findall(r"public static native \w+ (\w+)\(", src)
        assert len(natives) >= 9
        for m in natives:
            assert "Java_%s_%s" % (cls, m) in defined, m
    # and nothing but the C ABI is called: every undefined fy_ symbol is declared by the public headers
    undefined = set(re.findall(r"\bU (fy_\w+)", subprocess.check_output(["nm", "-u", obj], text=True)))
    from filmyou_core_b200 import engine
    assert undefined and undefined <= set(engine.EXPORTS) | set(engine.NMF_EXPORTS)


@pytest.mark.gpu
def test_stub_through_a_fake_jnienv_matches_the_oracles(tmp_path, golden, golden_ratings):
    from oracle import nmf_oracle as norc
    from oracle import rm2_oracle as orc
    fy.build_library()
    obj = _compile_stub(tmp_path)
    exe = str(tmp_path / "jni_harness")
    libdir = os.path.dirname(fy.library_path())
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-D_POSIX_C_SOURCE=200809L", "-I", os.path.join(ROOT, "tests", "mock_jni"),
                           os.path.join(ROOT, "tests", "jni_mock_harness.c"), obj, "-L", libdir, "-lfilmyou_rm2",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = golden_ratings
    lines = ["%d %d %d %d %d %r" % (r.nnz, r.n_users, r.n_clusters, golden["numberOfItems"], 10, 0.5)]
    lines += ["%d %d %r" % (u, i, float(s)) for u, i, s in zip(r.user, r.item, r.score)]
    lines += ["%d %d" % (u, c) for u, c in zip(r.cl_user, r.cl_cluster)]
    lines += ["%d" % c for c in r.cluster_size]
    lines.append("%d %d" % (golden["clusterSplit"], golden["splitSize"]))
    p = json.load(open(os.path.join(ROOT, "tests", "golden", "ppc_test_data.json")))
    pu, pi, ps = norc.coo_from_dense(p["A"])
    lines.append("%d %d %d %d %d" % (p["numberOfUsers"], p["numberOfItems"], p["numberOfClusters"], 10, len(pu)))
    lines += ["%d %d %r" % (u, i, float(s)) for u, i, s in zip(pu, pi, ps)]
    lines += ["%r" % v for row in p["H_init"] for v in row] + ["%r" % v for row in p["W_init"] for v in row]
    out = subprocess.run([exe], input="\n".join(lines) + "\n", text=True, capture_output=True, timeout=300)
    assert out.returncode == 0, out.stdout[-500:] + out.stderr[-500:]
    rows = out.stdout.split("\n")
    assert rows[0].startswith("rm2 ")
    n = int(rows[0].split()[1])
    got = np.array([[float(x) for x in ln.split()] for ln in rows[1:1 + n]])
    want = orc.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, r.cluster_size, 0.5, golden["numberOfItems"], 10)
    assert n == len(want["user"])
    assert np.array_equal(got[:, 0], want["user"]) and np.array_equal(got[:, 1], want["item"]) and np.array_equal(got[:, 3], want["cluster"])
    assert np.max(np.abs(got[:, 2] - want["score64"]) / np.abs(want["score64"])) < 1e-9
    assert rows[1 + n].startswith("state -8 ")                       # FY_E_STATE and its message cross the stub
    assert "fy_rm2_run needs" in rows[1 + n]
    assert rows[2 + n] == "capacity -1"                              # an undersized direct buffer is refused (FY_E_ARG)
    assert rows[3 + n].startswith("create 0 java/lang/RuntimeException: fy_rm2_create failed")
    # fine seam through the stub: every "c-split-nSplits" group of TestHDFSRM2's configuration (clusterSplit=5, splitSize=3),
    # bit for bit the coarse seam's triples, each user in the split its id selects (AbstractRM2Reducer.java:203-205)
    assert rows[4 + n] == "fine"
    k1 = 5 + n
    fine = []
    while not rows[k1].startswith("fine-end"):
        fine.append([float(x) for x in rows[k1].split()]); k1 += 1
    fine = np.array(fine)
    assert int(rows[k1].split()[1]) == len(fine) == n
    coarse = {(int(a), int(b)): (c, int(d)) for a, b, c, d in got}
    sizes = dict(zip(range(len(r.cluster_size)), r.cluster_size.tolist()))
    for uu, ii, sc, cl_, sp in fine:
        assert coarse[(int(uu), int(ii))] == (sc, int(cl_))
        K = sizes[int(cl_)]
        n_splits = -(-K // golden["splitSize"]) if K >= golden["clusterSplit"] else 1
        assert int(uu) % n_splits == int(sp)
    k0 = k1 + 1
    assert rows[k0] == "ppc 30 10"
    H = np.array([float(x) for x in rows[k0 + 1:k0 + 1 + 300]]).reshape(30, 10)
    cl = np.array([int(x) for x in rows[k0 + 301:k0 + 331]])
    assert np.max(np.abs(H - np.array(p["H_ten"]))) < 1e-10          # the reference's golden H after ten iterations
    assert np.array_equal(cl, norc.cluster_assign(H)[0])
