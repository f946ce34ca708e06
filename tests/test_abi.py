"""CPU suite: the C-ABI library loads and exports every symbol include/filmyou_rm2.h declares
(no compute calls without a GPU), and fails loudly when no device is present."""
import os
import re

import pytest

import filmyou_core_b200 as fy
from filmyou_core_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header="filmyou_rm2.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fy_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_list_the_same_symbols():
    assert _declared() == sorted(engine.EXPORTS)
    assert _declared("filmyou_seqfile.h") == sorted(engine.SEQ_EXPORTS)
    assert _declared("filmyou_nmf.h") == sorted(engine.NMF_EXPORTS)


def test_library_builds_and_exports_every_declared_symbol():
    fy.build_library()
    L = fy.load_library()
    for name in _declared() + _declared("filmyou_seqfile.h") + _declared("filmyou_nmf.h"):
        assert hasattr(L, name), name
    assert L.fy_rm2_abi_version() == 3


def test_default_params_match_the_reference_defaults():
    L = fy.load_library()
    p = fy.Rm2Params()
    L.fy_rm2_default_params(p)
    # M/rmrecommender/RMRecommenderDriver.java:95,114,119
    assert (p.lambda_, p.top_n, p.filter_users, p.shard_count) == (0.1, 1000, 0, 1)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fy.Rm2Error) as e:
        fy.Rm2Engine(number_of_items=10)
    assert e.value.code == -7


def test_no_cpu_fallback_for_the_clustering_step():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from filmyou_core_b200.nmf import PPC, NmfEngine
    with pytest.raises(fy.Rm2Error) as e:
        NmfEngine(PPC, 30, 100, 10)
    assert e.value.code == -7
    L = fy.load_library()
    from filmyou_core_b200.nmf import NmfParams, _lib
    _lib()
    p = NmfParams()
    L.fy_nmf_default_params(p)
    # M/rmrecommender/RMRecommenderDriver.java:94,115 (numberOfIterations 10, normalizationFrequency 12); PPC is what the driver runs
    assert (p.mode, p.number_of_iterations, p.normalization_frequency, p.apply_normalization, p.id_base) == (1, 10, 12, 0, 1)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "filmyou_core_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and False, os.path.join(d, f)


def test_struct_layouts_match_the_header(tmp_path):
    # the ctypes mirrors of fy_rm2_params / fy_rm2_profile (and what a Panama struct layout would declare) against the C header
    import ctypes
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "filmyou_rm2.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %d\\n", sizeof(fy_rm2_params), offsetof(fy_rm2_params, n_gpus), '
                   'sizeof(fy_rm2_profile), offsetof(fy_rm2_profile, ms_gather), offsetof(fy_rm2_profile, score_kernel), FY_RM2_ABI_VERSION); return 0; }\n')
    exe = str(tmp_path / "sizes")
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe])
    sp, on, sf, og, ok, ver = (int(x) for x in subprocess.check_output([exe], text=True).split())
    assert ctypes.sizeof(fy.Rm2Params) == sp == 48 and fy.Rm2Params.n_gpus.offset == on
    from filmyou_core_b200.engine import Rm2Profile
    assert ctypes.sizeof(Rm2Profile) == sf and Rm2Profile.ms_gather.offset == og and Rm2Profile.score_kernel.offset == ok
    assert ver == 3
