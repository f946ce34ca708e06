#!/usr/bin/env python3
"""Builds filmyou_core_b200/libfilmyou_rm2_checked.so = the product sources + -DFY_BOUNDS_CHECK (device-side range checks
on every accumulator / candidate-list / output index, counted in a device global).  Run the GPU suite against it with
    FY_RM2_LIB=$PWD/filmyou_core_b200/libfilmyou_rm2_checked.so python -m pytest tests -m gpu -k "..."
Rm2Engine.close() raises when the count is not zero.  compute-sanitizer is closed on the development pool, this is the
substitute it recommends ("bounds checks and asserts of your own, small cases, a comparison with the CPU reference")."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filmyou_core_b200 import engine
out = os.path.join(os.path.dirname(engine.library_path()), "libfilmyou_rm2_checked.so")
cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-DFY_BOUNDS_CHECK",
       "-Xcompiler", "-fPIC", "-shared", "-o", out] + engine.sources() + ["-ldl"]
subprocess.check_call(cmd)
print(out)
