#!/usr/bin/env python3
"""Builds tests/cpp/bin/test_dropin against include/ and libminsnap_b200.so (g++, no CUDA
headers needed: the C++ mirror only sees the C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIBDIR = os.path.join(ROOT, "mav_trajectory_generation_cmake_b200", "lib")
BIN = os.path.join(HERE, "bin", "test_dropin")
SRC = os.path.join(HERE, "test_dropin.cpp")
TIMING_BIN = os.path.join(HERE, "bin", "timing_evaluation")
TIMING_SRC = os.path.join(ROOT, "tools", "cpp", "timing_evaluation.cpp")


def build(force=False):
    deps = [SRC, TIMING_SRC] + [os.path.join(ROOT, "include", "mav_trajectory_generation", f)
                    for f in os.listdir(os.path.join(ROOT, "include", "mav_trajectory_generation"))]
    deps.append(os.path.join(ROOT, "include", "minsnap_b200.h"))
    if not force and os.path.exists(BIN) and os.path.exists(TIMING_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(BIN) for d in deps):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", BIN,
           "-L", LIBDIR, "-lminsnap_b200", "-Wl,-rpath," + LIBDIR]
    subprocess.check_call(cmd)
    # the reference's timing procedure through the mirror (tools/cpp/timing_evaluation.cpp)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), TIMING_SRC, "-o",
                           TIMING_BIN, "-L", LIBDIR, "-lminsnap_b200", "-Wl,-rpath," + LIBDIR])
    return BIN


if __name__ == "__main__":
    print("built", build(force="--force" in sys.argv))
