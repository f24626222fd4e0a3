"""CPU tests of the boundary: the C-ABI library builds, loads and exports exactly what
include/minsnap_b200.h declares; argument errors come back as codes; and there is no CPU
fallback -- compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "minsnap_b200.h")


def declared_functions():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MINSNAP_API\s+[\w\s\*]+?\b(minsnap_\w+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("minsnap_reorder", "minsnap_solve", "minsnap_solve_standard", "minsnap_sample_uniform",
                 "minsnap_sample_at", "minsnap_evaluate_range", "minsnap_cost_sweep", "minsnap_solve_standard_host"):
        assert must in names
    assert len(names) >= 25


def test_library_exports_every_declared_symbol(ms):
    from mav_trajectory_generation_cmake_b200 import capi
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH]).decode()
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    declared = declared_functions()
    missing = [n for n in declared if n not in exported]
    assert not missing, missing
    # and nothing undeclared leaks out of the library
    extra = [n for n in exported if n.startswith("minsnap_") and n not in declared]
    assert not extra, extra
    # the ctypes table covers the same set
    assert sorted(capi.SIGNATURES) == declared


def test_abi_version_and_strings(ms):
    lib = ms.load()
    assert lib.minsnap_abi_version() == 1
    assert lib.minsnap_error_string(0) == b"ok"
    assert b"argument" in lib.minsnap_error_string(1)


def test_argument_errors_are_codes_not_crashes(ms):
    lib = ms.load()
    from mav_trajectory_generation_cmake_b200 import capi
    # unsupported N, K < 1, NULL pointers
    assert lib.minsnap_solve_workspace_bytes(7, 10) == 0
    assert lib.minsnap_solve_workspace_bytes(10, 10) > 0
    assert lib.minsnap_reorder(10, 0, 1, None, None, None, None) == capi.ERR_ARG
    assert lib.minsnap_solve(1, 10, 3, 10, 4, None, None, None, None, None, None, None, None, None, 0, None) == capi.ERR_ARG
    assert lib.minsnap_solve_standard(1, 10, 3, 10, 7, None, None, None, 0.0, 0.0, 0.0, None, None, None, None, None,
                                      None) == capi.ERR_ARG
    assert lib.minsnap_sample_uniform(1, 10, 3, 10, None, None, 10, 5, None, None, None) == capi.ERR_ARG
    assert lib.minsnap_fp64_peak(0, None) == capi.ERR_ARG


def test_random_positions_match_oracle_generator(ms, oracle):
    lo, hi = [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0]
    got = ms.random_positions_host(33, 10, lo, hi, 12345)
    want = np.stack([oracle.create_random_positions(10, lo, hi, 12345 + b) for b in range(33)])
    assert np.array_equal(got, want)
    got1 = ms.random_positions_host(5, 100, [-50.0], [50.0], 0)
    want1 = np.stack([oracle.create_random_positions(100, [-50.0], [50.0], b) for b in range(5)])
    assert np.array_equal(got1, want1)


def test_no_cpu_fallback_without_gpu(ms):
    """Without a CUDA device every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU-only machines")
    with pytest.raises(ms.MinsnapError) as ei:
        ms.device_info()
    assert ei.value.code in (2, 4)
    pos = np.zeros((2, 11, 3))
    pos[:, :, 0] = np.arange(11)
    with pytest.raises(ms.MinsnapError):
        ms.solve_standard_host(pos, times=np.ones((2, 10)))
    with pytest.raises(ms.MinsnapError):
        ms.segment_matrices_host([1.0, 2.0])
    with pytest.raises(ms.MinsnapError):
        ms.fp64_peak()


def test_product_package_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "mav_trajectory_generation_cmake_b200")
    inc = os.path.join(ROOT, "include")
    offenders = []
    for base in (pkg, inc):
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"oracle_py|liboracle|orc_\w+\(|from oracle|import oracle", text):
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
