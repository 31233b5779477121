"""The C-ABI library loads on a machine without a GPU and exports every symbol include/hhe_b200.h declares;
compute entry points fail loudly (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

import common

pkg = common.package()


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def libpath():
    if not os.path.exists(pkg.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return pkg.LIB_PATH


def test_header_symbols_are_exported(libpath):
    header = open(os.path.join(common.ROOT, "include", "hhe_b200.h")).read()
    declared = set(re.findall(r"\b(hhe_[a-z0-9_]+)\s*\(", header))
    declared -= {"hhe_ctx"}
    exported = set(re.findall(r" T (hhe_\w+)", subprocess.check_output(["nm", "-D", "--defined-only", libpath], text=True)))
    assert declared, "no declarations found"
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert set(pkg.SYMBOLS) == declared, (sorted(set(pkg.SYMBOLS) ^ declared))


def test_library_loads_and_reports_version(libpath):
    lib = pkg.load_library()
    assert b"sm_100a" in lib.hhe_version()


def test_no_cpu_fallback(libpath):
    if _have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.HheNoDevice):
        pkg.Context(1024, common.T, common.small_params(1024, 6))


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(pkg.HheNoDevice):
        pkg.load_library(str(tmp_path / "libhhe_b200.so"))


def test_bad_parameters_are_invalid_argument(libpath):
    # parameter validation happens before any device is touched
    with pytest.raises(pkg.HheInvalidArgument):
        pkg.Context(1000, common.T, common.small_params(1024, 6))
    with pytest.raises(pkg.HheInvalidArgument):
        pkg.Context(1024, common.T, [15, 17])
    with pytest.raises(pkg.HheInvalidArgument):
        pkg.Context(1024, 65539, common.small_params(1024, 6))


def test_emulation_harness_is_refused_as_a_product_library(libpath):
    """The host emulation of the kernel bodies (tests/emul) is test infrastructure: the package refuses it unless the caller
    states it is the harness, whether it arrives through lib_path or through HHE_B200_LIB (there is no CPU path)."""
    emul = os.path.join(common.ROOT, "tests", "emul", "libhhe_emul.so")
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(emul)])
    assert pkg.load_library().hhe_build_is_cuda() == 1
    with pytest.raises(pkg.HheNoDevice):
        pkg.load_library(emul)
    with pytest.raises(pkg.HheNoDevice):
        pkg.Context(1024, common.T, common.small_params(1024, 6), lib_path=emul)
    lib = pkg.load_library(emul, emulation_harness=True)
    assert lib.hhe_build_is_cuda() == 0 and b"EMULATION" in lib.hhe_version()
    with pytest.raises(pkg.HheNoDevice):  # still refused afterwards (the handle is cached)
        pkg.load_library(emul)
