"""CPU: the C-ABI library builds, loads and exports every symbol include/motif_b200.h declares
(no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from motif_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "motif_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(motif_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_lists():
    assert declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_abi_version_and_launch_counter(lib):
    assert lib.motif_abi_version() == 5
    lib.motif_reset_launch_count()
    assert lib.motif_launch_count() == 0


def test_workspace_queries_need_no_gpu(lib):
    assert lib.motif_splat_workspace_bytes(1, 720, 1280) > 720 * 1280 * 4 * 17
    assert lib.motif_splat_workspace_bytes(0, 1, 1) == 0
    n = lib.motif_decode_workspace_bytes(1, 7, 180, 320, 720, 1280)
    assert n >= 720 * 1280 * 4 * (2 * 64 + 128 + 4 + 1)


def test_bad_arguments_are_reported_not_crashed(lib):
    rc = lib.motif_splat_fwd(None, None, None, None, 1, 1, 1, 1, 0, None, ctypes.c_size_t(0), None)
    assert rc == -1 and b"null" in lib.motif_last_error()
    rc = lib.motif_corr_fwd(None, None, None, 1, 1, 1, 1, None)
    assert rc == -1
    rc = lib.motif_decode(None, None)
    assert rc == -1


def test_ctypes_struct_matches_header_layout(lib):
    # 6 ints + 4 pointers + 1 float (+ padding) for the geometry block
    assert ctypes.sizeof(_lib.GeomT) == 6 * 4 + 4 * 8 + 8
    assert ctypes.sizeof(_lib.SirenT) == 8 + 5 * 8 * 2
    assert ctypes.sizeof(_lib.DecodeT) == lib.motif_sizeof_decode_t()  # the binding's struct is the library's


def test_operators_refuse_cpu_tensors():
    import torch
    from motif_b200 import correlation, softsplat_cp

    x = torch.zeros(1, 2, 4, 4)
    with pytest.raises(NotImplementedError):
        softsplat_cp.FunctionSoftsplat(x, x, None, "summation")
    with pytest.raises(NotImplementedError):
        correlation.FunctionCorrelation(x, x)
