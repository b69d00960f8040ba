"""The C-ABI library loads without a GPU and exports every symbol include/pn2_b200.h declares."""
import ctypes
import os

from conftest import load


def test_library_is_built_and_exports_header_symbols():
    import __graft_entry__ as entry
    entry.build()
    cabi = load("cabi")
    assert os.path.exists(cabi.LIB_PATH)
    lib = ctypes.CDLL(cabi.LIB_PATH)
    names = cabi.declared_symbols()
    assert len(names) >= 14
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.pn2_abi_version.restype = ctypes.c_int
    assert lib.pn2_abi_version() >= 1
    # pure host helper (no device needed)
    assert lib.pn2_fps_ref_block_size(16384) == 1024 and lib.pn2_fps_ref_block_size(512) == 512


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    p2 = load("pointnet2_cuda")
    cabi = load("cabi")
    x = torch.zeros(1, 8, 3)
    with pytest.raises(cabi.Pn2Error):
        p2.furthest_point_sampling_wrapper(1, 8, 2, x, torch.zeros(1, 8), torch.zeros(1, 2, dtype=torch.int32))
