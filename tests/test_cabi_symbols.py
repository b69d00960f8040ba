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


def test_product_path_has_no_cpu_fallback():
    """Without a CUDA device the product fails loudly instead of computing on the CPU: the model's forward raises at
    its first kernel call and bench.py exits non-zero with a message (the CPU port under oracle/ is never reached)."""
    import subprocess
    import sys
    import pytest
    import torch
    from conftest import ROOT
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cabi = load("cabi")
    model = load("inference").build_model(seed=0, device="cpu")
    with pytest.raises(cabi.Pn2Error):
        model({"pts_input": torch.zeros(1, 16384, 3)})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode != 0 and "CUDA device" in (r.stderr + r.stdout)
    assert '"metric"' not in r.stdout                       # no bench line was printed
