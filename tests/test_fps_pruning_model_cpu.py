"""The exactness claim of csrc/fps_cells.cu, pinned without a GPU: furthest point sampling that skips every cell whose
bounding-box lower bound (the reference's own float expression) is not below the cell's current maximum min-distance gives
the indices and the running distances of the literal restatement of sampling_gpu.cu:93-209 (oracle.fps), bit for bit --
whatever the assignment of points to cells (spatially sorted, the identity, a random permutation) and whatever the ties."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc

syn = importlib.import_module("3d_adapt_auto_driving_b200.synthetic")


def _orders(xyz, rng):
    n = xyz.shape[0]
    q = np.floor((xyz[:, [0, 2]] - xyz[:, [0, 2]].min(0)) / 2.0).astype(np.int64)        # 2 m grid, row-major: spatially coherent
    return {"grid": np.lexsort((q[:, 1], q[:, 0])), "identity": np.arange(n), "random": rng.permutation(n)}


@pytest.mark.parametrize("kind,n,m,cell", [("lidar", 4096, 700, 128), ("uniform", 3000, 500, 128), ("ties", 2500, 2500, 128),
                                           ("ties", 1024, 300, 32), ("lidar", 777, 777, 64)])
def test_pruned_sampling_is_exact_for_any_cell_assignment(kind, n, m, cell):
    rng = np.random.RandomState(n + m)
    xyz = syn.make_clouds(kind, 1, n, seed=n)[0]
    ref_idx, ref_temp = orc.fps(xyz[None], m)
    ncell = (n + cell - 1) // cell
    touched = {}
    for name, order in _orders(xyz, rng).items():
        idx, temp, t = orc.fps_pruned_model(xyz, m, order, cell=cell)
        assert np.array_equal(idx, ref_idx[0]), (kind, name, int((idx != ref_idx[0]).sum()))
        assert np.array_equal(temp, ref_temp[0]), (kind, name)
        touched[name] = t / float(max(m - 1, 1) * ncell)
    # the pruning only pays with a spatial order (that is what the kernel's Hilbert prepass is for)
    assert touched["grid"] < touched["random"]


def test_pruned_sampling_with_caller_distances():
    """the reference's scratch may come pre-filled (temp): maxima and boxes start from it"""
    rng = np.random.RandomState(3)
    xyz = syn.make_clouds("lidar", 1, 1500, seed=5)[0]
    temp0 = rng.uniform(0.0, 4.0, size=(1, 1500)).astype(np.float32)
    ref_idx, ref_temp = orc.fps(xyz[None], 200, temp=temp0)
    idx, temp, _ = orc.fps_pruned_model(xyz, 200, _orders(xyz, rng)["grid"], cell=128, temp=temp0[0])
    assert np.array_equal(idx, ref_idx[0]) and np.array_equal(temp, ref_temp[0])
