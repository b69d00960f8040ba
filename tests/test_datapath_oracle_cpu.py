"""The accumulation order the GPU data-path kernels (csrc/scene_prepare.cu, csrc/stat_norm.cu) assume for numpy's
np.dot -- a k-sequential FMA chain, first term a plain multiply (oracle/datapath_oracle.c) -- pinned against numpy on
this machine, for every matrix shape on the reference's data path (lib/utils/calibration.py:51-71 in float32;
utils/kitti_util.py:141-160 and stat_norm/norm.py:197,218 in float64).  If this fails the BLAS here accumulates
differently and the bit-exact GPU parity tests of the data path cannot hold."""
import ctypes

import numpy as np
import pytest

from conftest import ROOT  # noqa: F401  (puts the repo on sys.path)
from oracle import oracle as orc


def chain32(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    out = np.empty((a.shape[0], b.shape[1]), np.float32)
    orc.lib().orc_dot_chain_f32(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(a.shape[0]), a.shape[1],
                                b.ctypes.data_as(ctypes.c_void_p), b.shape[1], out.ctypes.data_as(ctypes.c_void_p))
    return out


def chain64(a, b):
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    out = np.empty((a.shape[0], b.shape[1]), np.float64)
    orc.lib().orc_dot_chain_f64(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(a.shape[0]), a.shape[1],
                                b.ctypes.data_as(ctypes.c_void_p), b.shape[1], out.ctypes.data_as(ctypes.c_void_p))
    return out


@pytest.mark.parametrize("n", [2, 3, 17, 1000, 120000])
def test_float32_products_of_the_pointrcnn_data_path(n):
    rng = np.random.RandomState(n)
    pts = rng.uniform(-80, 80, (n, 3)).astype(np.float32)
    hom = np.hstack((pts, np.ones((n, 1), dtype=np.float32)))
    V2C = rng.uniform(-1, 1, (3, 4)).astype(np.float32)
    R0 = rng.uniform(-1, 1, (3, 3)).astype(np.float32)
    P2 = (rng.uniform(-1, 1, (3, 4)) * np.array([[700, 1, 600, 45]])).astype(np.float32)
    M = np.dot(V2C.T, R0.T)                                            # calibration.py:55
    rect = np.dot(hom, M)
    assert np.array_equal(chain32(hom, M), rect)
    hom2 = np.hstack((rect, np.ones((n, 1), dtype=np.float32)))
    assert np.array_equal(chain32(hom2, P2.T), np.dot(hom2, P2.T))     # calibration.py:62


@pytest.mark.parametrize("n", [2, 5, 1000, 120000])
def test_float64_products_of_the_stat_norm_path(n):
    rng = np.random.RandomState(100 + n)
    velo = rng.uniform(-80, 80, (n, 3)).astype(np.float32)
    hom = np.hstack((velo, np.ones((n, 1))))
    V2C, R0 = rng.uniform(-1, 1, (3, 4)), rng.uniform(-1, 1, (3, 3)) + np.eye(3)
    ref = np.dot(hom, np.transpose(V2C))                               # kitti_util.py:142
    assert np.array_equal(chain64(hom, V2C.T), ref)
    rect = np.transpose(np.dot(R0, np.transpose(ref)))                 # kitti_util.py:154
    assert np.array_equal(chain64(ref, R0.T), rect)
    c, s = np.cos(0.37), np.sin(0.37)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    t = rng.uniform(-10, 10, 3).astype(np.float32)
    d = rect - t
    box = np.dot(d, R)                                                 # norm.py:197
    assert np.array_equal(chain64(d, R), box)
    sub = box[rng.rand(n) < 0.5] * np.array([[1.17, 1.08, 0.96]])
    if len(sub):
        assert np.array_equal(chain64(sub, np.ascontiguousarray(R.T)), np.dot(sub, R.T))      # norm.py:218
    inv = np.linalg.inv(R0)
    back = np.transpose(np.dot(inv, np.transpose(rect)))               # kitti_util.py:151
    assert np.array_equal(chain64(rect, inv.T), back)
