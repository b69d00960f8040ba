"""The KITTI AP evaluator on the GPU path: rotated BEV overlaps by pn2_rotate_iou_eval_f32, 3-D overlaps by
pn2_d3_overlap_f64, against the golden vectors of the reference evaluate/eval2.py (tests/golden/kitti_eval.*)."""
import numpy as np
import pytest

from conftest import load
from test_kitti_eval_cpu import _fixture, check_against_golden

pytestmark = pytest.mark.gpu


def test_evaluator_on_gpu_equals_reference_golden(cuda):
    fx, z, gts, dts = _fixture()
    ev = load("evaluate.eval2")
    check_against_golden(ev, z, gts, dts, exact=False)


def test_d3_overlap_kernel_equals_numba_expressions(cuda):
    ev = load("evaluate.eval2")
    rng = np.random.RandomState(3)
    n, k = 37, 53
    def boxes(m):
        return np.stack([rng.uniform(-10, 10, m), rng.uniform(1, 2, m), rng.uniform(5, 30, m), rng.uniform(3, 5, m),
                         rng.uniform(1.3, 1.9, m), rng.uniform(1.4, 2, m), rng.uniform(-3, 3, m)], 1)
    b, q = boxes(n), boxes(k)
    for crit in (-1, 0, 1, 2):
        rinc = np.where(rng.rand(n, k) < 0.5, rng.uniform(0, 8, (n, k)), 0.0).astype(np.float32).astype(np.float64)
        want = rinc.copy()
        for i in range(n):                                   # the literal loop of eval2.py:136-162
            for j in range(k):
                if want[i, j] > 0:
                    iw = min(b[i, 1], q[j, 1]) - max(b[i, 1] - b[i, 4], q[j, 1] - q[j, 4])
                    if iw > 0:
                        area1, area2 = b[i, 3] * b[i, 4] * b[i, 5], q[j, 3] * q[j, 4] * q[j, 5]
                        inc = iw * want[i, j]
                        ua = {-1: area1 + area2 - inc, 0: area1, 1: area2}.get(crit, inc)
                        want[i, j] = inc / ua
                    else:
                        want[i, j] = 0.0
        got = rinc.copy()
        ev.d3_box_overlap_kernel(b, q, got, crit)
        assert np.array_equal(got, want), crit
