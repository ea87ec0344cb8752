"""The two non-trivial primitives the TensorFlow stand-in (tests/golden/tf_shim.py) restates, checked independently:
`tf.image.non_max_suppression` against torchvision's C++ NMS on tie-free inputs, and the `tf.nn.top_k` tie rule."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import tf_shim  # noqa: E402


def test_top_k_prefers_lower_index_on_ties():
    x = np.array([0.5, 0.9, 0.5, 0.9, 0.1], np.float32)
    r = tf_shim._top_k(x, 4)
    assert list(r.indices) == [1, 3, 0, 2] and np.array_equal(r.values, x[[1, 3, 0, 2]])


def test_nms_cap_and_threshold_are_strict():
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30], [40, 40, 50, 50]], np.float32)
    scores = np.array([0.9, 0.8, 0.7, 0.6], np.float32)
    assert list(tf_shim._nms(boxes, scores, 10, 0.5)) == [0, 2, 3]          # IoU 1 > 0.5 drops box 1
    assert list(tf_shim._nms(boxes, scores, 10, 1.0)) == [0, 1, 2, 3]       # IoU == threshold is kept (suppress iff >)
    assert list(tf_shim._nms(boxes, scores, 2, 0.5)) == [0, 2]              # max_output_size
    deg = np.array([[0, 0, 0, 10], [0, 0, 0, 10]], np.float32)              # zero-area boxes never suppress each other
    assert list(tf_shim._nms(deg, np.array([0.9, 0.8], np.float32), 10, 0.0)) == [0, 1]


def test_nms_matches_torchvision_on_random_boxes():
    tv = pytest.importorskip('torchvision')
    import torch
    rng = np.random.default_rng(0)
    for n, thr in ((300, 0.45), (1000, 0.3), (50, 0.7)):
        xy = rng.uniform(0, 100, (n, 2)); wh = rng.uniform(5, 40, (n, 2))
        b = np.concatenate([xy, xy + wh], axis=1).astype(np.float32)         # (x0, y0, x1, y1)
        s = rng.permutation(n).astype(np.float32) / n                        # distinct scores: no ties
        ref = tv.ops.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
        got = tf_shim._nms(b[:, [1, 0, 3, 2]], s, n, thr)                    # TensorFlow's (y0, x0, y1, x1) order
        assert np.array_equal(got, ref)
