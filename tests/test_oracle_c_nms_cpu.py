"""The C restatement of tf.image.non_max_suppression (oracle/tf_nms.c, what bench.py's CPU arm times) against the two NumPy
restatements it follows (oracle/decoder.py: tf_nms -- the literal one -- and tf_nms_fast): identical selections, in order, on random
boxes, exact score ties, zero-area / reversed-corner / inf / NaN boxes, caps below the survivor count and empty inputs; and
decode_layer with it plugged in returns the same tensor as with the NumPy version."""
import numpy as np
import pytest

from oracle import cbuild
from oracle.decoder import decode_layer, tf_nms, tf_nms_c, tf_nms_fast


@pytest.fixture(scope='module', autouse=True)
def _compiled():
    if cbuild.build_c() is None:
        pytest.skip('no C compiler on this host')


def _boxes(rng, n, scale, degenerate=False):
    b = rng.uniform(0, scale, (n, 4)).astype(np.float32)
    if degenerate:
        b[rng.integers(0, n, n // 8)] = 0.0                                   # zero-area boxes
        k = rng.integers(0, n, n // 10); b[k, 2] = b[k, 0]                     # zero width
        k = rng.integers(0, n, n // 16); b[k, 1] = np.inf
        k = rng.integers(0, n, n // 16); b[k, 3] = np.nan
        k = rng.integers(0, n, n // 16); b[k, 0] = -np.inf
    return b


@pytest.mark.parametrize('seed', range(6))
def test_matches_the_literal_restatement_on_small_inputs(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 120))
    b = _boxes(rng, n, 40.0, degenerate=seed % 2 == 1)
    s = rng.uniform(0.02, 1.0, n).astype(np.float32)
    if seed % 3 == 0:
        s = np.round(s * 4) / 4                                              # many exact ties -> lower index first
    for cap in (1, 5, 400):
        for thr in (0.45, 0.0, 0.9):
            with np.errstate(all='ignore'):
                assert tf_nms_c(b, s, cap, thr) == tf_nms(b, s, cap, thr) == tf_nms_fast(b, s, cap, thr)


@pytest.mark.parametrize('n,scale,deg', [(8732, 300.0, False), (8732, 20.0, True), (3000, 5.0, True), (24564, 512.0, False)])
def test_matches_the_vectorised_restatement_at_model_sizes(n, scale, deg):
    rng = np.random.default_rng(n)
    b = _boxes(rng, n, scale, deg)
    s = rng.uniform(0.011, 1.0, n).astype(np.float32)
    with np.errstate(all='ignore'):
        assert tf_nms_c(b, s, 400, 0.45) == tf_nms_fast(b, s, 400, 0.45)


def test_empty_and_single():
    z = np.zeros((0, 4), np.float32)
    assert tf_nms_c(z, np.zeros(0, np.float32), 400, 0.45) == []
    assert tf_nms_c(np.array([[0, 0, 1, 1]], np.float32), np.array([0.5], np.float32), 400, 0.45) == [0]
    assert tf_nms_c(np.array([[0, 0, 1, 1]], np.float32), np.array([0.5], np.float32), 0, 0.45) == []


def test_decode_layer_with_the_c_nms_is_the_same_tensor():
    rng = np.random.default_rng(5)
    B, P, C = 2, 600, 6
    y = np.zeros((B, P, C + 12), np.float32)
    logits = rng.standard_normal((B, P, C)).astype(np.float32) * 2
    e = np.exp(logits - logits.max(-1, keepdims=True)); y[..., :C] = e / e.sum(-1, keepdims=True)
    y[..., C:C + 4] = rng.standard_normal((B, P, 4)).astype(np.float32) * 0.5
    y[..., C + 4:C + 6] = rng.uniform(0.1, 0.9, (B, P, 2)); y[..., C + 6:C + 8] = rng.uniform(0.05, 0.4, (B, P, 2))
    y[..., C + 8:] = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
    with np.errstate(all='ignore'):
        a, ia = decode_layer(y, 0.01, 0.45, 50, 30, True, 300, 300, return_indices=True)
        b, ib = decode_layer(y, 0.01, 0.45, 50, 30, True, 300, 300, return_indices=True, nms=tf_nms_c)
    assert np.array_equal(a, b) and np.array_equal(ia, ib)
