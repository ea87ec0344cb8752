"""The TensorFlow/Keras half of the oracle against golden vectors produced by THE REFERENCE'S OWN SOURCE
(keras_loss_function/keras_ssd_loss.py and keras_layers/keras_layer_{DecodeDetections,DecodeDetectionsFast,L2Normalization,
AnchorBoxes}.py) executed over a NumPy stand-in for the TensorFlow/Keras primitives they call (tests/golden/tf_shim.py,
tests/golden/make_tf_golden.py; TensorFlow cannot be installed offline).  This pins the oracle to the reference's code for
everything but the primitives' own semantics, which tf_shim.py states."""
import os

import numpy as np
import pytest
import torch

from oracle.anchors import anchor_boxes_for_layer
from oracle.decoder import decode_layer, decode_layer_fast
from oracle.loss import ssd_loss
from oracle.model import l2_normalize

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_tf_shim_golden.npz'))


@pytest.mark.parametrize('key', ['plain', 'ratio2_alpha', 'no_pos', 'no_pos_negmin', 'ties', 'neutral', 'zero_neg_losses'])
def test_ssd_loss_matches_reference_code(key):
    ratio, n_neg_min, alpha = G['loss/%s/kw' % key]
    out = ssd_loss(G['loss/%s/y_true' % key], G['loss/%s/y_pred' % key], int(ratio), int(n_neg_min), float(alpha))
    ref = G['loss/%s/out' % key]
    assert out.shape == ref.shape
    np.testing.assert_allclose(out, ref, rtol=2e-6, atol=1e-6)       # float32 summation order is the only freedom


@pytest.mark.parametrize('fast', [False, True])
@pytest.mark.parametrize('case', ['default', 'cap', 'topk_small', 'nonorm', 'none'])
def test_decode_layers_match_reference_code(case, fast):
    key = ('fast_' if fast else 'layer_') + case
    conf, iou, top_k, cap, norm = G['dec/%s/kw' % key]
    fn = decode_layer_fast if fast else decode_layer
    with np.errstate(all='ignore'):
        out = fn(G['dec/y_pred'], confidence_thresh=float(conf), iou_threshold=float(iou), top_k=int(top_k),
                 nms_max_output_size=int(cap), normalize_coords=bool(norm), img_height=120, img_width=160)
    ref = G['dec/%s/out' % key]
    assert out.shape == ref.shape                                   # (B, top_k, 6), zero padded
    np.testing.assert_array_equal(out[..., 0], ref[..., 0])         # class ids, row by row (same order, same padding)
    np.testing.assert_array_equal(out[..., 1], ref[..., 1])         # confidences are copied, not computed
    np.testing.assert_allclose(out[..., 2:], ref[..., 2:], rtol=1e-6, atol=1e-4)


def test_l2_normalization_matches_reference_code():
    x = G['l2norm/x']                                               # NHWC
    out = l2_normalize(torch.from_numpy(x).permute(0, 3, 1, 2), np.full((x.shape[-1],), 20.0, np.float32))
    np.testing.assert_allclose(out.permute(0, 2, 3, 1).numpy(), G['l2norm/out'], rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize('key,fmap,kw', [
    ('tiny0', (6, 8), dict(img_height=120, img_width=160, this_scale=0.2, next_scale=0.45, aspect_ratios=[0.5, 1.0, 2.0],
                           two_boxes_for_ar1=True, coords='centroids', normalize_coords=True)),
    ('tiny1_clip_corners', (3, 4), dict(img_height=120, img_width=160, this_scale=0.45, next_scale=0.8, aspect_ratios=[0.5, 3.0],
                                        two_boxes_for_ar1=False, this_steps=(40, 41), this_offsets=(0.4, 0.6), clip_boxes=True,
                                        coords='corners', normalize_coords=False)),
    ('ssd300_conv4_3', (38, 38), dict(img_height=300, img_width=300, this_scale=0.1, next_scale=0.2, aspect_ratios=[1.0, 2.0, 0.5],
                                      two_boxes_for_ar1=True, this_steps=8, this_offsets=0.5, clip_boxes=False,
                                      coords='centroids', normalize_coords=True)),
])
def test_anchor_boxes_layer_matches_reference_code(key, fmap, kw):
    ref = G['anchors/%s/out' % key]                                 # (B, H, W, n_boxes, 8) float32: boxes | variances
    a = anchor_boxes_for_layer(feature_map_size=fmap, **kw)
    assert ref.shape[1:4] == a.shape[:3]
    for b in range(ref.shape[0]):                                   # tiled over the batch
        np.testing.assert_array_equal(ref[b, ..., :4], a.astype(np.float32))
        np.testing.assert_array_equal(ref[b, ..., 4:], np.broadcast_to(np.float32([0.1, 0.1, 0.2, 0.2]), ref[b, ..., 4:].shape))


# ---------------------------------------------------------------------------------------------------------------
# The model graphs: oracle/model.py against the outputs of the reference's REAL builders (models/keras_ssd300.py,
# keras_ssd512.py, keras_ssd7.py) executed eagerly over stand-in Keras layers (tf_shim.make_keras_layers).  Pins which layer
# feeds which, paddings, the dilated fc6, pooling modes, reshape / concatenate order, L2Normalization placement, the
# AnchorBoxes arguments per source layer and the decoder wiring of mode='inference' / 'inference_fast'.
# ---------------------------------------------------------------------------------------------------------------
def _vgg_w(seed, variant, n_cls):
    from oracle import synth
    from oracle.model import vgg_weight_shapes
    w = synth.synth_weights(seed, vgg_weight_shapes(variant, n_cls), bias_scale=0.02)
    w['conv4_3_norm/gamma'] = np.random.default_rng(seed).uniform(10, 30, 512).astype(np.float32)
    return w


PRE = dict(subtract_mean=[123, 117, 104], divide_by_stddev=[64, 64, 64], swap_channels=[2, 1, 0])


@pytest.fixture(scope='module')
def ssd300_oracle_output():
    from oracle import synth
    from oracle.model import ssd_vgg_forward
    x = synth.synth_images(41, 1, 300, 300)
    return ssd_vgg_forward(x, _vgg_w(42, 300, 20), 300, 20, scales=[0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05], **PRE)


def test_ssd300_graph_matches_reference_builder(ssd300_oracle_output):
    y = ssd300_oracle_output
    assert y.shape == (1, 8732, 33)
    np.testing.assert_allclose(y[:, ::7], G['model/ssd300/rows7'], rtol=2e-3, atol=1e-4)     # two float32 evaluation orders, 23 layers
    np.testing.assert_allclose(y.astype(np.float64).sum(axis=1), G['model/ssd300/colsum'], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize('mode', ['inference', 'inference_fast'])
def test_ssd300_inference_modes_match_reference_builder(ssd300_oracle_output, mode):
    fn = decode_layer if mode == 'inference' else decode_layer_fast
    out = fn(ssd300_oracle_output, 0.01, 0.45, 200, 400, True, 300, 300)
    ref = G['model/ssd300/' + mode]
    assert out.shape == ref.shape == (1, 200, 6)
    np.testing.assert_array_equal(out[..., 0], ref[..., 0])
    np.testing.assert_allclose(out[..., 1], ref[..., 1], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out[..., 2:], ref[..., 2:], rtol=1e-4, atol=2e-3)


def test_ssd512_graph_matches_reference_builder():
    from oracle import synth
    from oracle.model import ssd_vgg_forward
    x = synth.synth_images(43, 1, 512, 512)
    y = ssd_vgg_forward(x, _vgg_w(44, 512, 20), 512, 20, scales=[0.04, 0.1, 0.26, 0.42, 0.58, 0.74, 0.9, 1.06], **PRE)
    assert y.shape == (1, 24564, 33)
    np.testing.assert_allclose(y[:, ::16], G['model/ssd512/rows16'], rtol=2e-3, atol=1e-4)
    np.testing.assert_allclose(y.astype(np.float64).sum(axis=1), G['model/ssd512/colsum'], rtol=1e-5, atol=1e-3)


def test_ssd7_graph_matches_reference_builder():
    """300 x 480 input: also pins the height / width order of every anchor and reshape."""
    from oracle import synth
    from oracle.model import ssd7_forward, ssd7_weight_shapes
    x = synth.synth_images(45, 1, 300, 480)
    w = synth.synth_weights(46, ssd7_weight_shapes(5), bias_scale=0.05)
    rng = np.random.default_rng(47)
    for i in range(1, 8):
        c = w['conv%d/bias' % i].shape[0]
        w['bn%d/gamma' % i] = rng.uniform(0.8, 1.2, c).astype(np.float32)
        w['bn%d/beta' % i] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        w['bn%d/moving_mean' % i] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        w['bn%d/moving_variance' % i] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    y = ssd7_forward(x, w, n_classes=5, scales=[0.08, 0.16, 0.32, 0.64, 0.96], normalize_coords=True, subtract_mean=127.5,
                     divide_by_stddev=127.5)
    assert tuple(G['model/ssd7/shape']) == y.shape
    np.testing.assert_allclose(y[:, ::5], G['model/ssd7/rows5'], rtol=2e-3, atol=1e-4)
    np.testing.assert_allclose(y.astype(np.float64).sum(axis=1), G['model/ssd7/colsum'], rtol=1e-5, atol=1e-3)
