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
