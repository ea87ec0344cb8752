#!/usr/bin/env python
"""Golden vectors for the TensorFlow/Keras half of the reference, produced by executing THE REFERENCE'S OWN SOURCE
(keras_loss_function/keras_ssd_loss.py, keras_layers/keras_layer_{DecodeDetections,DecodeDetectionsFast,L2Normalization,
AnchorBoxes}.py) over tests/golden/tf_shim.py, a NumPy stand-in for the ~45 TensorFlow / Keras primitives that code calls
(TensorFlow itself cannot be installed here).  Run in the build container only:

    python tests/golden/make_tf_golden.py        # writes tests/golden/ref_tf_shim_golden.npz

What the vectors pin and what they assume is spelled out in tf_shim.py.  Inputs are stored next to the outputs (they are
small) so that tests/test_oracle_tf_shim_golden.py needs neither the reference nor the shim.
"""
import os
import sys

import numpy as np

np.float = float   # noqa  the reference targets NumPy < 1.24 (caller-side aliases, as in make_golden.py)
np.int = int       # noqa

REF = os.environ.get('SSD_REFERENCE_ROOT', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

import tf_shim  # noqa: E402

tf_shim.install()

from keras_layers.keras_layer_AnchorBoxes import AnchorBoxes                      # noqa: E402  (the reference's files)
from keras_layers.keras_layer_DecodeDetections import DecodeDetections            # noqa: E402
from keras_layers.keras_layer_DecodeDetectionsFast import DecodeDetectionsFast    # noqa: E402
from keras_layers.keras_layer_L2Normalization import L2Normalization              # noqa: E402
from keras_loss_function.keras_ssd_loss import SSDLoss                            # noqa: E402

from models.keras_ssd300 import ssd_300                                            # noqa: E402  (the reference's builders)
from models.keras_ssd512 import ssd_512                                            # noqa: E402
from models.keras_ssd7 import build_model                                          # noqa: E402

from oracle import synth                                                           # noqa: E402
from oracle.encoder import OracleEncoder                                           # noqa: E402
from oracle.model import ssd7_weight_shapes, vgg_weight_shapes                     # noqa: E402

TINY = dict(img_height=120, img_width=160, n_classes=3, predictor_sizes=[(6, 8), (3, 4)], scales=[0.2, 0.45, 0.8],
            aspect_ratios_global=[0.5, 1.0, 2.0], two_boxes_for_ar1=True, variances=[0.1, 0.1, 0.2, 0.2],
            pos_iou_threshold=0.5, neg_iou_limit=0.3, normalize_coords=True)


def main():
    arrays = {}
    enc = OracleEncoder(**TINY)                                   # pinned bit-exact against the real SSDInputEncoder
    P, C = enc.anchors.shape[0], enc.n_classes

    # ---- SSDLoss.compute_loss (keras_ssd_loss.py:98-211) -------------------------------------------------
    def loss_case(key, y_true, y_pred, **kw):
        L = SSDLoss(**kw)
        out = L.compute_loss(np.asarray(y_true, np.float32), np.asarray(y_pred, np.float32))
        arrays['loss/%s/y_true' % key] = np.asarray(y_true, np.float32)
        arrays['loss/%s/y_pred' % key] = np.asarray(y_pred, np.float32)
        arrays['loss/%s/kw' % key] = np.array([kw.get('neg_pos_ratio', 3), kw.get('n_neg_min', 0), kw.get('alpha', 1.0)], np.float64)
        arrays['loss/%s/out' % key] = np.asarray(out, np.float32)

    yt, yp = synth.synth_y_true_pred_for_loss(31, enc, 3, 4, 3, sharp=2.0)
    loss_case('plain', yt, yp)
    loss_case('ratio2_alpha', yt, yp, neg_pos_ratio=2, alpha=0.5)
    yt0 = yt.copy(); yt0[:, :, :C] = 0; yt0[:, :, 0] = 1                        # background only -> n_positive = 0
    loss_case('no_pos', yt0, yp)
    loss_case('no_pos_negmin', yt0, yp, n_neg_min=7)
    ypt = yp.copy(); ypt[:, :, :C] = np.array([0.5, 0.25, 0.125, 0.125], np.float32)   # every negative loss identical: top_k ties
    loss_case('ties', yt, ypt)
    ytn = yt.copy(); ytn[0, :20, :C] = 0                                         # neutral boxes
    loss_case('neutral', ytn, yp)
    yp1 = yp.copy(); yp1[:, :, :C] = 0; yp1[:, :, 0] = 1                         # all negative losses are exactly zero
    loss_case('zero_neg_losses', yt, yp1)

    # ---- DecodeDetections / DecodeDetectionsFast (.call) ----------------------------------------------------
    ypd = synth.synth_y_pred(21, 3, enc.anchors, C, sharp=3.0, loc_scale=1.0)
    arrays['dec/y_pred'] = ypd

    def dec_case(key, cls, **kw):
        layer = cls(**kw)
        arrays['dec/%s/out' % key] = np.asarray(layer.call(np.asarray(ypd, np.float32)), np.float32)
        arrays['dec/%s/kw' % key] = np.array([kw['confidence_thresh'], kw['iou_threshold'], kw['top_k'], kw['nms_max_output_size'],
                                              1.0 if kw.get('normalize_coords', True) else 0.0], np.float64)

    for name, cls in (('layer', DecodeDetections), ('fast', DecodeDetectionsFast)):
        dec_case(name + '_default', cls, confidence_thresh=0.01, iou_threshold=0.45, top_k=200, nms_max_output_size=400,
                 normalize_coords=True, img_height=120, img_width=160)
        dec_case(name + '_cap', cls, confidence_thresh=0.05, iou_threshold=0.6, top_k=10, nms_max_output_size=5,
                 normalize_coords=True, img_height=120, img_width=160)
        dec_case(name + '_topk_small', cls, confidence_thresh=0.2, iou_threshold=0.3, top_k=4, nms_max_output_size=400,
                 normalize_coords=True, img_height=120, img_width=160)
        dec_case(name + '_nonorm', cls, confidence_thresh=0.3, iou_threshold=0.45, top_k=50, nms_max_output_size=100,
                 normalize_coords=False, img_height=120, img_width=160)
        dec_case(name + '_none', cls, confidence_thresh=0.999, iou_threshold=0.45, top_k=20, nms_max_output_size=30,
                 normalize_coords=True, img_height=120, img_width=160)

    # ---- L2Normalization (.call, keras_layer_L2Normalization.py:61-63) ----------------------------------------
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((2, 5, 4, 16)) * 3).astype(np.float32)
    x[0, 0, 0] = 0                                                 # an all-zero pixel: the epsilon clamp
    l2 = L2Normalization(gamma_init=20)
    arrays['l2norm/x'] = x
    arrays['l2norm/out'] = np.asarray(l2(x), np.float32)

    # ---- AnchorBoxes (.call, keras_layer_AnchorBoxes.py:133-255) ----------------------------------------------
    def anchor_case(key, fmap, **kw):
        layer = AnchorBoxes(**kw)
        out = layer.call(tf_shim.keras_tensor(np.zeros((2,) + fmap + (8,), np.float32)))
        arrays['anchors/%s/out' % key] = np.asarray(out, np.float32)

    anchor_case('tiny0', (6, 8), img_height=120, img_width=160, this_scale=0.2, next_scale=0.45, aspect_ratios=[0.5, 1.0, 2.0],
                two_boxes_for_ar1=True, variances=[0.1, 0.1, 0.2, 0.2], coords='centroids', normalize_coords=True)
    anchor_case('tiny1_clip_corners', (3, 4), img_height=120, img_width=160, this_scale=0.45, next_scale=0.8,
                aspect_ratios=[0.5, 3.0], two_boxes_for_ar1=False, this_steps=(40, 41), this_offsets=(0.4, 0.6), clip_boxes=True,
                variances=[0.1, 0.1, 0.2, 0.2], coords='corners', normalize_coords=False)
    anchor_case('ssd300_conv4_3', (38, 38), img_height=300, img_width=300, this_scale=0.1, next_scale=0.2,
                aspect_ratios=[1.0, 2.0, 0.5], two_boxes_for_ar1=True, this_steps=8, this_offsets=0.5, clip_boxes=False,
                variances=[0.1, 0.1, 0.2, 0.2], coords='centroids', normalize_coords=True)

    # ---- the model builders themselves (models/keras_ssd300.py, keras_ssd512.py, keras_ssd7.py), executed eagerly ---------
    # Keras layers are eager NumPy/torch-CPU stand-ins (tf_shim.make_keras_layers); weights and the image are synthetic and
    # regenerated from their seeds in the test.  Stored: every `stride`-th prior row of the (1, P, C+12) output, plus the
    # decoded (1, top_k, 6) output of mode='inference' / 'inference_fast'.
    def run_builder(builder, x, w, **kw):
        tf_shim.STATE['input'], tf_shim.STATE['weights'] = x, w
        return np.asarray(builder(**kw).output, np.float32)

    def vgg_w(seed, variant, n_cls):
        w = synth.synth_weights(seed, vgg_weight_shapes(variant, n_cls), bias_scale=0.02)
        w['conv4_3_norm/gamma'] = np.random.default_rng(seed).uniform(10, 30, 512).astype(np.float32)
        return w

    pre = dict(subtract_mean=[123, 117, 104], divide_by_stddev=[64, 64, 64], swap_channels=[2, 1, 0])
    sc300 = [0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05]
    x = synth.synth_images(41, 1, 300, 300)
    w = vgg_w(42, 300, 20)
    common = dict(image_size=(300, 300, 3), n_classes=20, scales=sc300, **pre)
    y = run_builder(ssd_300, x, w, mode='training', **common)
    assert y.shape == (1, 8732, 33), y.shape
    arrays['model/ssd300/rows7'] = y[:, ::7]
    arrays['model/ssd300/colsum'] = y.astype(np.float64).sum(axis=1)
    arrays['model/ssd300/inference'] = run_builder(ssd_300, x, w, mode='inference', confidence_thresh=0.01, iou_threshold=0.45,
                                                   top_k=200, nms_max_output_size=400, **common)
    arrays['model/ssd300/inference_fast'] = run_builder(ssd_300, x, w, mode='inference_fast', confidence_thresh=0.01,
                                                        iou_threshold=0.45, top_k=200, nms_max_output_size=400, **common)

    sc512 = [0.04, 0.1, 0.26, 0.42, 0.58, 0.74, 0.9, 1.06]
    x = synth.synth_images(43, 1, 512, 512)
    w = vgg_w(44, 512, 20)
    y = run_builder(ssd_512, x, w, mode='training', image_size=(512, 512, 3), n_classes=20, scales=sc512, **pre)
    assert y.shape == (1, 24564, 33), y.shape
    arrays['model/ssd512/rows16'] = y[:, ::16]
    arrays['model/ssd512/colsum'] = y.astype(np.float64).sum(axis=1)

    x = synth.synth_images(45, 1, 300, 480)
    w = synth.synth_weights(46, ssd7_weight_shapes(5), bias_scale=0.05)
    rng = np.random.default_rng(47)
    for i in range(1, 8):
        c = w['conv%d/bias' % i].shape[0]
        w['bn%d/gamma' % i] = rng.uniform(0.8, 1.2, c).astype(np.float32)
        w['bn%d/beta' % i] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        w['bn%d/moving_mean' % i] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        w['bn%d/moving_variance' % i] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    y = run_builder(build_model, x, w, mode='training', image_size=(300, 480, 3), n_classes=5, scales=[0.08, 0.16, 0.32, 0.64, 0.96],
                    normalize_coords=True, subtract_mean=127.5, divide_by_stddev=127.5)
    arrays['model/ssd7/rows5'] = y[:, ::5]
    arrays['model/ssd7/colsum'] = y.astype(np.float64).sum(axis=1)
    arrays['model/ssd7/shape'] = np.array(y.shape)

    np.savez_compressed(os.path.join(HERE, 'ref_tf_shim_golden.npz'), **arrays)
    print('wrote %d arrays' % len(arrays))


if __name__ == '__main__':
    main()
