#!/usr/bin/env python
"""Golden vectors for the smaller public helpers of the reference's hot-path modules (REAL reference code; the SSDLoss
helpers over tests/golden/tf_shim.py): intersection_area, SSDInputEncoder.generate_anchor_boxes_for_layer,
SSDLoss.smooth_L1_loss / log_loss.  Writes tests/golden/ref_api_extra.npz.  Build container only."""
import os
import sys

import numpy as np

np.float = float   # noqa
np.int = int       # noqa
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get('SSD_REFERENCE_ROOT', '/root/reference'))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

tf_shim.install()
from bounding_box_utils.bounding_box_utils import convert_coordinates, intersection_area      # noqa: E402
from keras_loss_function.keras_ssd_loss import SSDLoss                                         # noqa: E402
from ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder                              # noqa: E402


def main():
    rng = np.random.default_rng(17)
    arrays = {}

    def boxes(n):
        xy = rng.uniform(0, 60, (n, 2)); wh = rng.uniform(4, 50, (n, 2))
        return np.concatenate([xy, xy + wh], axis=1)                    # corners
    b1, b2 = boxes(7), boxes(5)
    b2[0] = b1[0]                                                       # identical boxes
    b2[1, :2] = b1[1, 2:]                                               # touching at a corner
    arrays['inter/b1'], arrays['inter/b2'] = b1, b2
    for coords in ('corners', 'minmax', 'centroids'):
        c1 = b1 if coords == 'corners' else convert_coordinates(b1, 0, 'corners2' + coords)
        c2 = b2 if coords == 'corners' else convert_coordinates(b2, 0, 'corners2' + coords)
        for border in ('half', 'include', 'exclude'):
            arrays['inter/%s/%s/outer' % (coords, border)] = intersection_area(c1, c2, coords=coords, mode='outer_product', border_pixels=border)
            arrays['inter/%s/%s/elem' % (coords, border)] = intersection_area(c1[:5], c2, coords=coords, mode='element-wise', border_pixels=border)

    enc = SSDInputEncoder(img_height=120, img_width=160, n_classes=3, predictor_sizes=[(6, 8), (3, 4)], scales=[0.2, 0.45, 0.8],
                          aspect_ratios_per_layer=[[1.0, 2.0], [0.5, 3.0]], two_boxes_for_ar1=True, steps=[20, (40, 41)],
                          offsets=[0.5, (0.4, 0.6)], clip_boxes=True, variances=[0.1, 0.1, 0.2, 0.2], matching_type='bipartite',
                          pos_iou_threshold=0.5, neg_iou_limit=0.2, normalize_coords=False)
    arrays['anchors_layer/a'] = enc.generate_anchor_boxes_for_layer((5, 7), [1.0, 2.0, 0.5], 0.3, 0.5)
    bx, centers, wh, step, off = enc.generate_anchor_boxes_for_layer((3, 4), [0.5, 3.0], 0.45, 0.8, this_steps=(40, 41),
                                                                     this_offsets=(0.4, 0.6), diagnostics=True)
    arrays['anchors_layer/b'] = bx
    arrays['anchors_layer/b_cy'], arrays['anchors_layer/b_cx'] = centers
    arrays['anchors_layer/b_wh'] = wh
    arrays['anchors_layer/b_step'] = np.array(step, dtype=np.float64)
    arrays['anchors_layer/b_off'] = np.array(off, dtype=np.float64)

    L = SSDLoss()
    yt = rng.standard_normal((2, 9, 4)).astype(np.float32) * 1.5
    yp = rng.standard_normal((2, 9, 4)).astype(np.float32) * 1.5
    arrays['loss_helpers/l1_true'], arrays['loss_helpers/l1_pred'] = yt, yp
    arrays['loss_helpers/l1_out'] = np.asarray(L.smooth_L1_loss(yt, yp), np.float32)
    p = rng.uniform(0, 1, (2, 9, 5)).astype(np.float32); p[0, 0] = 0                      # a zero probability: the 1e-15 clamp
    t = np.eye(5, dtype=np.float32)[rng.integers(0, 5, (2, 9))]
    arrays['loss_helpers/log_true'], arrays['loss_helpers/log_pred'] = t, p
    arrays['loss_helpers/log_out'] = np.asarray(L.log_loss(t, p), np.float32)

    np.savez_compressed(os.path.join(HERE, 'ref_api_extra.npz'), **arrays)
    print('wrote %d arrays' % len(arrays))


if __name__ == '__main__':
    main()
