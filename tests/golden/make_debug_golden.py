#!/usr/bin/env python
"""Golden vectors for the debug / stand-alone utilities of the reference's decoder module, produced by the REAL reference:
decode_detections_debug (+ variance_encoded_in_target), greedy_nms, get_num_boxes_per_pred_layer, get_pred_layers
(ssd_encoder_decoder/ssd_output_decoder.py:27-75, 342-530).  Writes tests/golden/ref_debug_golden.npz.  Build container only."""
import os
import sys

import numpy as np

np.float = float   # noqa
np.int = int       # noqa
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get('SSD_REFERENCE_ROOT', '/root/reference'))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder                                                   # noqa: E402
from ssd_encoder_decoder.ssd_output_decoder import (decode_detections, decode_detections_debug, get_num_boxes_per_pred_layer,  # noqa: E402
                                                    get_pred_layers, greedy_nms)
from oracle import synth                                                                                            # noqa: E402

TINY = dict(img_height=120, img_width=160, n_classes=3, predictor_sizes=[(6, 8), (3, 4)], scales=[0.2, 0.45, 0.8],
            aspect_ratios_global=[0.5, 1.0, 2.0], two_boxes_for_ar1=True, variances=[0.1, 0.1, 0.2, 0.2], normalize_coords=True)


def main():
    arrays = {}
    for coords in ('centroids', 'corners', 'minmax'):
        enc = SSDInputEncoder(coords=coords, **TINY)
        anchors = np.concatenate([b.reshape(-1, 4) for b in enc.boxes_list], axis=0)
        yp = synth.synth_y_pred(71, 3, anchors.astype(np.float32), 4, sharp=3.0, loc_scale=0.3)
        arrays['dbg/%s/y_pred' % coords] = yp
        for tag, kw in (('a', dict(confidence_thresh=0.05, iou_threshold=0.45, top_k=200)),
                        ('topk', dict(confidence_thresh=0.05, iou_threshold=0.45, top_k=7)),
                        ('vit', dict(confidence_thresh=0.05, iou_threshold=0.45, top_k=200, variance_encoded_in_target=True))):
            res = decode_detections_debug(yp, input_coords=coords, normalize_coords=True, img_height=120, img_width=160, **kw)
            for i, r in enumerate(res):
                arrays['dbg/%s/%s/out%d' % (coords, tag, i)] = np.asarray(r, np.float64).reshape(-1, 7)
            if tag == 'a':
                layers = get_pred_layers(res, get_num_boxes_per_pred_layer(TINY['predictor_sizes'], [[0.5, 1.0, 2.0]] * 2, True))
                for i, l in enumerate(layers):
                    arrays['dbg/%s/layers%d' % (coords, i)] = np.asarray(l, np.int64)
    arrays['dbg/num_boxes'] = np.asarray(get_num_boxes_per_pred_layer(TINY['predictor_sizes'], [[0.5, 1.0, 2.0]] * 2, True), np.int64)
    arrays['dbg/num_boxes_one'] = np.asarray(get_num_boxes_per_pred_layer([(38, 38), (19, 19)], [[1.0, 2.0, 0.5], [1.0, 2.0, 0.5, 3.0, 1 / 3]], False), np.int64)
    # greedy_nms on decoded boxes: random overlapping boxes incl. exact duplicates and score ties
    rng = np.random.default_rng(9)
    items = []
    for n in (40, 1, 17):
        xy = rng.uniform(0, 80, (n, 2)); wh = rng.uniform(10, 60, (n, 2))
        b = np.concatenate([rng.integers(1, 4, (n, 1)).astype(np.float64), rng.uniform(0, 1, (n, 1)), xy, xy + wh], axis=1)
        if n > 10:
            b[3] = b[2]; b[5, 1] = b[4, 1]
        items.append(b)
    for i, b in enumerate(items):
        arrays['nms/in%d' % i] = b
    for bp in ('half', 'include', 'exclude'):
        for thr in (0.45, 0.1):
            res = greedy_nms(items, iou_threshold=thr, coords='corners', border_pixels=bp)
            for i, r in enumerate(res):
                arrays['nms/%s/%g/out%d' % (bp, thr, i)] = np.asarray(r, np.float64)
    np.savez_compressed(os.path.join(HERE, 'ref_debug_golden.npz'), **arrays)
    print('wrote %d arrays' % len(arrays))


if __name__ == '__main__':
    main()
