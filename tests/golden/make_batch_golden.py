#!/usr/bin/env python
"""Golden vectors for the device-side batch assembly: the REAL reference's CropPad / Flip / Resize / BoxFilter
(data_generator/object_detection_2d_patch_sampling_ops.py, ..._geometric_ops.py, ..._image_boxes_validation_utils.py) applied
to synthetic label arrays on dummy images, in the order of the original SSD chain (expand -> crop -> flip -> resize), plus the
degenerate-box removal of DataGenerator.generate.  Writes tests/golden/ref_batch_golden.npz.  Build container only."""
import json
import os
import sys

import numpy as np

np.float = float   # noqa
np.int = int       # noqa
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get('SSD_REFERENCE_ROOT', '/root/reference'))

from data_generator.object_detection_2d_geometric_ops import Flip, Resize                         # noqa: E402
from data_generator.object_detection_2d_image_boxes_validation_utils import BoxFilter             # noqa: E402
from data_generator.object_detection_2d_patch_sampling_ops import CropPad                         # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    arrays, cases = {}, []
    B = 24
    for b in range(B):
        H, W = int(rng.integers(60, 200)), int(rng.integers(60, 240))
        n = int(rng.integers(0, 12))
        x0 = rng.integers(0, W - 8, n); y0 = rng.integers(0, H - 8, n)
        w = rng.integers(1, W // 2, n); h = rng.integers(1, H // 2, n)
        labels = np.stack([rng.integers(1, 21, n), x0, y0, np.minimum(x0 + w, W - 1), np.minimum(y0 + h, H - 1)], axis=1).astype(np.int64).reshape(-1, 5)
        if b % 5 == 4 and n:
            labels = labels.astype(np.float64) + rng.uniform(0, 1, labels.shape) * (np.arange(5) > 0)       # float labels too
        if b % 7 == 3 and n > 1:
            labels[1, 3] = labels[1, 1]                                                                      # a degenerate box from the start
        image = np.zeros((H, W, 3), np.uint8)
        ops = []
        lab = labels.copy()
        img = image
        if b % 2 == 0:                                               # SSDExpand: larger canvas, image somewhere inside, no filter, no clip
            ph, pw = int(H * rng.uniform(1, 3)), int(W * rng.uniform(1, 3))
            py, px = -int(rng.integers(0, ph - H + 1)), -int(rng.integers(0, pw - W + 1))
            img, lab = CropPad(py, px, ph, pw, clip_boxes=False, box_filter=None)(img, lab)
            ops.append(['crop_pad', py, px, ph, pw, False, False])
        if b % 3 != 1:                                               # SSDRandomCrop: centre-point filter + clip
            ih, iw = img.shape[:2]
            ph, pw = int(ih * rng.uniform(0.3, 1.0)), int(iw * rng.uniform(0.3, 1.0))
            py, px = int(rng.integers(0, ih - ph + 1)), int(rng.integers(0, iw - pw + 1))
            bf = BoxFilter(check_overlap=True, check_min_area=False, check_degenerate=False, overlap_criterion='center_point')
            img, lab = CropPad(py, px, ph, pw, clip_boxes=True, box_filter=bf)(img, lab)
            ops.append(['crop_pad', py, px, ph, pw, True, True])
        if b % 2 == 1:
            ops.append(['flip', img.shape[1], 'horizontal'])
            img, lab = Flip(dim='horizontal')(img, lab)
        if b % 11 == 5:
            ops.append(['flip', img.shape[0], 'vertical'])
            img, lab = Flip(dim='vertical')(img, lab)
        ih, iw = img.shape[:2]
        bf = BoxFilter(check_overlap=False, check_min_area=False, check_degenerate=True)
        img, lab = Resize(height=300, width=300, box_filter=bf)(img, lab)
        ops.append(['resize', ih, iw, 300, 300, True])
        if b % 4 == 2:
            bf2 = BoxFilter(check_overlap=False, check_min_area=True, check_degenerate=True, min_area=400)
            lab = bf2(lab)
            ops.append(['filter', True, 400])
        arrays['in%d' % b] = labels.astype(np.float64)
        arrays['out%d' % b] = np.asarray(lab, dtype=np.float64).reshape(-1, 5)
        cases.append(ops)
    np.savez_compressed(os.path.join(HERE, 'ref_batch_golden.npz'), **arrays)
    with open(os.path.join(HERE, 'ref_batch_golden.json'), 'w') as f:
        json.dump({'n': B, 'ops': cases}, f)
    print('wrote', len(arrays), 'arrays; boxes in', sum(arrays['in%d' % b].shape[0] for b in range(B)), 'out',
          sum(arrays['out%d' % b].shape[0] for b in range(B)))


if __name__ == '__main__':
    main()
