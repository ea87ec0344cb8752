#!/usr/bin/env python
"""Golden vectors for the Evaluator: the REAL reference's match_predictions / compute_precision_recall /
compute_average_precisions / compute_mean_average_precision (eval_utils/average_precision_evaluator.py:490-905) on synthetic
predictions and ground truth (stand-in object for the DataGenerator attributes it reads).  'mergesort' is used so that equal
confidences have a defined order.  Writes tests/golden/ref_eval_golden.npz.  Build container only."""
import contextlib
import io
import os
import sys
import types

import numpy as np

np.float = float   # noqa
np.int = int       # noqa
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get('SSD_REFERENCE_ROOT', '/root/reference'))

import warnings                                                                   # noqa: E402
warnings.simplefilter('ignore')
from eval_utils.average_precision_evaluator import Evaluator                    # noqa: E402


def synth(seed, n_images, n_classes, n_gt, n_pred_per_image, neutral_frac, ties):
    rng = np.random.default_rng(seed)
    labels, neutral, preds = [], [], [list() for _ in range(n_classes + 1)]
    ids = ['%06d' % (i * 3 + 1) for i in range(n_images)]
    for i in range(n_images):
        g = int(rng.integers(0, n_gt + 1))
        x0 = rng.integers(0, 200, g); y0 = rng.integers(0, 200, g)
        lab = np.stack([rng.integers(1, n_classes + 1, g), x0, y0, x0 + rng.integers(8, 90, g), y0 + rng.integers(8, 90, g)], axis=1).reshape(-1, 5)
        labels.append(lab)
        neutral.append(rng.uniform(0, 1, g) < neutral_frac)
        for _ in range(int(rng.integers(0, n_pred_per_image + 1))):
            if g and rng.uniform() < 0.7:                      # a jittered copy of a ground-truth box
                j = int(rng.integers(0, g))
                b = lab[j, 1:].astype(np.float64) + rng.normal(0, 6, 4)
                c = int(lab[j, 0]) if rng.uniform() < 0.85 else int(rng.integers(1, n_classes + 1))
            else:
                xy = rng.uniform(0, 200, 2); b = np.concatenate([xy, xy + rng.uniform(8, 90, 2)])
                c = int(rng.integers(1, n_classes + 1))
            conf = float(np.float32(rng.uniform(0.01, 1)))
            if ties:
                conf = round(conf, 1)
            preds[c].append((ids[i], conf, round(float(b[0]), 1), round(float(b[1]), 1), round(float(b[2]), 1), round(float(b[3]), 1)))
    return ids, labels, neutral, preds


def main():
    arrays = {}
    cases = {'a': dict(seed=1, n_images=40, n_classes=5, n_gt=6, n_pred_per_image=12, neutral_frac=0.0, ties=False),
             'neutral': dict(seed=2, n_images=30, n_classes=4, n_gt=7, n_pred_per_image=10, neutral_frac=0.3, ties=False),
             'ties': dict(seed=3, n_images=25, n_classes=3, n_gt=5, n_pred_per_image=15, neutral_frac=0.1, ties=True)}
    for name, kw in cases.items():
        ids, labels, neutral, preds = synth(**kw)
        for use_neutral in (False, True):
            dg = types.SimpleNamespace(labels=labels, image_ids=ids, eval_neutral=neutral if use_neutral else None)
            for bp in ('include', 'half'):
                ev = Evaluator(model=None, n_classes=kw['n_classes'], data_generator=dg)
                ev.prediction_results = preds
                with contextlib.redirect_stdout(io.StringIO()):
                    ev.get_num_gt_per_class(ignore_neutral_boxes=True, verbose=False)
                    tp, fp, ctp, cfp = ev.match_predictions(ignore_neutral_boxes=True, matching_iou_threshold=0.5, border_pixels=bp,
                                                            sorting_algorithm='mergesort', verbose=True, ret=True)
                    prec, rec = ev.compute_precision_recall(verbose=False, ret=True)
                    ap_s = ev.compute_average_precisions(mode='sample', num_recall_points=11, verbose=False, ret=True)
                    map_s = ev.compute_mean_average_precision()
                    ap_i = ev.compute_average_precisions(mode='integrate', verbose=False, ret=True)
                    map_i = ev.compute_mean_average_precision()
                key = '%s/%d/%s' % (name, int(use_neutral), bp)
                arrays[key + '/num_gt'] = np.asarray(ev.num_gt_per_class)
                for c in range(1, kw['n_classes'] + 1):
                    arrays[key + '/tp%d' % c] = np.asarray(tp[c]); arrays[key + '/fp%d' % c] = np.asarray(fp[c])
                    arrays[key + '/ctp%d' % c] = np.asarray(ctp[c]) if len(np.atleast_1d(tp[c])) else np.zeros(0, int)
                    arrays[key + '/cfp%d' % c] = np.asarray(cfp[c]) if len(np.atleast_1d(fp[c])) else np.zeros(0, int)
                    arrays[key + '/prec%d' % c] = np.asarray(prec[c], dtype=np.float64); arrays[key + '/rec%d' % c] = np.asarray(rec[c], dtype=np.float64)
                arrays[key + '/ap_sample'] = np.asarray(ap_s, dtype=np.float64); arrays[key + '/ap_integrate'] = np.asarray(ap_i, dtype=np.float64)
                arrays[key + '/map'] = np.asarray([map_s, map_i], dtype=np.float64)
        # the inputs
        arrays[name + '/ids'] = np.asarray(ids)
        for i, (l, e) in enumerate(zip(labels, neutral)):
            arrays[name + '/labels%d' % i] = l; arrays[name + '/neutral%d' % i] = np.asarray(e, dtype=bool)
        for c in range(1, kw['n_classes'] + 1):
            arrays[name + '/pred_ids%d' % c] = np.asarray([p[0] for p in preds[c]])
            arrays[name + '/pred%d' % c] = np.asarray([p[1:] for p in preds[c]], dtype=np.float64).reshape(-1, 5)
        arrays[name + '/cfg'] = np.asarray([kw['n_images'], kw['n_classes']])
    np.savez_compressed(os.path.join(HERE, 'ref_eval_golden.npz'), **arrays)
    print('wrote %d arrays' % len(arrays))


if __name__ == '__main__':
    main()
