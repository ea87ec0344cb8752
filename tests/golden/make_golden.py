#!/usr/bin/env python
"""Generate the committed golden fixtures by running the REAL reference (read-only, /root/reference).

Run in the build container only (``python tests/golden/make_golden.py``); the GPU box has no
/root/reference and only ever reads the committed ``ref_golden.npz`` / ``ref_golden.json``.

The reference targets NumPy < 1.24 (``np.float`` / ``np.int``); the two aliases are restored in
THIS process before importing it -- the reference tree itself is untouched.
"""
import hashlib
import json
import os
import sys

import numpy as np

np.float = float   # noqa  (caller-side shim)
np.int = int       # noqa

REF = os.environ.get('SSD_REFERENCE_ROOT', '/root/reference')
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from bounding_box_utils.bounding_box_utils import convert_coordinates, iou                    # noqa: E402
from ssd_encoder_decoder.matching_utils import match_bipartite_greedy, match_multi            # noqa: E402
from ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder                             # noqa: E402
from ssd_encoder_decoder.ssd_output_decoder import decode_detections, decode_detections_fast  # noqa: E402

from oracle import synth                                                                       # noqa: E402
from oracle.model import SSD300_AR, SSD512_AR                                                  # noqa: E402

CONFIGS = {
    'ssd300': dict(img_height=300, img_width=300, n_classes=20,
                   predictor_sizes=[(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)],
                   scales=[0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05], aspect_ratios_per_layer=SSD300_AR,
                   two_boxes_for_ar1=True, steps=[8, 16, 32, 64, 100, 300], offsets=[0.5] * 6, clip_boxes=False,
                   variances=[0.1, 0.1, 0.2, 0.2], matching_type='multi', pos_iou_threshold=0.5,
                   neg_iou_limit=0.5, normalize_coords=True),
    'ssd512': dict(img_height=512, img_width=512, n_classes=80,
                   predictor_sizes=[(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2), (1, 1)],
                   scales=[0.04, 0.1, 0.26, 0.42, 0.58, 0.74, 0.9, 1.06], aspect_ratios_per_layer=SSD512_AR,
                   two_boxes_for_ar1=True, steps=[8, 16, 32, 64, 128, 256, 512], offsets=[0.5] * 7,
                   clip_boxes=False, variances=[0.1, 0.1, 0.2, 0.2], matching_type='multi',
                   pos_iou_threshold=0.5, neg_iou_limit=0.3, normalize_coords=True),
    'ssd7': dict(img_height=300, img_width=300, n_classes=5,
                 predictor_sizes=[(37, 37), (18, 18), (9, 9), (4, 4)],
                 scales=[0.08, 0.16, 0.32, 0.64, 0.96], aspect_ratios_global=[0.5, 1.0, 2.0],
                 two_boxes_for_ar1=True, steps=None, offsets=None, clip_boxes=False,
                 variances=[1.0, 1.0, 1.0, 1.0], matching_type='multi', pos_iou_threshold=0.5,
                 neg_iou_limit=0.3, normalize_coords=True),
    'micro': dict(img_height=1000, img_width=1600, n_classes=20, predictor_sizes=[(125, 200)],
                  scales=[0.1, 0.2], aspect_ratios_global=[0.5, 1.0, 2.0], variances=[0.1, 0.1, 0.2, 0.2],
                  matching_type='multi', pos_iou_threshold=0.5, neg_iou_limit=0.5),
    # small configs whose full tensors are stored
    'tiny': dict(img_height=120, img_width=160, n_classes=3, predictor_sizes=[(6, 8), (3, 4)],
                 scales=[0.2, 0.45, 0.8], aspect_ratios_global=[0.5, 1.0, 2.0], two_boxes_for_ar1=True,
                 variances=[0.1, 0.1, 0.2, 0.2], matching_type='multi', pos_iou_threshold=0.5,
                 neg_iou_limit=0.3, normalize_coords=True),
    'tiny_clip_abs': dict(img_height=120, img_width=160, n_classes=3, predictor_sizes=[(6, 8), (3, 4)],
                          scales=[0.2, 0.45, 0.8], aspect_ratios_per_layer=[[1.0, 2.0], [0.5, 3.0]],
                          two_boxes_for_ar1=True, steps=[20, (40, 41)], offsets=[0.5, (0.4, 0.6)],
                          clip_boxes=True, variances=[0.1, 0.1, 0.2, 0.2], matching_type='bipartite',
                          pos_iou_threshold=0.5, neg_iou_limit=0.2, normalize_coords=False),
}


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    arrays, meta = {}, {}
    rng = np.random.default_rng(1234)

    # ---- box math ---------------------------------------------------------------------------
    b = rng.uniform(0, 100, size=(7, 4))
    b[:, 2:] += b[:, :2] + 1
    for conv in ('minmax2centroids', 'centroids2minmax', 'corners2centroids', 'centroids2corners',
                 'minmax2corners', 'corners2minmax'):
        for bp in ('half', 'include', 'exclude'):
            arrays['cc/%s/%s' % (conv, bp)] = convert_coordinates(b, 0, conv, bp)
    arrays['cc/input'] = b
    b32 = b.astype(np.float32)
    arrays['cc32/centroids2corners'] = convert_coordinates(b32, 0, 'centroids2corners')
    b1 = rng.uniform(0, 50, size=(5, 4)); b1[:, 2:] += b1[:, :2]
    b2 = rng.uniform(0, 50, size=(9, 4)); b2[:, 2:] += b2[:, :2]
    arrays['iou/b1'], arrays['iou/b2'] = b1, b2
    for bp in ('half', 'include', 'exclude'):
        arrays['iou/outer/corners/' + bp] = iou(b1, b2, coords='corners', mode='outer_product', border_pixels=bp)
        arrays['iou/elem/corners/' + bp] = iou(b2, b1[0], coords='corners', mode='element-wise', border_pixels=bp)
    c1 = convert_coordinates(b1, 0, 'corners2centroids'); c2 = convert_coordinates(b2, 0, 'corners2centroids')
    arrays['iou/outer/centroids'] = iou(c1, c2, coords='centroids')
    m1 = convert_coordinates(b1, 0, 'corners2minmax'); m2 = convert_coordinates(b2, 0, 'corners2minmax')
    arrays['iou/outer/minmax'] = iou(m1, m2, coords='minmax')
    meta['iou_known'] = {bp: float(iou(np.array([0., 0, 10, 10]), np.array([5., 5, 15, 15]), coords='corners',
                                       mode='element-wise', border_pixels=bp)[0]) for bp in ('half', 'include', 'exclude')}

    # ---- matching ---------------------------------------------------------------------------
    mats = [np.array([[.1, .9], [0, 0]]), np.array([[.6, .2, .5], [.6, .7, .5]]),
            rng.uniform(0, 1, (6, 40)), np.round(rng.uniform(0, 1, (5, 30)), 1),
            np.zeros((3, 10)), (rng.uniform(0, 1, (8, 50)) > 0.8) * rng.uniform(0, 1, (8, 50))]
    for i, m in enumerate(mats):
        arrays['match/%d/w' % i] = m
        arrays['match/%d/bip' % i] = match_bipartite_greedy(m)
        g, a = match_multi(m, 0.5)
        arrays['match/%d/multi_g' % i], arrays['match/%d/multi_a' % i] = g, a

    # ---- anchors ----------------------------------------------------------------------------
    for name, cfg in CONFIGS.items():
        enc = SSDInputEncoder(**cfg)
        tmpl = enc.generate_encoding_template(1)[0]
        anc = tmpl[:, -8:-4]
        meta['anchors/' + name] = dict(P=int(anc.shape[0]), width=int(tmpl.shape[1]), sum=float(anc.sum()),
                                       sha_f32=sha16(anc.astype(np.float32)), first=anc[0].tolist(),
                                       last=anc[-1].tolist(), n_boxes=enc.n_boxes)
        if name.startswith('tiny'):
            arrays['anchors/' + name] = anc

    # ---- encoder ----------------------------------------------------------------------------
    def enc_case(key, cfg, gts, **over):
        c = dict(cfg); c.update(over)
        enc = SSDInputEncoder(**c)
        y = enc(gts)
        for i, g in enumerate(gts):
            arrays['enc/%s/gt%d' % (key, i)] = np.asarray(g, dtype=np.float32)
        meta['enc/' + key] = dict(n_gt=len(gts), over={k: v for k, v in over.items()}, shape=list(y.shape),
                                  sum=float(y.sum()), abs_off=float(np.abs(y[:, :, -12:-8]).sum()),
                                  sha_f32=sha16(y.astype(np.float32)))
        return enc, y

    tiny_gt = synth.synth_gt(11, 5, 4, 160, 120, 3)
    tiny_gt[2] = np.zeros((0, 5), np.float32)                          # empty image
    tiny_gt[3] = np.concatenate([tiny_gt[3], tiny_gt[3][:1]], axis=0)   # duplicate gt box
    tiny_gt[4] = np.array([[1, 2., 2., 9., 9.], [2, 150., 100., 159., 119.]], np.float32)  # tiny boxes: zero-IoU rows
    _, y = enc_case('tiny', CONFIGS['tiny'], tiny_gt)
    arrays['enc/tiny/y'] = y
    for co in ('corners', 'minmax'):
        _, y = enc_case('tiny_' + co, CONFIGS['tiny'], tiny_gt, coords=co)
        arrays['enc/tiny_%s/y' % co] = y
    _, y = enc_case('tiny_bip', CONFIGS['tiny_clip_abs'], tiny_gt)
    arrays['enc/tiny_bip/y'] = y
    _, y = enc_case('tiny_bg3', CONFIGS['tiny'], tiny_gt, background_id=3, neg_iou_limit=0.1)
    arrays['enc/tiny_bg3/y'] = y
    _, y = enc_case('tiny_incl', CONFIGS['tiny'], tiny_gt, border_pixels='include', normalize_coords=False)
    arrays['enc/tiny_incl/y'] = y

    gt300 = synth.synth_gt(2, 4, 8, 300, 300, 20)
    _, y = enc_case('ssd300', CONFIGS['ssd300'], gt300)
    pos = np.argwhere(y[:, :, 1:-12].max(axis=-1) > 0)
    arrays['enc/ssd300/pos'] = pos.astype(np.int32)
    arrays['enc/ssd300/pos_rows'] = y[pos[:, 0], pos[:, 1]].astype(np.float32)
    arrays['enc/ssd300/neutral'] = np.argwhere(y[:, :, :-12].sum(axis=-1) == 0).astype(np.int32)
    _, y = enc_case('ssd300_neg03', CONFIGS['ssd300'], gt300, neg_iou_limit=0.3)
    arrays['enc/ssd300_neg03/neutral'] = np.argwhere(y[:, :, :-12].sum(axis=-1) == 0).astype(np.int32)
    gtm = synth.synth_gt(4, 1, 128, 1600, 1000, 20)
    _, y = enc_case('micro', CONFIGS['micro'], gtm)
    pos = np.argwhere(y[:, :, 1:-12].max(axis=-1) > 0)
    arrays['enc/micro/pos'] = pos.astype(np.int32)
    arrays['enc/micro/pos_cls'] = y[pos[:, 0], pos[:, 1], :-12].argmax(-1).astype(np.int32)
    arrays['enc/micro/pos_off'] = y[pos[:, 0], pos[:, 1], -12:-8].astype(np.float32)
    arrays['enc/micro/neutral'] = np.argwhere(y[:, :, :-12].sum(axis=-1) == 0).astype(np.int32)

    # ---- decoders ---------------------------------------------------------------------------
    def dec_case(key, y_pred, fn, **kw):
        res = fn(y_pred, **kw)
        arrays['dec/%s/y_pred' % key] = y_pred
        meta['dec/' + key] = dict(n=len(res), kw={k: (v if not isinstance(v, np.generic) else v.item()) for k, v in kw.items()},
                                  counts=[int(r.shape[0]) for r in res])
        for i, r in enumerate(res):
            arrays['dec/%s/out%d' % (key, i)] = np.asarray(r, dtype=np.float64).reshape(-1, 6)

    enc_t = SSDInputEncoder(**CONFIGS['tiny'])
    anc_t = enc_t.generate_encoding_template(1)[0][:, -8:-4]
    yp = synth.synth_y_pred(21, 3, anc_t, 4, sharp=3.0, loc_scale=1.0)
    dec_case('tiny', yp, decode_detections, confidence_thresh=0.01, iou_threshold=0.45, top_k=200,
             img_height=120, img_width=160)
    dec_case('tiny_topk', yp, decode_detections, confidence_thresh=0.2, iou_threshold=0.3, top_k=10,
             img_height=120, img_width=160)
    dec_case('tiny_fast', yp, decode_detections_fast, confidence_thresh=0.5, iou_threshold=0.45, top_k='all',
             img_height=120, img_width=160)
    dec_case('tiny_fast_topk', yp, decode_detections_fast, confidence_thresh=0.3, iou_threshold=0.45, top_k=7,
             img_height=120, img_width=160)
    dec_case('tiny_nonorm', yp, decode_detections, confidence_thresh=0.3, iou_threshold=0.45, top_k=200,
             normalize_coords=False)
    yp_hi = synth.synth_y_pred(22, 2, anc_t, 4, sharp=0.2)
    dec_case('tiny_empty', yp_hi, decode_detections, confidence_thresh=0.9, iou_threshold=0.45, top_k=200,
             img_height=120, img_width=160)
    enc3 = SSDInputEncoder(**CONFIGS['ssd300'])
    anc3 = enc3.generate_encoding_template(1)[0][:, -8:-4]
    yp3 = synth.synth_y_pred(23, 1, anc3, 21, sharp=6.0, loc_scale=1.0)
    res = decode_detections(yp3, confidence_thresh=0.5, iou_threshold=0.45, top_k=200, img_height=300, img_width=300)
    meta['dec/ssd300'] = dict(seed=23, sharp=6.0, counts=[int(r.shape[0]) for r in res])
    arrays['dec/ssd300/out0'] = res[0]
    res = decode_detections_fast(yp3, confidence_thresh=0.5, iou_threshold=0.45, top_k=200, img_height=300, img_width=300)
    meta['dec/ssd300_fast'] = dict(seed=23, sharp=6.0, counts=[int(r.shape[0]) for r in res])
    arrays['dec/ssd300_fast/out0'] = res[0]

    np.savez_compressed(os.path.join(HERE, 'ref_golden.npz'), **arrays)
    with open(os.path.join(HERE, 'ref_golden.json'), 'w') as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print('wrote %d arrays, %d meta entries' % (len(arrays), len(meta)))


if __name__ == '__main__':
    main()
