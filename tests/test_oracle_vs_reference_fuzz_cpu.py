"""Differential fuzzing of the NumPy half of the oracle against the REAL reference, beyond the fixed golden vectors: random
encoder configurations and ground truth, random prediction tensors for the two NumPy decoders, random box sets for iou /
convert_coordinates.  Needs the reference checkout (build container: /root/reference); skipped where it is absent (GPU boxes),
nothing from it is copied or stored.  Bit-exact float64 equality is required, as in tests/test_oracle_golden.py."""
import os
import sys

import numpy as np
import pytest

REF = os.environ.get('SSD_REFERENCE_ROOT', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'ssd_encoder_decoder')), reason='reference checkout not present')


@pytest.fixture(scope='module')
def ref():
    np.float = float    # noqa  the reference targets NumPy < 1.24 (caller-side aliases)
    np.int = int        # noqa
    sys.path.insert(0, REF)
    try:
        from bounding_box_utils.bounding_box_utils import convert_coordinates, convert_coordinates2, iou
        from ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
        from ssd_encoder_decoder.ssd_output_decoder import decode_detections, decode_detections_fast
        yield dict(convert_coordinates=convert_coordinates, convert_coordinates2=convert_coordinates2, iou=iou, SSDInputEncoder=SSDInputEncoder,
                   decode_detections=decode_detections, decode_detections_fast=decode_detections_fast)
    finally:
        sys.path.remove(REF)
        for alias in ('float', 'int'):
            if alias in vars(np):
                delattr(np, alias)
        for k in [m for m in sys.modules if m.startswith(('bounding_box_utils', 'ssd_encoder_decoder')) and not m.startswith('ssd_keras_b200')]:
            sys.modules.pop(k, None)


def _random_encoder_cfg(rng):
    n_layers = int(rng.integers(1, 4))
    H, W = int(rng.integers(60, 200)), int(rng.integers(60, 200))
    sizes = [(int(rng.integers(1, 7)), int(rng.integers(1, 7))) for _ in range(n_layers)]
    scales = sorted(rng.uniform(0.05, 1.0, n_layers + 1).tolist())
    per_layer = bool(rng.integers(0, 2))
    pool = [0.5, 1.0, 2.0, 3.0, 1.0 / 3.0, 1.5]
    ars = [list(rng.choice(pool, size=int(rng.integers(1, 5)), replace=False)) for _ in range(n_layers)]
    cfg = dict(img_height=H, img_width=W, n_classes=int(rng.integers(1, 6)), predictor_sizes=sizes, scales=scales,
               aspect_ratios_global=ars[0] if not per_layer else None, aspect_ratios_per_layer=ars if per_layer else None,
               two_boxes_for_ar1=bool(rng.integers(0, 2)), clip_boxes=bool(rng.integers(0, 2)),
               variances=rng.choice([0.1, 0.2, 1.0], size=4).tolist(), matching_type=str(rng.choice(['multi', 'bipartite'])),
               pos_iou_threshold=float(rng.choice([0.3, 0.5, 0.7])), neg_iou_limit=float(rng.choice([0.2, 0.3, 0.5])),
               border_pixels=str(rng.choice(['half', 'include', 'exclude'])), coords=str(rng.choice(['centroids', 'minmax', 'corners'])),
               normalize_coords=bool(rng.integers(0, 2)))
    cfg['neg_iou_limit'] = min(cfg['neg_iou_limit'], cfg['pos_iou_threshold'])
    if rng.integers(0, 2):
        cfg['steps'] = [(float(rng.uniform(8, 40)), float(rng.uniform(8, 40))) if rng.integers(0, 2) else float(rng.uniform(8, 40))
                        for _ in range(n_layers)]
    if rng.integers(0, 2):
        cfg['offsets'] = [float(rng.uniform(0.2, 0.8)) for _ in range(n_layers)]
    cfg['background_id'] = int(rng.integers(0, cfg['n_classes'] + 1)) if rng.integers(0, 3) == 0 else 0
    return cfg


def _random_gt(rng, cfg, B):
    out = []
    for _ in range(B):
        G = int(rng.integers(0, 6))
        x0 = rng.uniform(0, 0.7 * cfg['img_width'], G); y0 = rng.uniform(0, 0.7 * cfg['img_height'], G)
        w = rng.uniform(5, 0.6 * cfg['img_width'], G); h = rng.uniform(5, 0.6 * cfg['img_height'], G)
        ids = [c for c in range(cfg['n_classes'] + 1) if c != cfg['background_id']]
        g = np.stack([rng.choice(ids, G) if G else np.zeros(0), x0, y0, np.minimum(x0 + w, cfg['img_width'] - 1),
                      np.minimum(y0 + h, cfg['img_height'] - 1)], axis=1) if G else np.zeros((0, 5))
        if G >= 2 and rng.integers(0, 3) == 0:
            g[1] = g[0]                                    # duplicate box: exercises the tie rules
        out.append(g.astype(np.float64))
    return out


@pytest.mark.parametrize('seed', range(40))
def test_encoder_fuzz(ref, seed):
    from oracle.encoder import OracleEncoder
    rng = np.random.default_rng(1000 + seed)
    cfg = _random_encoder_cfg(rng)
    gt = _random_gt(rng, cfg, int(rng.integers(1, 4)))
    r = ref['SSDInputEncoder'](**cfg)
    o = OracleEncoder(**cfg)
    np.testing.assert_array_equal(o.anchors, r.generate_encoding_template(1)[0][:, -8:-4])
    np.testing.assert_array_equal(o(gt), r(gt))


@pytest.mark.parametrize('seed', range(25))
def test_numpy_decoders_fuzz(ref, seed):
    from oracle import synth
    from oracle.decoder import decode_detections, decode_detections_fast
    rng = np.random.default_rng(2000 + seed)
    P, C, B = int(rng.integers(20, 200)), int(rng.integers(2, 7)), int(rng.integers(1, 3))
    anchors = np.concatenate([rng.uniform(0.1, 0.9, (P, 2)), rng.uniform(0.05, 0.5, (P, 2))], axis=1)
    y = synth.synth_y_pred(seed, B, anchors, C, sharp=float(rng.uniform(1, 5)), loc_scale=float(rng.uniform(0.3, 1.5)))
    kw = dict(confidence_thresh=float(rng.choice([0.01, 0.2, 0.5])), iou_threshold=float(rng.choice([0.3, 0.45, 0.6])),
              top_k=int(rng.choice([5, 20, 200])), normalize_coords=bool(rng.integers(0, 2)), img_height=120, img_width=160,
              border_pixels=str(rng.choice(['half', 'include', 'exclude'])))
    for name, fn in (('decode_detections', decode_detections), ('decode_detections_fast', decode_detections_fast)):
        got, want = fn(y, **kw), ref[name](y, **kw)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            a, b = np.asarray(a, np.float64).reshape(-1, 6), np.asarray(b, np.float64).reshape(-1, 6)
            assert a.shape == b.shape
            # top-k of the reference is an unordered argpartition set: compare as sorted rows
            ka = np.lexsort(a.T[::-1]); kb = np.lexsort(b.T[::-1])
            np.testing.assert_array_equal(a[ka], b[kb])


@pytest.mark.parametrize('seed', range(10))
def test_box_math_fuzz(ref, seed):
    from oracle.boxes import convert_coordinates, iou
    from ssd_keras_b200.bounding_box_utils.bounding_box_utils import convert_coordinates as mirror_cc
    rng = np.random.default_rng(3000 + seed)
    m, n = int(rng.integers(1, 9)), int(rng.integers(1, 9))

    def boxes(k):
        xy = rng.uniform(0, 80, (k, 2)); wh = rng.uniform(0, 40, (k, 2))
        return np.concatenate([xy, xy + wh], axis=1)
    b1, b2 = boxes(m), boxes(n)
    from ssd_keras_b200.bounding_box_utils.bounding_box_utils import convert_coordinates2 as mirror_cc2
    wide = np.concatenate([rng.standard_normal((m, 2)), b1], axis=1)          # conversion in the middle of a wider row
    for conv in ('minmax2centroids', 'centroids2minmax'):
        np.testing.assert_array_equal(mirror_cc2(wide, 2, conv), ref['convert_coordinates2'](wide, 2, conv))
    for border in ('half', 'include', 'exclude'):
        for conv in ('minmax2centroids', 'centroids2minmax', 'corners2centroids', 'centroids2corners', 'minmax2corners', 'corners2minmax'):
            np.testing.assert_array_equal(convert_coordinates(b1, 0, conv, border), ref['convert_coordinates'](b1, 0, conv, border))
            np.testing.assert_array_equal(mirror_cc(b1, 0, conv, border), ref['convert_coordinates'](b1, 0, conv, border))   # product (host side)
        for coords in ('corners', 'minmax', 'centroids'):
            c1 = b1 if coords == 'corners' else ref['convert_coordinates'](b1, 0, 'corners2' + coords)
            c2 = b2 if coords == 'corners' else ref['convert_coordinates'](b2, 0, 'corners2' + coords)
            np.testing.assert_array_equal(iou(c1, c2, coords, 'outer_product', border), ref['iou'](c1, c2, coords, 'outer_product', border))
            k = min(m, n)
            np.testing.assert_array_equal(iou(c1[:k], c2[:k], coords, 'element-wise', border),
                                          ref['iou'](c1[:k], c2[:k], coords, 'element-wise', border))


@pytest.mark.parametrize('seed', range(40))
def test_product_anchor_generation_fuzz(ref, seed):
    """PRODUCT code: the library's host-side anchor generator (`ssdk_anchors_generate`, csrc/api.cu) behind the mirror's
    SSDInputEncoder constructor against the real reference on the same random configurations (no GPU needed)."""
    from ssd_keras_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
    rng = np.random.default_rng(1000 + seed)
    cfg = _random_encoder_cfg(rng)
    r = ref['SSDInputEncoder'](**cfg)
    m = SSDInputEncoder(**cfg)
    np.testing.assert_array_equal(m.anchors, r.generate_encoding_template(1)[0][:, -8:-4])
    tpl = m.generate_encoding_template(2)
    np.testing.assert_array_equal(tpl, r.generate_encoding_template(2))
    for a, b in zip(m.boxes_list, r.boxes_list):
        np.testing.assert_array_equal(a, b)
