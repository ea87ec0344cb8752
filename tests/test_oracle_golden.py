"""Pin the oracle restatement against fixtures produced by the REAL reference (make_golden.py)."""
import hashlib

import numpy as np
import pytest

from oracle import synth
from oracle.boxes import convert_coordinates, iou
from oracle.decoder import decode_detections, decode_detections_fast
from oracle.encoder import DegenerateBoxError, OracleEncoder
from oracle.matching import match_bipartite_greedy, match_multi


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_convert_coordinates(golden):
    arr, _ = golden
    b = arr['cc/input']
    for conv in ('minmax2centroids', 'centroids2minmax', 'corners2centroids', 'centroids2corners',
                 'minmax2corners', 'corners2minmax'):
        for bp in ('half', 'include', 'exclude'):
            np.testing.assert_array_equal(convert_coordinates(b, 0, conv, bp), arr['cc/%s/%s' % (conv, bp)])
    np.testing.assert_array_equal(convert_coordinates(b.astype(np.float32), 0, 'centroids2corners'),
                                  arr['cc32/centroids2corners'])
    with pytest.raises(ValueError):
        convert_coordinates(b, 0, 'nope')


def test_iou(golden):
    arr, meta = golden
    b1, b2 = arr['iou/b1'], arr['iou/b2']
    for bp in ('half', 'include', 'exclude'):
        np.testing.assert_array_equal(iou(b1, b2, coords='corners', border_pixels=bp), arr['iou/outer/corners/' + bp])
        np.testing.assert_array_equal(iou(b2, b1[0], coords='corners', mode='element-wise', border_pixels=bp),
                                      arr['iou/elem/corners/' + bp])
    c1 = convert_coordinates(b1, 0, 'corners2centroids'); c2 = convert_coordinates(b2, 0, 'corners2centroids')
    np.testing.assert_array_equal(iou(c1, c2, coords='centroids'), arr['iou/outer/centroids'])
    m1 = convert_coordinates(b1, 0, 'corners2minmax'); m2 = convert_coordinates(b2, 0, 'corners2minmax')
    np.testing.assert_array_equal(iou(m1, m2, coords='minmax'), arr['iou/outer/minmax'])
    # SURVEY 8c known answers (the 'include' value shows the border_pixels quirk)
    k = meta['iou_known']
    assert abs(k['half'] - 0.142857) < 1e-6 and abs(k['include'] - 0.115207) < 1e-6 and abs(k['exclude'] - 0.182482) < 1e-6
    for bp in k:
        v = iou(np.array([0., 0, 10, 10]), np.array([5., 5, 15, 15]), coords='corners', mode='element-wise', border_pixels=bp)[0]
        assert v == k[bp]


def test_matching(golden):
    arr, _ = golden
    i = 0
    while 'match/%d/w' % i in arr:
        w = arr['match/%d/w' % i]
        np.testing.assert_array_equal(match_bipartite_greedy(w), arr['match/%d/bip' % i])
        g, a = match_multi(w, 0.5)
        np.testing.assert_array_equal(g, arr['match/%d/multi_g' % i])
        np.testing.assert_array_equal(a, arr['match/%d/multi_a' % i])
        i += 1
    assert i >= 6
    np.testing.assert_array_equal(match_bipartite_greedy(np.array([[.1, .9], [0, 0]])), [0, 0])


@pytest.mark.parametrize('name', ['ssd300', 'ssd512', 'ssd7', 'micro', 'tiny', 'tiny_clip_abs'])
def test_anchors(golden, configs, name):
    arr, meta = golden
    enc = OracleEncoder(**configs[name])
    m = meta['anchors/' + name]
    anc = enc.anchors
    assert anc.shape[0] == m['P'] and enc.n_classes + 12 == m['width']
    assert sha16(anc.astype(np.float32)) == m['sha_f32']
    assert abs(anc.sum() - m['sum']) < 1e-9 * abs(m['sum'])
    np.testing.assert_array_equal(anc[0], m['first']); np.testing.assert_array_equal(anc[-1], m['last'])
    if name.startswith('tiny'):
        np.testing.assert_array_equal(anc, arr['anchors/' + name])


def test_anchor_known_answers(golden):
    _, meta = golden
    assert meta['anchors/ssd300']['P'] == 8732 and meta['anchors/ssd300']['sha_f32'] == '327b6cefbab2f6a4'
    assert meta['anchors/ssd512']['P'] == 24564 and meta['anchors/ssd512']['sha_f32'] == '5e6173d806d7d6d2'
    assert meta['anchors/ssd7']['P'] == 7160 and meta['anchors/ssd7']['sha_f32'] == '28a06c7e377a1d34'
    assert meta['anchors/micro']['P'] == 100000


def _gts(arr, key, n):
    return [arr['enc/%s/gt%d' % (key, i)] for i in range(n)]


@pytest.mark.parametrize('key,cfg', [('tiny', 'tiny'), ('tiny_corners', 'tiny'), ('tiny_minmax', 'tiny'),
                                     ('tiny_bip', 'tiny_clip_abs'), ('tiny_bg3', 'tiny'), ('tiny_incl', 'tiny')])
def test_encoder_full_tensor(golden, configs, key, cfg):
    arr, meta = golden
    m = meta['enc/' + key]
    c = dict(configs[cfg]); c.update(m['over'])
    y = OracleEncoder(**c)(_gts(arr, key, m['n_gt']))
    np.testing.assert_array_equal(y, arr['enc/%s/y' % key])


@pytest.mark.parametrize('key,cfg', [('ssd300', 'ssd300'), ('ssd300_neg03', 'ssd300'), ('micro', 'micro')])
def test_encoder_large(golden, configs, key, cfg):
    arr, meta = golden
    m = meta['enc/' + key]
    c = dict(configs[cfg]); c.update(m['over'])
    y = OracleEncoder(**c)(_gts(arr, key, m['n_gt']))
    assert list(y.shape) == m['shape']
    assert sha16(y.astype(np.float32)) == m['sha_f32']
    assert abs(y.sum() - m['sum']) <= 1e-9 * abs(m['sum'])
    np.testing.assert_array_equal(np.argwhere(y[:, :, :-12].sum(axis=-1) == 0), arr['enc/%s/neutral' % key])
    pos = np.argwhere(y[:, :, 1:-12].max(axis=-1) > 0)
    np.testing.assert_array_equal(pos, arr['enc/%s/pos' % key if key != 'ssd300_neg03' else 'enc/ssd300/pos'])


def test_encoder_gt_generator_matches_survey(configs):
    # synth_gt(2, 32, 8, 300, 300, 20) -> 2397 positives, 0 neutral (neg limit 0.5)
    y = OracleEncoder(**configs['ssd300'])(synth.synth_gt(2, 32, 8, 300, 300, 20))
    assert int((y[:, :, 1:-12].max(axis=-1) > 0).sum()) == 2397
    assert int((y[:, :, :-12].sum(axis=-1) == 0).sum()) == 0
    assert abs(y.sum() - 828803.593393) < 1e-3


def test_encoder_degenerate(configs):
    enc = OracleEncoder(**configs['tiny'])
    with pytest.raises(DegenerateBoxError):
        enc([np.array([[1, 10., 10., 10., 20.]])])


def _same_rows(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 6); b = np.asarray(b, dtype=np.float64).reshape(-1, 6)
    assert a.shape == b.shape
    ka = np.lexsort(a.T[::-1]); kb = np.lexsort(b.T[::-1])
    np.testing.assert_array_equal(a[ka], b[kb])


@pytest.mark.parametrize('key,fn', [('tiny', decode_detections), ('tiny_topk', decode_detections),
                                    ('tiny_nonorm', decode_detections), ('tiny_empty', decode_detections),
                                    ('tiny_fast', decode_detections_fast), ('tiny_fast_topk', decode_detections_fast)])
def test_decoders_small(golden, key, fn):
    arr, meta = golden
    m = meta['dec/' + key]
    res = fn(arr['dec/%s/y_pred' % key], **m['kw'])
    assert [int(np.asarray(r).reshape(-1, 6).shape[0]) for r in res] == m['counts']
    for i, r in enumerate(res):
        _same_rows(r, arr['dec/%s/out%d' % (key, i)])     # top-k via argpartition is unordered: compare as sets


def test_decoders_ssd300(golden, configs):
    arr, meta = golden
    enc = OracleEncoder(**configs['ssd300'])
    yp = synth.synth_y_pred(23, 1, enc.anchors, 21, sharp=6.0, loc_scale=1.0)
    res = decode_detections(yp, confidence_thresh=0.5, iou_threshold=0.45, top_k=200, img_height=300, img_width=300)
    assert [r.shape[0] for r in res] == meta['dec/ssd300']['counts']
    _same_rows(res[0], arr['dec/ssd300/out0'])
    res = decode_detections_fast(yp, confidence_thresh=0.5, iou_threshold=0.45, top_k=200, img_height=300, img_width=300)
    assert [r.shape[0] for r in res] == meta['dec/ssd300_fast']['counts']
    _same_rows(res[0], arr['dec/ssd300_fast/out0'])
