"""GPU tests of the decoder module's debug / stand-alone utilities against outputs of the REAL reference
(tests/golden/make_debug_golden.py): decode_detections_debug incl. variance_encoded_in_target, greedy_nms; host-only helpers
get_num_boxes_per_pred_layer / get_pred_layers / apply_inverse_transforms are in test_api_helpers_cpu.py."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_debug_golden.npz'))


@pytest.fixture(scope='module', autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


def _same_rows(a, b, rtol=1e-6, atol=1e-4):
    a = np.asarray(a, np.float64).reshape(-1, 7); b = np.asarray(b, np.float64).reshape(-1, 7)
    assert a.shape == b.shape, (a.shape, b.shape)
    ka = np.lexsort((a[:, 0], a[:, 1], -a[:, 2])); kb = np.lexsort((b[:, 0], b[:, 1], -b[:, 2]))
    a, b = a[ka], b[kb]
    np.testing.assert_array_equal(a[:, :2], b[:, :2])                 # prior index and class id: bit-exact
    np.testing.assert_allclose(a[:, 2:], b[:, 2:], rtol=rtol, atol=atol)


@pytest.mark.parametrize('coords', ['centroids', 'corners', 'minmax'])
@pytest.mark.parametrize('tag', ['a', 'topk', 'vit'])
def test_decode_detections_debug_vs_reference(coords, tag):
    from ssd_keras_b200.ssd_encoder_decoder.ssd_output_decoder import (decode_detections_debug, get_num_boxes_per_pred_layer,
                                                                        get_pred_layers)
    kw = dict(a=dict(top_k=200), topk=dict(top_k=7), vit=dict(top_k=200, variance_encoded_in_target=True))[tag]
    yp = G['dbg/%s/y_pred' % coords]
    res = decode_detections_debug(yp, confidence_thresh=0.05, iou_threshold=0.45, input_coords=coords, normalize_coords=True,
                                  img_height=120, img_width=160, **kw)
    assert len(res) == 3
    for i, r in enumerate(res):
        ref = G['dbg/%s/%s/out%d' % (coords, tag, i)]
        if tag == 'topk':                                             # argpartition keeps an arbitrary order: compare as sets
            assert r.shape == ref.shape
        _same_rows(r, ref)
    if tag == 'a':
        nb = get_num_boxes_per_pred_layer([(6, 8), (3, 4)], [[0.5, 1.0, 2.0]] * 2, True)
        for i, (r, layers) in enumerate(zip(res, get_pred_layers(res, nb))):
            # same multiset of layer indices per image
            np.testing.assert_array_equal(np.sort(layers), np.sort(G['dbg/%s/layers%d' % (coords, i)]))


@pytest.mark.parametrize('bp', ['half', 'include', 'exclude'])
@pytest.mark.parametrize('thr', [0.45, 0.1])
def test_greedy_nms_vs_reference(bp, thr):
    from ssd_keras_b200.ssd_encoder_decoder.ssd_output_decoder import greedy_nms
    items = [G['nms/in%d' % i] for i in range(3)]
    res = greedy_nms(items, iou_threshold=thr, coords='corners', border_pixels=bp)
    for i, r in enumerate(res):
        np.testing.assert_array_equal(r, G['nms/%s/%g/out%d' % (bp, thr, i)])      # same survivors, same order, same bits
