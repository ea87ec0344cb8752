"""The smaller public helpers of the mirror package against outputs of the REAL reference (tests/golden/ref_golden.npz,
ref_api_extra.npz): stand-alone match_bipartite_greedy / match_multi, intersection_area, the per-layer anchor method and the
SSDLoss helper losses.  The tensor routines are device-agnostic; here they run on CPU tensors (the public wrappers refuse to
run without CUDA), the per-layer anchors use the library's host routine."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'ref_golden.npz'))
X = np.load(os.path.join(HERE, 'golden', 'ref_api_extra.npz'))


@pytest.mark.parametrize('i', range(6))
def test_standalone_matching_matches_reference(i):
    from ssd_keras_b200.ssd_encoder_decoder.matching_utils import _bipartite_t, _multi_t
    w = torch.from_numpy(G['match/%d/w' % i].astype(np.float64))
    np.testing.assert_array_equal(_bipartite_t(w).numpy(), G['match/%d/bip' % i])
    g, a = _multi_t(w, 0.5)
    np.testing.assert_array_equal(g.numpy(), G['match/%d/multi_g' % i])
    np.testing.assert_array_equal(a.numpy(), G['match/%d/multi_a' % i])


def test_standalone_matching_needs_cuda():
    from ssd_keras_b200.ssd_encoder_decoder.matching_utils import match_bipartite_greedy
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    with pytest.raises(RuntimeError):
        match_bipartite_greedy(np.zeros((2, 3)))


@pytest.mark.parametrize('coords', ['corners', 'minmax', 'centroids'])
@pytest.mark.parametrize('border', ['half', 'include', 'exclude'])
def test_intersection_area_matches_reference(coords, border):
    from ssd_keras_b200 import _ffi
    from ssd_keras_b200.bounding_box_utils.bounding_box_utils import _intersection_t, convert_coordinates
    b1, b2 = X['inter/b1'], X['inter/b2']
    if coords == 'minmax':
        b1, b2 = convert_coordinates(b1, 0, 'corners2minmax'), convert_coordinates(b2, 0, 'corners2minmax')
    inner = 'minmax' if coords == 'minmax' else 'corners'     # 'centroids' inputs are converted to corners first (:136-139)
    if coords == 'centroids':                                 # round trip through the mirror's own conversion, like the public function
        c1 = convert_coordinates(X['inter/b1'], 0, 'corners2centroids'); c2 = convert_coordinates(X['inter/b2'], 0, 'corners2centroids')
        b1, b2 = convert_coordinates(c1, 0, 'centroids2corners'), convert_coordinates(c2, 0, 'centroids2corners')
    d = float(_ffi.BORDER_D[border])
    t1, t2 = torch.from_numpy(np.ascontiguousarray(b1)), torch.from_numpy(np.ascontiguousarray(b2))
    np.testing.assert_array_equal(_intersection_t(t1, t2, inner, False, d).numpy(), X['inter/%s/%s/outer' % (coords, border)])
    np.testing.assert_array_equal(_intersection_t(t1[:5], t2, inner, True, d).numpy(), X['inter/%s/%s/elem' % (coords, border)])


def test_generate_anchor_boxes_for_layer_matches_reference():
    from ssd_keras_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
    enc = SSDInputEncoder(img_height=120, img_width=160, n_classes=3, predictor_sizes=[(6, 8), (3, 4)], scales=[0.2, 0.45, 0.8],
                          aspect_ratios_per_layer=[[1.0, 2.0], [0.5, 3.0]], two_boxes_for_ar1=True, steps=[20, (40, 41)],
                          offsets=[0.5, (0.4, 0.6)], clip_boxes=True, variances=[0.1, 0.1, 0.2, 0.2], matching_type='bipartite',
                          pos_iou_threshold=0.5, neg_iou_limit=0.2, normalize_coords=False)
    np.testing.assert_array_equal(enc.generate_anchor_boxes_for_layer((5, 7), [1.0, 2.0, 0.5], 0.3, 0.5), X['anchors_layer/a'])
    bx, (cy, cx), wh, step, off = enc.generate_anchor_boxes_for_layer((3, 4), [0.5, 3.0], 0.45, 0.8, this_steps=(40, 41),
                                                                      this_offsets=(0.4, 0.6), diagnostics=True)
    np.testing.assert_array_equal(bx, X['anchors_layer/b'])
    np.testing.assert_array_equal(cy, X['anchors_layer/b_cy']); np.testing.assert_array_equal(cx, X['anchors_layer/b_cx'])
    np.testing.assert_array_equal(wh, X['anchors_layer/b_wh'])
    np.testing.assert_array_equal(np.array(step, dtype=np.float64), X['anchors_layer/b_step'])
    np.testing.assert_array_equal(np.array(off, dtype=np.float64), X['anchors_layer/b_off'])


def test_loss_helpers_match_reference():
    from ssd_keras_b200.keras_loss_function.keras_ssd_loss import SSDLoss
    f = torch.from_numpy
    out = SSDLoss._smooth_l1_t(f(X['loss_helpers/l1_true']), f(X['loss_helpers/l1_pred'])).numpy()
    np.testing.assert_allclose(out, X['loss_helpers/l1_out'], rtol=1e-6, atol=1e-7)
    out = SSDLoss._log_loss_t(f(X['loss_helpers/log_true']), f(X['loss_helpers/log_pred'])).numpy()
    np.testing.assert_allclose(out, X['loss_helpers/log_out'], rtol=1e-6, atol=1e-6)


def test_pred_layer_helpers_and_inverse_transforms():
    """get_num_boxes_per_pred_layer / get_pred_layers (ssd_output_decoder.py:488-530) against the real reference's outputs,
    apply_inverse_transforms (data_generator/object_detection_2d_misc_utils.py:22-73) against its documented behaviour."""
    import os
    from ssd_keras_b200.data_generator.object_detection_2d_misc_utils import apply_inverse_transforms
    from ssd_keras_b200.ssd_encoder_decoder.ssd_output_decoder import get_num_boxes_per_pred_layer, get_pred_layers
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_debug_golden.npz'))
    nb = get_num_boxes_per_pred_layer([(6, 8), (3, 4)], [[0.5, 1.0, 2.0]] * 2, True)
    np.testing.assert_array_equal(nb, G['dbg/num_boxes'])
    np.testing.assert_array_equal(get_num_boxes_per_pred_layer([(38, 38), (19, 19)], [[1.0, 2.0, 0.5], [1.0, 2.0, 0.5, 3.0, 1 / 3]], False),
                                  G['dbg/num_boxes_one'])
    dec = [G['dbg/centroids/a/out%d' % i] for i in range(3)]
    for i, layers in enumerate(get_pred_layers(dec, nb)):
        np.testing.assert_array_equal(layers, G['dbg/centroids/layers%d' % i])
    with pytest.raises(ValueError):
        get_pred_layers([np.array([[1e9, 1, 0.5, 0, 0, 1, 1]])], nb)
    shift = lambda a: a + np.array([0, 0, 10, 20, 10, 20.])       # noqa: E731
    lst = [np.ones((2, 6)), np.zeros((0, 6)), np.ones((1, 6))]
    out = apply_inverse_transforms(lst, [[shift, None], [shift], [shift, shift]])
    np.testing.assert_array_equal(out[0], np.ones((2, 6)) + [0, 0, 10, 20, 10, 20])
    assert out[1].shape == (0, 6)
    np.testing.assert_array_equal(out[2], np.ones((1, 6)) + [0, 0, 20, 40, 20, 40])
    arr = np.ones((2, 3, 6))
    out = apply_inverse_transforms(arr, [[shift], [None]])
    np.testing.assert_array_equal(out[0], arr[0] + [0, 0, 10, 20, 10, 20]); np.testing.assert_array_equal(out[1], arr[1])
    with pytest.raises(ValueError):
        apply_inverse_transforms(3, [])
