"""Cross-pin of the TensorFlow-side decoder restatement (oracle.decoder.decode_layer / decode_layer_fast, which restate
keras_layers/keras_layer_DecodeDetections.py:109-265 and ..Fast.py:111-248 and cannot be run against TensorFlow offline)
against outputs of the REAL reference's NumPy twins of the same algorithms (ssd_output_decoder.py:111-226 / :228-333),
stored in tests/golden/ref_golden.npz by make_golden.py.  The two implementations share the algorithm (per-class or
arg-max-class threshold, greedy IoU NMS, top-k) and differ in arithmetic width (float32 TF ops vs float64 NumPy), the
layer's nms_max_output_size cap and zero padding; on these tie-free inputs with fewer survivors than the cap the kept
detections must be the same set."""
import json
import os

import numpy as np
import pytest

from oracle.decoder import decode_layer, decode_layer_fast

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'ref_golden.npz'))
META = json.load(open(os.path.join(HERE, 'golden', 'ref_golden.json')))


def _rows(a):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 6)
    a = a[np.any(a != 0, axis=1)]                      # the layer pads with all-zero rows
    return a[np.lexsort((a[:, 5], a[:, 4], a[:, 3], a[:, 2], -a[:, 1], a[:, 0]))]


@pytest.mark.parametrize('key,fast', [('tiny', False), ('tiny_topk', False), ('tiny_nonorm', False), ('tiny_empty', False),
                                      ('tiny_fast', True), ('tiny_fast_topk', True)])
def test_layer_oracle_equals_reference_numpy_decoder(key, fast):
    y = G['dec/%s/y_pred' % key]
    kw = dict(META['dec/' + key]['kw'])
    P = y.shape[1]
    top_k = kw.get('top_k', 200)
    top_k = P * (y.shape[2] - 12) if top_k == 'all' else int(top_k)
    fn = decode_layer_fast if fast else decode_layer
    with np.errstate(all='ignore'):
        out = fn(y, confidence_thresh=kw['confidence_thresh'], iou_threshold=kw['iou_threshold'], top_k=top_k,
                 nms_max_output_size=max(400, P), normalize_coords=kw.get('normalize_coords', True),
                 img_height=kw.get('img_height'), img_width=kw.get('img_width'))
    for i in range(META['dec/' + key]['n']):
        ref = _rows(G['dec/%s/out%d' % (key, i)])
        got = _rows(out[i])
        assert got.shape == ref.shape, (key, i, got.shape, ref.shape)
        if ref.size:
            np.testing.assert_array_equal(got[:, 0], ref[:, 0])                       # classes
            np.testing.assert_allclose(got[:, 1], ref[:, 1], rtol=1e-6)               # confidences
            np.testing.assert_allclose(got[:, 2:], ref[:, 2:], rtol=1e-5, atol=1e-3)   # pixel coordinates
