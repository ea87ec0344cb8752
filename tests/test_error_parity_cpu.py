"""Invalid-argument behaviour of the mirror package against the REAL reference: same exception type and message
(tests/golden/ref_errors.json, captured by tests/golden/make_errors_golden.py from the reference's own code).  Everything
here fails before any device work, so it runs without a GPU."""
import json
import os

import numpy as np
import pytest

CASES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_errors.json')))


def _target(name):
    if name == 'SSDInputEncoder':
        from ssd_keras_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder as t
    elif name == 'decode_detections':
        from ssd_keras_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections as t
    elif name == 'decode_detections_fast':
        from ssd_keras_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections_fast as t
    elif name == 'DecodeDetections':
        from ssd_keras_b200.keras_layers.keras_layer_DecodeDetections import DecodeDetections as t
    elif name == 'DecodeDetectionsFast':
        from ssd_keras_b200.keras_layers.keras_layer_DecodeDetectionsFast import DecodeDetectionsFast as t
    elif name == 'ssd_300':
        from ssd_keras_b200.models.keras_ssd300 import ssd_300 as t
    elif name == 'ssd_512':
        from ssd_keras_b200.models.keras_ssd512 import ssd_512 as t
    else:
        from ssd_keras_b200.models.keras_ssd7 import build_model as t
    return t


@pytest.mark.parametrize('name', sorted(CASES))
def test_same_exception_as_reference(name):
    c = CASES[name]
    if c['target'] == 'SSDInputEncoder.__call__':               # degenerate ground truth: detected on the host before any device work
        from ssd_keras_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
        with pytest.raises(Exception) as ei:
            SSDInputEncoder(**c['kwargs'])([np.array([[1, 5., 5., 50., 60.]]), np.array([c['args'][0]])])
        assert type(ei.value).__name__ == c['result'][0] and str(ei.value) == c['result'][1]
        return
    args = [np.zeros((1, 10, 15), np.float32) if a == 'zeros(1,10,15)' else (tuple(a) if isinstance(a, list) else a) for a in c['args']]
    kind, msg = c['result']
    fn = _target(c['target'])
    if kind == 'OK':
        fn(*args, **c['kwargs'])                      # construction succeeds in the reference: must succeed here too
        return
    with pytest.raises(Exception) as ei:
        fn(*args, **c['kwargs'])
    assert type(ei.value).__name__ == kind, (type(ei.value).__name__, str(ei.value))
    assert str(ei.value) == msg
