#!/usr/bin/env python
"""Public signatures (parameter names, order, defaults) of the reference's hot-path entry points, captured from the REAL
reference into tests/golden/ref_signatures.json; tests/test_signature_parity_cpu.py checks that the mirror package accepts the
same parameters in the same order with the same defaults (extra trailing parameters are allowed).  Build container only."""
import importlib
import inspect
import json
import os
import sys

import numpy as np

np.float = float   # noqa
np.int = int       # noqa
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get('SSD_REFERENCE_ROOT', '/root/reference'))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

tf_shim.install()

ENTRY_POINTS = [
    ('ssd_encoder_decoder.ssd_input_encoder', 'SSDInputEncoder.__init__'), ('ssd_encoder_decoder.ssd_input_encoder', 'SSDInputEncoder.__call__'),
    ('ssd_encoder_decoder.ssd_output_decoder', 'decode_detections'), ('ssd_encoder_decoder.ssd_output_decoder', 'decode_detections_fast'),
    ('bounding_box_utils.bounding_box_utils', 'iou'), ('bounding_box_utils.bounding_box_utils', 'convert_coordinates'),
    ('keras_loss_function.keras_ssd_loss', 'SSDLoss.__init__'), ('keras_loss_function.keras_ssd_loss', 'SSDLoss.compute_loss'),
    ('keras_layers.keras_layer_DecodeDetections', 'DecodeDetections.__init__'),
    ('keras_layers.keras_layer_DecodeDetectionsFast', 'DecodeDetectionsFast.__init__'),
    ('keras_layers.keras_layer_AnchorBoxes', 'AnchorBoxes.__init__'), ('keras_layers.keras_layer_L2Normalization', 'L2Normalization.__init__'),
    ('ssd_encoder_decoder.matching_utils', 'match_bipartite_greedy'), ('ssd_encoder_decoder.matching_utils', 'match_multi'),
    ('bounding_box_utils.bounding_box_utils', 'intersection_area'), ('bounding_box_utils.bounding_box_utils', 'intersection_area_'),
    ('ssd_encoder_decoder.ssd_input_encoder', 'SSDInputEncoder.generate_anchor_boxes_for_layer'),
    ('ssd_encoder_decoder.ssd_input_encoder', 'SSDInputEncoder.generate_encoding_template'),
    ('keras_loss_function.keras_ssd_loss', 'SSDLoss.smooth_L1_loss'), ('keras_loss_function.keras_ssd_loss', 'SSDLoss.log_loss'),
    ('models.keras_ssd300', 'ssd_300'), ('models.keras_ssd512', 'ssd_512'), ('models.keras_ssd7', 'build_model'),
]


def signature(mod, path):
    o = importlib.import_module(mod)
    for p in path.split('.'):
        o = getattr(o, p)
    return [[n, None if p.default is inspect.Parameter.empty else repr(p.default)]
            for n, p in inspect.signature(o).parameters.items() if p.kind != inspect.Parameter.VAR_KEYWORD]


if __name__ == '__main__':
    out = {'%s:%s' % (m, p): signature(m, p) for m, p in ENTRY_POINTS}
    with open(os.path.join(HERE, 'ref_signatures.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print('wrote %d signatures' % len(out))
