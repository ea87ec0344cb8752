#!/usr/bin/env python
"""Error behaviour of the reference's public entry points for invalid arguments (exception type + message), captured from
the REAL reference (NumPy half imported directly, Keras half over tests/golden/tf_shim.py) into tests/golden/ref_errors.json.
tests/test_error_parity_cpu.py replays the same calls against the mirror package.  Build container only."""
import json
import os
import sys

import numpy as np

np.float = float   # noqa
np.int = int       # noqa
REF = os.environ.get('SSD_REFERENCE_ROOT', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

tf_shim.install()
tf_shim.STATE['input'] = np.zeros((1, 300, 300, 3), np.float32)
tf_shim.STATE['weights'] = {}

from keras_layers.keras_layer_DecodeDetections import DecodeDetections                                  # noqa: E402
from keras_layers.keras_layer_DecodeDetectionsFast import DecodeDetectionsFast                          # noqa: E402
from models.keras_ssd300 import ssd_300                                                                  # noqa: E402
from models.keras_ssd512 import ssd_512                                                                  # noqa: E402
from models.keras_ssd7 import build_model                                                                # noqa: E402
from ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder                                        # noqa: E402
from ssd_encoder_decoder.ssd_output_decoder import decode_detections, decode_detections_fast            # noqa: E402

ENC = dict(img_height=120, img_width=160, n_classes=3, predictor_sizes=[(6, 8), (3, 4)], scales=[0.2, 0.45, 0.8],
           aspect_ratios_global=[0.5, 1.0, 2.0])
SC300 = [.1, .2, .37, .54, .71, .88, 1.05]
SC512 = [.04, .1, .26, .42, .58, .74, .9, 1.06]
Y = 'zeros(1,10,15)'        # placeholder understood by the test: np.zeros((1, 10, 15), float32)

# (target, positional args, keyword overrides)
CASES = {
    'enc_scales_len': ('SSDInputEncoder', [], dict(ENC, scales=[0.2, 0.45])),
    'enc_no_scales': ('SSDInputEncoder', [], dict(ENC, scales=None, min_scale=None, max_scale=None)),
    'enc_ar_len': ('SSDInputEncoder', [], dict(ENC, aspect_ratios_per_layer=[[1.0]])),
    'enc_no_ar': ('SSDInputEncoder', [], dict(ENC, aspect_ratios_global=None, aspect_ratios_per_layer=None)),
    'enc_steps_len': ('SSDInputEncoder', [], dict(ENC, steps=[8])),
    'enc_offsets_len': ('SSDInputEncoder', [], dict(ENC, offsets=[0.5])),
    'enc_variances_len': ('SSDInputEncoder', [], dict(ENC, variances=[0.1, 0.1, 0.2])),
    'enc_variances_neg': ('SSDInputEncoder', [], dict(ENC, variances=[0.1, 0.1, 0.2, -0.2])),
    'enc_coords': ('SSDInputEncoder', [], dict(ENC, coords='xyxy')),
    'enc_scales_neg': ('SSDInputEncoder', [], dict(ENC, scales=[0.2, -0.45, 0.8])),
    'enc_ar_neg': ('SSDInputEncoder', [], dict(ENC, aspect_ratios_global=[0.5, -1.0])),
    'enc_ok': ('SSDInputEncoder', [], dict(ENC)),
    'dd_norm_nosize': ('decode_detections', [Y], dict(normalize_coords=True)),
    'dd_coords': ('decode_detections', [Y], dict(input_coords='xyxy', normalize_coords=False)),
    'ddf_norm_nosize': ('decode_detections_fast', [Y], dict(normalize_coords=True)),
    'ddf_coords': ('decode_detections_fast', [Y], dict(input_coords='xyxy', normalize_coords=False)),
    'layer_coords': ('DecodeDetections', [], dict(coords='corners', img_height=10, img_width=10)),
    'layer_norm_nosize': ('DecodeDetections', [], dict(normalize_coords=True)),
    'layerfast_coords': ('DecodeDetectionsFast', [], dict(coords='minmax', img_height=10, img_width=10)),
    'layerfast_norm_nosize': ('DecodeDetectionsFast', [], dict(normalize_coords=True)),
    'ssd300_no_ar': ('ssd_300', [(300, 300, 3), 20], dict(aspect_ratios_global=None, aspect_ratios_per_layer=None, scales=SC300)),
    'ssd300_ar_len': ('ssd_300', [(300, 300, 3), 20], dict(aspect_ratios_per_layer=[[1.0]], scales=SC300)),
    'ssd300_no_scales': ('ssd_300', [(300, 300, 3), 20], dict()),
    'ssd300_scales_len': ('ssd_300', [(300, 300, 3), 20], dict(scales=[.1, .2])),
    'ssd300_var_len': ('ssd_300', [(300, 300, 3), 20], dict(scales=SC300, variances=[.1, .1, .2])),
    'ssd300_var_neg': ('ssd_300', [(300, 300, 3), 20], dict(scales=SC300, variances=[.1, .1, .2, -.2])),
    'ssd300_steps_len': ('ssd_300', [(300, 300, 3), 20], dict(scales=SC300, steps=[8])),
    'ssd300_offsets_len': ('ssd_300', [(300, 300, 3), 20], dict(scales=SC300, offsets=[.5])),
    'ssd512_scales_len': ('ssd_512', [(512, 512, 3), 20], dict(scales=SC300)),
    'ssd512_ar_len': ('ssd_512', [(512, 512, 3), 20], dict(scales=SC512, aspect_ratios_per_layer=[[1.0]] * 6)),
    'ssd7_no_scales': ('build_model', [(300, 300, 3), 5], dict(min_scale=None, max_scale=None)),
    'ssd7_ar_len': ('build_model', [(300, 300, 3), 5], dict(aspect_ratios_per_layer=[[1.0]])),
    'ssd7_var_neg': ('build_model', [(300, 300, 3), 5], dict(variances=[1.0, 1.0, 0.0, 1.0])),
}
TARGETS = dict(SSDInputEncoder=SSDInputEncoder, decode_detections=decode_detections, decode_detections_fast=decode_detections_fast,
               DecodeDetections=DecodeDetections, DecodeDetectionsFast=DecodeDetectionsFast, ssd_300=ssd_300, ssd_512=ssd_512,
               build_model=build_model)


def run(fn, args, kw):
    args = [np.zeros((1, 10, 15), np.float32) if a == Y else (tuple(a) if isinstance(a, list) else a) for a in args]
    try:
        fn(*args, **kw)
        return ['OK', '']
    except (ValueError, TypeError) as e:
        return [type(e).__name__, str(e)]


def main():
    out = {}
    # SSDInputEncoder.__call__ on degenerate ground truth (ssd_input_encoder.py:333-336): its own exception class
    for name, box in (('enc_call_degenerate_x', [1, 10., 10., 10., 50.]), ('enc_call_degenerate_y', [2, 10., 40., 30., 40.])):
        try:
            SSDInputEncoder(**ENC)([np.array([[1, 5., 5., 50., 60.]]), np.array([box])])
            res = ['OK', '']
        except Exception as e:                                   # DegenerateBoxError
            res = [type(e).__name__, str(e)]
        out[name] = dict(target='SSDInputEncoder.__call__', args=[box], kwargs=dict(ENC), result=res)
    for name, (target, args, kw) in CASES.items():
        res = run(TARGETS[target], list(args), dict(kw))
        if target in ('ssd_300', 'ssd_512', 'build_model') and res[0] == 'OK':
            continue
        out[name] = dict(target=target, args=[list(a) if isinstance(a, tuple) else a for a in args], kwargs=kw, result=res)
    with open(os.path.join(HERE, 'ref_errors.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print('wrote %d cases; %d raise' % (len(out), sum(1 for v in out.values() if v['result'][0] != 'OK')))


if __name__ == '__main__':
    main()
