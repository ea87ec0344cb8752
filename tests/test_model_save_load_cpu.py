"""``SSDModel.save`` / ``models.load_model`` / ``save_weights('*.h5')`` (host code over misc_utils/hdf5_lite.py; reference use:
``ModelCheckpoint`` in ssd300_training.ipynb:404-413 and ``load_model(model_path, custom_objects=...)`` in ssd300_inference.ipynb).
No device needed: no plan is built."""
import inspect

import numpy as np
import pytest

from ssd_keras_b200.misc_utils.hdf5_lite import read_attributes, read_datasets
from ssd_keras_b200.models import load_model
from ssd_keras_b200.models.keras_ssd300 import ssd_300
from ssd_keras_b200.models.keras_ssd7 import build_model


def _randomise(model, seed):
    rng = np.random.default_rng(seed)
    model.set_weights({k: rng.standard_normal(s).astype(np.float32) for k, s in model.weight_shapes().items()})


def test_save_and_load_model_ssd7(tmp_path):
    m = build_model((96, 128, 3), 4, mode='inference', l2_regularization=1e-4, scales=[0.08, 0.16, 0.32, 0.64, 0.96],
                    aspect_ratios_global=[0.5, 1.0, 2.0], variances=np.array([0.1, 0.1, 0.2, 0.2]), normalize_coords=True,
                    subtract_mean=127.5, divide_by_stddev=127.5, confidence_thresh=0.3, top_k=50)
    _randomise(m, 1)
    p = tmp_path / 'ssd7.h5'
    m.save(p)
    assert '/model_weights/conv1/conv1/kernel:0' in read_datasets(str(p))             # Keras' full-model layout
    assert b'SSDModel' in bytes(read_attributes(str(p))['model_config'])
    m2 = load_model(str(p), custom_objects={'AnchorBoxes': None, 'compute_loss': None})
    assert m2.mode == 'inference' and m2.n_classes == 5 and (m2.img_height, m2.img_width) == (96, 128)
    assert m2.decode_cfg == m.decode_cfg and m2.l2_regularization == m.l2_regularization
    assert np.array_equal(m2.anchors, m.anchors) and np.array_equal(m2.variances, m.variances)
    w, w2 = m.get_weights(), m2.get_weights()
    assert sorted(w) == sorted(w2) and all(np.array_equal(w[k], w2[k]) for k in w)


def test_save_and_load_model_ssd300_with_predictor_sizes(tmp_path):
    m, sizes = ssd_300((300, 300, 3), 20, mode='training', scales=[0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05], return_predictor_sizes=True)
    assert sizes.shape == (6, 2)
    m.weights['conv4_3_norm/gamma'][:] = 17.0
    m.weights['conv9_2_mbox_loc/bias'][:] = 0.25
    p = tmp_path / 'ssd300.h5'
    m.save(str(p))
    m2 = load_model(str(p))
    assert m2.mode == 'training' and m2.n_boxes_total == 8732
    assert np.all(m2.weights['conv4_3_norm/gamma'] == 17.0) and np.all(m2.weights['conv9_2_mbox_loc/bias'] == 0.25)
    assert np.array_equal(m2.weights['conv1_1/kernel'], m.weights['conv1_1/kernel'])
    # a plain model built by hand reads the same file by name, like Keras' load_weights(by_name=True)
    m3 = ssd_300((300, 300, 3), 20, mode='inference', scales=[0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05], weights_seed=5)
    m3.load_weights(str(p), by_name=True)
    assert np.array_equal(m3.weights['fc7/kernel'], m.weights['fc7/kernel'])


def test_save_weights_h5_roundtrip_and_errors(tmp_path):
    m = build_model((96, 96, 3), 5, mode='training', scales=[0.08, 0.16, 0.32, 0.64, 0.96])
    _randomise(m, 2)
    p = tmp_path / 'w.h5'
    m.save_weights(str(p))
    assert '/conv1/conv1/kernel:0' in read_datasets(str(p))                          # Keras' weights-file layout
    m2 = build_model((96, 96, 3), 5, mode='training', scales=[0.08, 0.16, 0.32, 0.64, 0.96], weights_seed=9)
    m2.load_weights(str(p))
    assert all(np.array_equal(m.weights[k], m2.weights[k]) for k in m.weights)
    with pytest.raises(ValueError):
        load_model(str(p))                                                           # a weights file has no model_config
    m.save_weights(str(tmp_path / 'w.npz'))
    m2.load_weights(str(tmp_path / 'w.npz'))


def test_builders_keep_their_signatures():
    # the config-recording wrapper must not hide the reference's parameter list (tests/test_signature_parity_cpu.py reads it)
    assert list(inspect.signature(ssd_300).parameters)[:3] == ['image_size', 'n_classes', 'mode']
    assert 'return_predictor_sizes' in inspect.signature(build_model).parameters


def test_count_params_summary_and_shapes():
    m = ssd_300((300, 300, 3), 20, mode='training', scales=[0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05])
    assert m.count_params() == 26285486                      # SSD300 / Pascal VOC, the figure of the reference's model.summary()
    assert m.input_shape == (None, 300, 300, 3) and m.output_shape == (None, 8732, 33)
    mi = ssd_300((300, 300, 3), 20, mode='inference', scales=[0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05], top_k=150)
    assert mi.output_shape == (None, 150, 6)
    lines = []
    m.summary(print_fn=lines.append)
    text = '\n'.join(lines)
    assert 'conv4_3_norm (L2Normalization)' in text and 'conv4_3_norm_mbox_conf (Conv2D)' in text and 'fc7_mbox_loc (Conv2D)' in text
    assert 'Total params: 26,285,486' in text and 'Non-trainable params: 0' in text
    m7 = build_model((300, 480, 3), 5, mode='training', scales=[0.08, 0.16, 0.32, 0.64, 0.96])
    lines = []
    m7.summary(print_fn=lines.append)
    assert any(l.startswith('Non-trainable params:') and not l.endswith(' 0') for l in lines)     # BatchNormalization statistics
