"""Keras-HDF5 weight ingestion without h5py: the reader of misc_utils/hdf5_lite.py on synthetic files
in the layout Keras writes (``/<layer>/<layer>/kernel:0``, ``layer_names`` / ``weight_names`` attributes; full-model files with
``/model_weights``; chunked + shuffle + gzip datasets), and SSDModel.load_weights(by_name=True) on top of it.
The files are produced by the module's own writer, which emits the same structures as h5py's default (superblock 0, symbol-table
groups, version-1 object headers); no real h5py file can be produced or fetched offline."""
import struct

import numpy as np
import pytest

from ssd_keras_b200.misc_utils import hdf5_lite as h5


def _weights(n_extra=40):
    rng = np.random.default_rng(0)
    w = {'conv1_1/kernel': rng.standard_normal((3, 3, 3, 64)).astype(np.float32), 'conv1_1/bias': rng.standard_normal(64).astype(np.float32),
         'conv4_3_norm/gamma': np.full(512, 20, np.float32), 'fc6/kernel': rng.standard_normal((3, 3, 8, 16)).astype(np.float64),
         'bn1/moving_variance': rng.uniform(0.5, 1.5, 32).astype(np.float32)}
    for i in range(n_extra):                                        # more links than one symbol-table node holds
        w['extra%02d/kernel' % i] = rng.standard_normal((1, 1, 4, 5)).astype(np.float32)
    return w


@pytest.mark.parametrize('full_model', [False, True])
def test_keras_layout_round_trip(tmp_path, full_model):
    w = _weights()
    p = str(tmp_path / 'w.h5')
    h5.write_keras_weights(p, w, full_model=full_model, chunked=('conv1_1/kernel', 'fc6/kernel'))
    raw = open(p, 'rb').read()
    assert raw[:8] == b'\x89HDF\r\n\x1a\n' and raw[8] == 0 and raw[13] == 8 and raw[14] == 8        # superblock 0, 8-byte offsets
    assert struct.unpack_from('<HH', raw, 16) == (4, 16)                                            # group leaf / internal K
    r = h5.read_keras_weights(p)
    assert set(r) == set(w)
    for k in w:
        assert r[k].dtype == w[k].dtype and np.array_equal(r[k], w[k]), k
    ds = h5.read_datasets(p)
    prefix = '/model_weights' if full_model else ''
    assert prefix + '/conv1_1/conv1_1/kernel:0' in ds and prefix + '/conv4_3_norm/conv4_3_norm/gamma:0' in ds
    attrs = h5.read_attributes(p, prefix or '/')
    assert [s.decode() for s in attrs['layer_names']][:2] == ['conv1_1', 'conv4_3_norm']
    assert [s.decode() for s in h5.read_attributes(p, prefix + '/conv1_1')['weight_names']] == ['conv1_1/kernel:0', 'conv1_1/bias:0']


def test_unsupported_files_fail_loudly(tmp_path):
    p = str(tmp_path / 'x.h5')
    open(p, 'wb').write(b'not hdf5 at all')
    with pytest.raises(ValueError):
        h5.read_datasets(p)
    w = {'a/kernel': np.zeros((2, 2), np.float32)}
    h5.write_keras_weights(p, w)
    raw = bytearray(open(p, 'rb').read())
    raw[8] = 2                                                       # pretend superblock version 2 (libver='latest')
    open(p, 'wb').write(bytes(raw))
    with pytest.raises(NotImplementedError):
        h5.read_datasets(p)


def test_model_load_weights_by_name_from_hdf5(tmp_path):
    """Host-side only (no plan is built): names present in the file replace the model's weights, others are skipped / kept."""
    from ssd_keras_b200.models.keras_ssd7 import build_model
    m = build_model((96, 96, 3), 5, scales=[0.08, 0.16, 0.32, 0.64, 0.96], weights_seed=1)
    src = build_model((96, 96, 3), 5, scales=[0.08, 0.16, 0.32, 0.64, 0.96], weights_seed=2).weights
    before = {k: v.copy() for k, v in m.weights.items()}
    subset = {k: v for k, v in src.items() if not k.startswith('conv7')}
    subset['not_in_this_model/kernel'] = np.zeros((1, 1, 2, 2), np.float32)
    p = str(tmp_path / 'ssd7.h5')
    h5.write_keras_weights(p, subset, chunked=('conv1/kernel',))
    m.load_weights(p, by_name=True)
    for k in m.weights:
        if k.startswith('conv7'):
            np.testing.assert_array_equal(m.weights[k], before[k])
        else:
            np.testing.assert_array_equal(m.weights[k], src[k])
    bad = dict(subset); bad['conv1/kernel'] = np.zeros((3, 3, 3, 32), np.float32)
    h5.write_keras_weights(p, bad)
    with pytest.raises(ValueError):
        m.load_weights(p)
    with pytest.raises(ValueError):
        open(str(tmp_path / 'w.bin'), 'wb').write(b'0' * 64)
        m.load_weights(str(tmp_path / 'w.bin'))
