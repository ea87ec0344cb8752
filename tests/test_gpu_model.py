"""GPU parity tests for the tcgen05 convolution plan: SSD7 / SSD300 / SSD512 forward against the float32
torch-CPU oracle graphs (oracle/model.py) on identical synthetic weights and images.

Tolerances (stated here, measured in DESIGN.md): the default 'bf16x3' mode splits every operand into bf16
hi+lo and issues hi*hi + hi*lo + lo*hi with fp32 accumulation, i.e. ~16 significant bits per product; layer
outputs are compared at 2e-4 of the layer's max magnitude and the final class probabilities / offsets at 1e-4
absolute.  The single-pass 'bf16' mode is checked at 5e-2."""
import numpy as np
import pytest

from oracle import synth
from oracle import decoder as odec
from oracle.model import ssd7_forward, ssd7_weight_shapes, ssd_vgg_forward, vgg_weight_shapes

pytestmark = pytest.mark.gpu

SC300 = [0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05]
SC512 = [0.04, 0.1, 0.26, 0.42, 0.58, 0.74, 0.9, 1.06]
SC7 = [0.08, 0.16, 0.32, 0.64, 0.96]


@pytest.fixture(scope='module', autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()
    import torch
    assert torch.cuda.is_available()


def _ssd7_weights(seed, n_classes=5):
    w = synth.synth_weights(seed, ssd7_weight_shapes(n_classes), bias_scale=0.05)
    rng = np.random.default_rng(seed + 100)
    for i in range(1, 8):
        c = w['conv%d/bias' % i].shape[0]
        w['bn%d/gamma' % i] = rng.uniform(0.8, 1.2, c).astype(np.float32)
        w['bn%d/beta' % i] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        w['bn%d/moving_mean' % i] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        w['bn%d/moving_variance' % i] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    return w


def _vgg_weights(seed, variant, n_classes):
    w = synth.synth_weights(seed, vgg_weight_shapes(variant, n_classes), bias_scale=0.02)
    w['conv4_3_norm/gamma'] = np.full((512,), 20.0, np.float32)
    return w


def _prob_atol(feats, rel=3e-5):
    """Class probabilities are softmax(logits): an error d in a logit moves a probability by at most d/4... d.
    The conv path is accurate to ~1e-5 RELATIVE to the logit magnitude (which is in the hundreds for he_normal
    weights on raw 0..255 images), so the absolute tolerance on probabilities scales with max|logit|."""
    return max(1e-4, rel * float(np.abs(feats['logits']).max()))


def _cmp_layers(model, feats, B, names, tol):
    for n in names:
        got = model.read_layer(n, B)
        ref = feats[n]
        scale = np.abs(ref).max()
        err = np.abs(got - ref).max()
        assert got.shape == ref.shape, (n, got.shape, ref.shape)
        assert err <= tol * scale + 1e-6, 'layer %s: max err %.3e vs scale %.3e' % (n, err, scale)


def test_ssd7_config0_forward_and_decode():
    """BASELINE config 0: SSD7 forward + decode_detections on one synthetic 300x300 image."""
    from ssd_keras_b200.models.keras_ssd7 import build_model
    from ssd_keras_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections
    w = _ssd7_weights(1)
    x = synth.synth_images(0, 1, 300, 300)
    model = build_model((300, 300, 3), 5, mode='training', scales=SC7, normalize_coords=True)
    model.set_weights(w)
    y = model.predict(x)
    y_ref, feats = ssd7_forward(x, w, n_classes=5, scales=SC7, normalize_coords=True, return_features=True)
    assert y.shape == (1, 7160, 18)
    _cmp_layers(model, feats, 1, ['conv1', 'conv4', 'conv7'], 2e-4)
    np.testing.assert_allclose(y[:, :, :6], y_ref[:, :, :6], atol=_prob_atol(feats))   # softmax probabilities
    np.testing.assert_allclose(y[:, :, 6:10], y_ref[:, :, 6:10], atol=2e-4 * np.abs(y_ref[:, :, 6:10]).max())
    np.testing.assert_array_equal(y[:, :, 10:], y_ref[:, :, 10:])                    # anchors + variances: bit-exact
    # decode on identical inputs (the oracle's y_pred): same survivors as the NumPy reference path
    kw = dict(confidence_thresh=0.3, iou_threshold=0.45, top_k=200, img_height=300, img_width=300)
    got = decode_detections(y_ref, **kw)
    exp = odec.decode_detections(y_ref, **kw)
    assert got[0].shape == np.asarray(exp[0]).reshape(-1, 6).shape
    # the identity-preprocessing SSD7 variant (divide_by_stddev / subtract_mean / swap) also runs
    m2 = build_model((300, 300, 3), 5, mode='training', scales=SC7, normalize_coords=True, subtract_mean=127.5 * np.ones(3),
                     divide_by_stddev=127.5 * np.ones(3), swap_channels=[2, 1, 0])
    m2.set_weights(w)
    y2 = m2.predict(x)
    y2_ref, f2 = ssd7_forward(x, w, n_classes=5, scales=SC7, normalize_coords=True, subtract_mean=[127.5] * 3,
                              divide_by_stddev=[127.5] * 3, swap_channels=[2, 1, 0], return_features=True)
    np.testing.assert_allclose(y2[:, :, :6], y2_ref[:, :, :6], atol=_prob_atol(f2))


@pytest.mark.parametrize('precision,tol', [('bf16x3', 2e-4), ('bf16', 5e-2)])
def test_ssd300_forward_layers(precision, tol):
    from ssd_keras_b200.models.keras_ssd300 import ssd_300
    B = 2
    w = _vgg_weights(1, 300, 20)
    x = synth.synth_images(0, B, 300, 300)
    model = ssd_300((300, 300, 3), 20, mode='training', scales=SC300, precision=precision)
    model.set_weights(w)
    y = model.predict(x)
    y_ref, feats = ssd_vgg_forward(x, w, 300, 20, scales=SC300, return_features=True)
    assert y.shape == (B, 8732, 33)
    _cmp_layers(model, feats, B, ['conv1_1', 'conv1_2', 'conv2_2', 'conv3_3', 'conv4_3', 'conv5_3', 'fc6', 'fc7', 'conv6_2',
                                  'conv7_2', 'conv8_2', 'conv9_2', 'conv4_3_norm'], tol)
    np.testing.assert_array_equal(y[:, :, 25:], y_ref[:, :, 25:])
    ptol = _prob_atol(feats) if precision == 'bf16x3' else _prob_atol(feats, 1e-2)
    np.testing.assert_allclose(y[:, :, :21], y_ref[:, :, :21], atol=ptol)
    np.testing.assert_allclose(y[:, :, 21:25], y_ref[:, :, 21:25], atol=max(tol, 2e-4) * np.abs(y_ref[:, :, 21:25]).max())


def test_ssd300_batch32_layers():
    """The benchmark configuration itself (BASELINE configs[1]: batch 32): at this size the 64/128-channel layers run with
    two m-tiles per work unit and the deep layers with the cross-term accumulator, plans that small batches do not select."""
    from ssd_keras_b200.models.keras_ssd300 import ssd_300
    B = 32
    w = _vgg_weights(5, 300, 20)
    x = synth.synth_images(6, B, 300, 300)
    model = ssd_300((300, 300, 3), 20, mode='training', scales=SC300)
    model.set_weights(w)
    y = model.predict(x)
    import os
    import torch
    threads = torch.get_num_threads()
    torch.set_num_threads(min(32, os.cpu_count() or 1))        # torch's CPU convolutions are slowest with every core of a big host
    try:
        y_ref, feats = ssd_vgg_forward(x, w, 300, 20, scales=SC300, return_features=True)
    finally:
        torch.set_num_threads(threads)
    assert y.shape == (B, 8732, 33)
    _cmp_layers(model, feats, B, ['conv1_2', 'conv2_1', 'conv2_2', 'conv3_1', 'conv3_3', 'conv4_3', 'fc7', 'conv6_2'], 2e-4)
    np.testing.assert_array_equal(y[:, :, 25:], y_ref[:, :, 25:])
    np.testing.assert_allclose(y[:, :, :21], y_ref[:, :, :21], atol=_prob_atol(feats))
    np.testing.assert_allclose(y[:, :, 21:25], y_ref[:, :, 21:25], atol=2e-4 * np.abs(y_ref[:, :, 21:25]).max())


def test_ssd300_inference_mode_matches_layer_oracle():
    """mode='inference': (B,200,6); decoded from identical y_pred the output equals the layer oracle bit for bit in
    survivor set; against the fp32 oracle's own y_pred the boxes agree within 1e-4 relative where survivors coincide."""
    import torch
    from ssd_keras_b200.models.keras_ssd300 import ssd_300
    B = 2
    w = _vgg_weights(1, 300, 20)
    x = synth.synth_images(0, B, 300, 300)
    model = ssd_300((300, 300, 3), 20, mode='inference', scales=SC300)
    model.set_weights(w)
    xd = torch.from_numpy(x).cuda()
    y_pred = model.forward_device(xd)
    out = model.predict(x)
    assert out.shape == (B, 200, 6)
    ref = odec.decode_layer(y_pred.cpu().numpy(), 0.01, 0.45, 200, 400, True, 300, 300)
    np.testing.assert_array_equal(out[:, :, :2], ref[:, :, :2])
    np.testing.assert_allclose(out[:, :, 2:], ref[:, :, 2:], rtol=1e-6, atol=1e-4)
    fast = ssd_300((300, 300, 3), 20, mode='inference_fast', scales=SC300)
    fast.set_weights(w)
    outf = fast.predict(x)
    reff = odec.decode_layer_fast(y_pred.cpu().numpy(), 0.01, 0.45, 200, 400, True, 300, 300)
    np.testing.assert_array_equal(outf[:, :, :2], reff[:, :, :2])


def test_ssd512_forward():
    from ssd_keras_b200.models.keras_ssd512 import ssd_512
    B = 1
    w = _vgg_weights(2, 512, 80)
    x = synth.synth_images(3, B, 512, 512)
    model = ssd_512((512, 512, 3), 80, mode='training', scales=SC512)
    model.set_weights(w)
    y = model.predict(x)
    y_ref, feats = ssd_vgg_forward(x, w, 512, 80, scales=SC512, return_features=True)
    assert y.shape == (B, 24564, 93)
    _cmp_layers(model, feats, B, ['conv4_3', 'fc7', 'conv8_2', 'conv9_2', 'conv10_2'], 2e-4)
    np.testing.assert_allclose(y[:, :, :81], y_ref[:, :, :81], atol=_prob_atol(feats))
    np.testing.assert_array_equal(y[:, :, 85:], y_ref[:, :, 85:])


def test_l2_normalization_layer():
    import torch
    from ssd_keras_b200.keras_layers.keras_layer_L2Normalization import L2Normalization
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 7, 5, 512)).astype(np.float32) * 3
    x[0, 0, 0] = 0                                             # zero vector: the 1e-12 clamp
    out = L2Normalization(gamma_init=20)(x)
    ss = np.maximum((x.astype(np.float64) ** 2).sum(-1, keepdims=True), 1e-12)
    np.testing.assert_allclose(out, x / np.sqrt(ss) * 20.0, rtol=2e-6, atol=1e-6)


def test_weights_roundtrip(tmp_path):
    from ssd_keras_b200.models.keras_ssd7 import build_model
    m = build_model((96, 96, 3), 5, scales=SC7)
    p = str(tmp_path / 'w.npz')
    m.save_weights(p)
    m2 = build_model((96, 96, 3), 5, scales=SC7, weights_seed=9)
    m2.load_weights(p, by_name=True)
    for k, v in m.get_weights().items():
        np.testing.assert_array_equal(v, m2.get_weights()[k])
    with pytest.raises(ValueError):
        m2.set_weights({'conv1/kernel': np.zeros((3, 3, 3, 32), np.float32)})
