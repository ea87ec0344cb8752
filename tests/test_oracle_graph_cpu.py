"""CPU consistency checks for the training oracle (oracle/graph.py, torch autograd over the product's layer specs):
its forward must equal the hand-written SSD300 forward oracle (oracle/model.py), its loss the NumPy loss oracle
(oracle/loss.py), and autograd's dL/dy_pred the hand-derived gradient (oracle/loss.py::ssd_loss_grad).  None of these is
pinned against TensorFlow (not installable offline); they are three independent restatements agreeing with each other."""
import numpy as np
import torch

from oracle import graph as og
from oracle import synth
from oracle.loss import ssd_loss, ssd_loss_grad


def _tiny_ssd300(hw=96, n_classes=3):
    from ssd_keras_b200.models.keras_ssd300 import ssd_300
    sc = [0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05]
    m = ssd_300((hw, hw, 3), n_classes, mode='training', scales=sc, divide_by_stddev=[64.0] * 3, weights_seed=1)
    return m, sc


def test_graph_forward_matches_model_oracle():
    """Same weights, same image: Spec-list executor == oracle.model.ssd_vgg_forward (float32 both)."""
    from oracle.model import ssd_vgg_forward
    m, sc = _tiny_ssd300(300, 3)           # the hand-written oracle hard-codes the SSD300 layer geometry for 300x300 inputs
    w = m.get_weights()
    x = synth.synth_images(4, 1, 300, 300)
    params = og.make_params(m.specs, w, dtype=torch.float32, requires_grad=False)
    with torch.no_grad():
        y, _ = og.forward(m.specs, params, x, m.n_classes, m.anchors, [0.1, 0.1, 0.2, 0.2], dtype=torch.float32)
    y_ref = ssd_vgg_forward(x, w, 300, 3, scales=sc, divide_by_stddev=[64.0] * 3)
    assert y.shape == y_ref.shape
    np.testing.assert_allclose(y.numpy(), y_ref, rtol=2e-4, atol=2e-5)


def test_loss_and_gradient_match_numpy_oracle():
    rng = np.random.default_rng(0)
    B, P, C = 3, 500, 6
    anchors = rng.uniform(0, 1, (P, 4))
    y_pred = synth.synth_y_pred(1, B, anchors, C, sharp=2.0)
    y_true = np.zeros_like(y_pred)
    cls = rng.integers(0, C, (B, P))
    cls[rng.uniform(size=(B, P)) < 0.8] = 0                     # mostly background
    y_true[np.arange(B)[:, None], np.arange(P)[None, :], cls] = 1.0
    y_true[0, :10, :C] = 0.0                                    # a few neutral boxes
    y_true[:, :, C:C + 4] = rng.standard_normal((B, P, 4))
    yp = torch.tensor(y_pred, dtype=torch.float64, requires_grad=True)
    lvec = og.ssd_loss_torch(y_true, yp)
    ref = ssd_loss(y_true.astype(np.float32), y_pred.astype(np.float32))
    np.testing.assert_allclose(lvec.detach().numpy(), ref, rtol=2e-5)
    lvec.mean().backward()
    g_ref = ssd_loss_grad(y_true.astype(np.float32), y_pred.astype(np.float32))
    g = yp.grad.numpy()
    np.testing.assert_allclose(g[..., :C + 4], g_ref[..., :C + 4], rtol=2e-4, atol=1e-7)
    assert np.all(g[..., C + 4:] == 0)                          # anchors / variances carry no gradient


def test_sgd_step_matches_keras_formula():
    w = {'a/kernel': np.array([1.0, -2.0], np.float32), 'a/bias': np.array([0.5], np.float32)}
    g = {'a/kernel': np.array([0.1, 0.2], np.float32), 'a/bias': np.array([-0.3], np.float32)}
    v = {'a/kernel': np.array([0.01, 0.0]), 'a/bias': np.array([0.0])}
    nw, nv = og.sgd_step(w, g, v, lr=0.1, momentum=0.9, l2_reg=0.01)
    # kernel: g + 2*l2*w, v = m*v - lr*g, w += v ; bias: no regulariser
    np.testing.assert_allclose(nv['a/kernel'], [0.9 * 0.01 - 0.1 * (0.1 + 0.02), -0.1 * (0.2 - 0.04)], rtol=1e-6)
    np.testing.assert_allclose(nw['a/bias'], [0.5 + 0.03], rtol=1e-6)


def test_product_ssd512_specs_match_reference_builder_output():
    """The PRODUCT's layer specs for SSD512 (ssd_keras_b200/models/keras_ssd512.py), executed by the torch Spec executor, against
    the output of the reference's real `ssd_512` builder run over the Keras stand-ins (tests/golden/ref_tf_shim_golden.npz):
    checks paddings, strides, the 4x4 conv10_2, head order and anchors of the plan the CUDA path is built from -- on the CPU."""
    import os
    from oracle.model import vgg_weight_shapes
    from ssd_keras_b200.models.keras_ssd512 import ssd_512
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_tf_shim_golden.npz'))
    sc = [0.04, 0.1, 0.26, 0.42, 0.58, 0.74, 0.9, 1.06]
    m = ssd_512((512, 512, 3), 20, mode='training', scales=sc, subtract_mean=[123, 117, 104], divide_by_stddev=[64, 64, 64],
                swap_channels=[2, 1, 0])
    w = synth.synth_weights(44, vgg_weight_shapes(512, 20), bias_scale=0.02)
    w['conv4_3_norm/gamma'] = np.random.default_rng(44).uniform(10, 30, 512).astype(np.float32)
    assert set(w) == set(m.weight_shapes())
    x = synth.synth_images(43, 1, 512, 512)
    params = og.make_params(m.specs, w, dtype=torch.float32, requires_grad=False)
    with torch.no_grad():
        y, _ = og.forward(m.specs, params, x, m.n_classes, m.anchors, [0.1, 0.1, 0.2, 0.2], dtype=torch.float32)
    y = y.numpy()
    assert y.shape == (1, 24564, 33)
    np.testing.assert_allclose(y[:, ::16], G['model/ssd512/rows16'], rtol=2e-3, atol=1e-4)


def test_product_ssd7_specs_match_reference_builder_output():
    """SSD7 on a 300 x 480 input: the product's builder (BatchNorm folded into the conv plan, ELU, 'same' 5x5 / 3x3 convs, max
    pools, four heads, non-square anchors) through the Spec executor against the reference's real `build_model` output."""
    import os
    from oracle.model import ssd7_weight_shapes
    from ssd_keras_b200.models.keras_ssd7 import build_model
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_tf_shim_golden.npz'))
    m = build_model((300, 480, 3), 5, mode='training', scales=[0.08, 0.16, 0.32, 0.64, 0.96], normalize_coords=True,
                    subtract_mean=127.5, divide_by_stddev=127.5)
    w = synth.synth_weights(46, ssd7_weight_shapes(5), bias_scale=0.05)
    rng = np.random.default_rng(47)
    for i in range(1, 8):
        c = w['conv%d/bias' % i].shape[0]
        w['bn%d/gamma' % i] = rng.uniform(0.8, 1.2, c).astype(np.float32)
        w['bn%d/beta' % i] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        w['bn%d/moving_mean' % i] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        w['bn%d/moving_variance' % i] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    assert set(w) == set(m.weight_shapes())
    x = synth.synth_images(45, 1, 300, 480)
    params = og.make_params(m.specs, w, dtype=torch.float32, requires_grad=False)
    with torch.no_grad():
        y, _ = og.forward(m.specs, params, x, m.n_classes, m.anchors, [1.0, 1.0, 1.0, 1.0], dtype=torch.float32)
    y = y.numpy()
    assert tuple(G['model/ssd7/shape']) == y.shape
    np.testing.assert_allclose(y[:, ::5], G['model/ssd7/rows5'], rtol=2e-3, atol=1e-4)


def test_product_ssd300_specs_match_reference_builder_output():
    """The product's SSD300 layer specs against the reference's real `ssd_300` builder output (golden, 1247 of 8732 prior rows)."""
    import os
    from oracle.model import vgg_weight_shapes
    from ssd_keras_b200.models.keras_ssd300 import ssd_300
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_tf_shim_golden.npz'))
    m = ssd_300((300, 300, 3), 20, mode='training', scales=[0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05], subtract_mean=[123, 117, 104],
                divide_by_stddev=[64, 64, 64], swap_channels=[2, 1, 0])
    w = synth.synth_weights(42, vgg_weight_shapes(300, 20), bias_scale=0.02)
    w['conv4_3_norm/gamma'] = np.random.default_rng(42).uniform(10, 30, 512).astype(np.float32)
    assert set(w) == set(m.weight_shapes())
    x = synth.synth_images(41, 1, 300, 300)
    params = og.make_params(m.specs, w, dtype=torch.float32, requires_grad=False)
    with torch.no_grad():
        y, _ = og.forward(m.specs, params, x, m.n_classes, m.anchors, [0.1, 0.1, 0.2, 0.2], dtype=torch.float32)
    np.testing.assert_allclose(y.numpy()[:, ::7], G['model/ssd300/rows7'], rtol=2e-3, atol=1e-4)
