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
