"""GPU test of the Keras-style training surface: ``compile`` + ``train_on_batch`` / ``fit_generator`` / ``test_on_batch`` on an
SSD7 graph (BatchNormalization + ELU, Adam: ssd7_training.ipynb:153-156, 330-340) and on a small VGG-style graph with SGD-momentum
(ssd300_training.ipynb:169-173).  The step itself is checked against float64 autograd in test_gpu_train.py; here: the scalar Keras
reports equals the mean of the trainer's per-image loss, the loop trains (loss falls on a fixed batch), the trained weights belong
to the model, and a different batch size is refused."""
import numpy as np
import pytest

from oracle import synth
from oracle.model import ssd7_weight_shapes

pytestmark = pytest.mark.gpu

SC7 = [0.08, 0.16, 0.32, 0.64, 0.96]


@pytest.fixture(scope='module', autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()
    import torch
    assert torch.cuda.is_available()


def _ssd7():
    from ssd_keras_b200.models.keras_ssd7 import build_model
    m = build_model((96, 96, 3), 5, mode='training', scales=SC7, normalize_coords=True, subtract_mean=[127.5] * 3, divide_by_stddev=[127.5] * 3)
    w = synth.synth_weights(1, ssd7_weight_shapes(5), bias_scale=0.05)
    for i in range(1, 8):
        c = w['conv%d/bias' % i].shape[0]
        w['bn%d/gamma' % i] = np.ones(c, np.float32); w['bn%d/beta' % i] = np.zeros(c, np.float32)
        w['bn%d/moving_mean' % i] = np.zeros(c, np.float32); w['bn%d/moving_variance' % i] = np.ones(c, np.float32)
    m.set_weights(w)
    return m, w


def _batches(model, B, n):
    from ssd_keras_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
    enc = SSDInputEncoder(96, 96, 5, model.predictor_sizes, scales=SC7, variances=[1.0] * 4, pos_iou_threshold=0.5, neg_iou_limit=0.3,
                          normalize_coords=True)
    out = []
    for i in range(n):
        x = synth.synth_images(10 + i, B, 96, 96)
        y = enc(synth.synth_gt(20 + i, B, 4, 96, 96, 5)).astype(np.float32)
        out.append((x, y))
    return out


def test_ssd7_compile_fit_generator_adam():
    import itertools
    import torch
    from ssd_keras_b200.keras_loss_function.keras_ssd_loss import SSDLoss
    from ssd_keras_b200.optimizers import Adam
    from ssd_keras_b200.training import SSDTrainer
    model, w0 = _ssd7()
    B = 4
    data = _batches(model, B, 2)
    # the scalar of the first step = mean of the per-image loss a bare trainer computes on the same weights
    ref_model, _ = _ssd7()
    tr = SSDTrainer(ref_model, B, lr=1e-3, optimizer='adam')
    l_ref, _, _ = tr._loss_and_dy(torch.from_numpy(data[0][0]).cuda(), torch.from_numpy(data[0][1]).cuda())
    model.compile(optimizer=Adam(lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-08, decay=0.0), loss=SSDLoss(neg_pos_ratio=3, alpha=1.0).compute_loss)
    l0 = model.train_on_batch(*data[0])
    assert abs(l0 - float(l_ref.mean().item())) <= 1e-5 * abs(l0)
    # fit on a repeating pair of batches: the epoch loss falls, the history has Keras' shape
    gen = itertools.cycle(data)
    epochs_seen = []

    class Log:
        def on_epoch_end(self, epoch, logs):
            epochs_seen.append((epoch, logs['loss'], logs['val_loss']))
    h = model.fit_generator(gen, steps_per_epoch=6, epochs=4, callbacks=[Log()], validation_data=itertools.cycle(data), validation_steps=2, initial_epoch=1)
    assert h.epoch == [1, 2, 3] and len(h.history['loss']) == 3 and len(h.history['val_loss']) == 3
    assert [e[0] for e in epochs_seen] == [1, 2, 3]
    assert all(np.isfinite(v) for v in h.history['loss'] + h.history['val_loss'])
    assert h.history['loss'][-1] < l0 and h.history['loss'][-1] < h.history['loss'][0]
    # the trained weights are the model's: get_weights / predict see them, test_on_batch agrees with evaluate_generator
    w1 = model.get_weights()
    assert np.abs(w1['conv1/kernel'] - w0['conv1/kernel']).max() > 1e-4
    assert not np.allclose(w1['bn1/moving_mean'], 0.0)
    v = model.test_on_batch(*data[0])
    assert np.isfinite(v) and abs(model.evaluate_generator(iter([data[0]]), 1) - v) < 1e-6 * max(1.0, abs(v))
    # another batch size would need its own optimizer state
    with pytest.raises(ValueError):
        model.train_on_batch(data[0][0][:2], data[0][1][:2])


def test_small_vgg_graph_sgd_momentum_lr_change():
    import importlib.util
    import os
    from ssd_keras_b200.keras_loss_function.keras_ssd_loss import SSDLoss
    from ssd_keras_b200.optimizers import SGD
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
    spec = importlib.util.spec_from_file_location('train_check', os.path.join(root, 'tools', 'train_check.py'))
    tc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tc)
    case = tc.CASES[1]
    m, w, n_cls = tc.build(case)
    hw, B = case[1], case[2]
    from oracle.encoder import OracleEncoder
    enc = OracleEncoder(hw, hw, n_cls - 1, m.predictor_sizes, scales=m.anchor_cfg['scales'], aspect_ratios_per_layer=m.anchor_cfg['aspect_ratios_per_layer'],
                        variances=[0.1, 0.1, 0.2, 0.2], pos_iou_threshold=0.3, neg_iou_limit=0.2)
    x = np.random.default_rng(11).integers(0, 256, size=(B, hw, hw, 3)).astype(np.float32)
    y = enc(tc.small_gt(5, B, 3, hw, n_cls - 1)).astype(np.float32)
    m.compile(optimizer=SGD(lr=1e-3, momentum=0.9, decay=0.0, nesterov=False), loss=SSDLoss(neg_pos_ratio=3, alpha=1.0).compute_loss)
    first = m.train_on_batch(x, y)
    for _ in range(5):
        last = m.train_on_batch(x, y)
    assert np.isfinite(first) and last < first
    m.optimizer.lr = 0.0                                 # what a LearningRateScheduler does; momentum still carries the weights on
    m.train_on_batch(x, y)
    assert m._fit_trainer.lr == 0.0
