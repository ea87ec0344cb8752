"""Host logic of the Keras-style training surface (models/_keras_api.py, optimizers.py): argument handling of ``compile`` and the
``fit_generator`` loop (epochs / initial_epoch / history / callbacks / validation) with the device step stubbed out.  Reference use:
ssd300_training.ipynb:153-173 (compile) and :437-448 (fit_generator)."""
import itertools

import numpy as np
import pytest

from ssd_keras_b200 import optimizers
from ssd_keras_b200.keras_loss_function.keras_ssd_loss import SSDLoss
from ssd_keras_b200.models.keras_ssd7 import build_model


def _model():
    return build_model((96, 96, 3), 5, mode='training', scales=[0.08, 0.16, 0.32, 0.64, 0.96])


def test_optimizer_holders_follow_keras_signatures():
    s = optimizers.SGD(lr=0.001, momentum=0.9, decay=0.0, nesterov=False)
    assert (s.lr, s.momentum, s.kind) == (0.001, 0.9, 'sgd')
    a = optimizers.Adam(lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-08, decay=0.0)
    assert (a.lr, a.beta_1, a.beta_2, a.epsilon, a.kind) == (0.001, 0.9, 0.999, 1e-08, 'adam')
    assert optimizers.SGD().lr == 0.01 and optimizers.Adam().lr == 0.001          # Keras' defaults
    with pytest.raises(ValueError):
        optimizers.SGD(nesterov=True)
    with pytest.raises(ValueError):
        optimizers.SGD(decay=1e-4)
    with pytest.raises(ValueError):
        optimizers.Adam(decay=1e-4)


def test_compile_argument_handling():
    m = _model()
    loss = SSDLoss(neg_pos_ratio=3, n_neg_min=0, alpha=1.0)
    m.compile(optimizer=optimizers.SGD(lr=1e-3, momentum=0.9), loss=loss.compute_loss)      # the reference's call
    assert m._compiled[0] == 'sgd' and m._compiled[1] is loss
    m.compile(optimizer='adam', loss=loss)
    assert m._compiled[0] == 'adam' and m.optimizer.lr == 0.001
    with pytest.raises(ValueError):
        m.compile(optimizer='rmsprop', loss=loss.compute_loss)
    with pytest.raises(ValueError):
        m.compile(optimizer=optimizers.SGD(), loss='mse')
    with pytest.raises(ValueError):
        m.compile(optimizer=object(), loss=loss.compute_loss)
    with pytest.raises(TypeError):
        m.compile(optimizer='sgd', loss=loss.compute_loss, metrics=['accuracy'])
    with pytest.raises(RuntimeError):
        _model().train_on_batch(np.zeros((1, 96, 96, 3), np.float32), np.zeros((1, 10, 18), np.float32))


class _Recorder:
    def __init__(self):
        self.events = []

    def set_model(self, model):
        self.events.append('set_model')

    def on_train_begin(self, logs):
        self.events.append('train_begin')

    def on_epoch_begin(self, epoch, logs):
        self.events.append(('epoch_begin', epoch))

    def on_epoch_end(self, epoch, logs):
        self.events.append(('epoch_end', epoch, round(logs['loss'], 6), round(logs.get('val_loss', -1.0), 6)))

    def on_train_end(self, logs):
        self.events.append('train_end')


def test_fit_generator_loop_with_stubbed_device_step(monkeypatch):
    m = _model()
    m.compile(optimizer=optimizers.SGD(lr=1e-3, momentum=0.9), loss=SSDLoss().compute_loss)
    losses = itertools.count(10, -1)                    # 10, 9, 8, ...
    seen_lr = []
    monkeypatch.setattr(m, 'train_on_batch', lambda x, y: (seen_lr.append(m.optimizer.lr), float(next(losses)))[1])
    monkeypatch.setattr(m, 'test_on_batch', lambda x, y: 0.5)
    gen = ((np.zeros((4, 96, 96, 3), np.float32), np.zeros((4, 10, 18), np.float32)) for _ in itertools.count())
    rec = _Recorder()

    class HalveLR:                                      # what keras.callbacks.LearningRateScheduler does for the reference
        def on_epoch_begin(self, epoch, logs):
            self.model.optimizer.lr = 1e-3 * 0.5 ** epoch

        def set_model(self, model):
            self.model = model
    h = m.fit_generator(gen, steps_per_epoch=3, epochs=4, callbacks=[rec, HalveLR()], validation_data=gen, validation_steps=2, initial_epoch=1)
    assert h.epoch == [1, 2, 3]
    assert h.history['loss'] == [9.0, 6.0, 3.0]         # means of (10, 9, 8), (7, 6, 5), (4, 3, 2)
    assert h.history['val_loss'] == [0.5, 0.5, 0.5]
    assert seen_lr == [5e-4] * 3 + [2.5e-4] * 3 + [1.25e-4] * 3
    assert rec.events[0] == 'set_model' and rec.events[1] == 'train_begin' and rec.events[-1] == 'train_end'
    assert ('epoch_end', 2, 6.0, 0.5) in rec.events and ('epoch_begin', 1) in rec.events and ('epoch_begin', 0) not in rec.events
    with pytest.raises(ValueError):
        m.fit_generator(gen, steps_per_epoch=1, epochs=1, validation_data=gen)

    class StopAfterFirst:
        def set_model(self, model):
            self.model = model

        def on_batch_end(self, batch, logs):
            self.model.stop_training = True
    h = m.fit_generator(gen, steps_per_epoch=5, epochs=3, callbacks=[StopAfterFirst()])
    assert h.epoch == [0] and len(h.history['loss']) == 1
