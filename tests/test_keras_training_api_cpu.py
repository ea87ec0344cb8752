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


def test_reference_notebook_callbacks_with_stubbed_device_step(monkeypatch, tmp_path):
    """ModelCheckpoint / LearningRateScheduler / TerminateOnNaN / CSVLogger as ssd300_training.ipynb:404-425 sets them up, plus
    EarlyStopping / ReduceLROnPlateau of ssd7_training.ipynb:300-325."""
    from ssd_keras_b200 import callbacks as cb
    from ssd_keras_b200.misc_utils.hdf5_lite import read_datasets
    from ssd_keras_b200.models import load_model
    m = _model()
    m.compile(optimizer=optimizers.SGD(lr=1e-3, momentum=0.9), loss=SSDLoss().compute_loss)
    train = iter([5.0, 5.0, 4.0, 4.0, 3.5, 3.5, 3.6, 3.6, 3.7, 3.7, 3.8, 3.8, 9.0, 9.0])
    val = iter([4.0, 3.0, 3.2, 3.3, 3.4, 3.5, 3.6])
    monkeypatch.setattr(m, 'train_on_batch', lambda x, y: next(train))
    monkeypatch.setattr(m, 'evaluate_generator', lambda g, steps: next(val))
    gen = ((np.zeros((2, 96, 96, 3), np.float32), np.zeros((2, 10, 18), np.float32)) for _ in itertools.count())

    def lr_schedule(epoch):
        return 0.001 if epoch < 2 else 0.0001
    ckpt = cb.ModelCheckpoint(filepath=str(tmp_path / 'ssd_epoch-{epoch:02d}_loss-{loss:.4f}_val_loss-{val_loss:.4f}.h5'), monitor='val_loss',
                              verbose=0, save_best_only=True, save_weights_only=False, mode='auto', period=1)
    log = cb.CSVLogger(filename=str(tmp_path / 'log.csv'), separator=',', append=True)
    stop = cb.EarlyStopping(monitor='val_loss', min_delta=0.0, patience=3, verbose=0)
    h = m.fit_generator(gen, steps_per_epoch=2, epochs=7, callbacks=[ckpt, log, cb.LearningRateScheduler(schedule=lr_schedule, verbose=0),
                                                                      cb.TerminateOnNaN(), stop], validation_data=gen, validation_steps=1)
    # val_loss: 4.0, 3.0 (best), 3.2, 3.3, 3.4 -> three epochs without improvement: stop after epoch index 4
    assert h.epoch == [0, 1, 2, 3, 4] and stop.stopped_epoch == 4
    saved = sorted(p.name for p in tmp_path.glob('*.h5'))
    assert saved == ['ssd_epoch-01_loss-5.0000_val_loss-4.0000.h5', 'ssd_epoch-02_loss-4.0000_val_loss-3.0000.h5']     # only improvements
    assert '/model_weights/conv1/conv1/kernel:0' in read_datasets(str(tmp_path / saved[1]))
    assert load_model(str(tmp_path / saved[1])).n_classes == m.n_classes
    assert m.optimizer.lr == 0.0001
    rows = (tmp_path / 'log.csv').read_text().strip().split('\n')
    assert rows[0] == 'epoch,loss,val_loss' and rows[2].startswith('1,4.0,3.0') and len(rows) == 6
    # ReduceLROnPlateau: factor 0.5 after 2 epochs without an improvement of more than epsilon
    m.optimizer.lr = 0.01
    r = cb.ReduceLROnPlateau(monitor='val_loss', factor=0.5, patience=2, verbose=0, epsilon=0.001, cooldown=0, min_lr=0.004)
    r.set_model(m)
    for e, v in enumerate([1.0, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9]):
        r.on_epoch_end(e, {'val_loss': v})
    assert m.optimizer.lr == 0.004                          # 0.01 -> 0.005 -> max(0.0025, min_lr)
    # TerminateOnNaN
    t = cb.TerminateOnNaN(); t.set_model(m); m.stop_training = False
    t.on_batch_end(0, {'loss': float('nan')})
    assert m.stop_training
    with pytest.raises(ValueError):
        cb.ModelCheckpoint('x.h5', mode='sideways')
    with pytest.raises(ValueError):
        cb.ReduceLROnPlateau(factor=1.0)
    # weights-only checkpoints every second epoch
    m.stop_training = False
    c2 = cb.ModelCheckpoint(str(tmp_path / 'w-{epoch:02d}.h5'), save_weights_only=True, period=2)
    c2.set_model(m)
    for e in range(4):
        c2.on_epoch_end(e, {'loss': 1.0})
    assert sorted(p.name for p in tmp_path.glob('w-*.h5')) == ['w-02.h5', 'w-04.h5']
    assert '/conv1/conv1/kernel:0' in read_datasets(str(tmp_path / 'w-02.h5'))
