"""The four Keras callbacks the reference's training notebooks pass to ``fit_generator`` (``ssd300_training.ipynb:404-425``,
``ssd7_training.ipynb:300-325``): ``ModelCheckpoint``, ``LearningRateScheduler``, ``TerminateOnNaN``, ``CSVLogger`` -- and
``EarlyStopping`` / ``ReduceLROnPlateau`` from the SSD7 notebook -- with Keras' argument names, for ``SSDModel.fit_generator``
(models/_keras_api.py), which calls ``set_model``, ``on_train_begin/end``, ``on_epoch_begin/end`` and ``on_batch_end``."""
import csv
import os

import numpy as np


class Callback:
    def set_model(self, model):
        self.model = model


def _mode(mode, monitor):
    if mode not in ('auto', 'min', 'max'):
        raise ValueError("mode must be 'auto', 'min' or 'max'")
    if mode == 'auto':
        mode = 'max' if ('acc' in monitor or monitor.startswith('fmeasure')) else 'min'
    return mode


def _better(mode, monitor):
    """(is a better than b?, the worst possible value) for Keras' ``mode`` of a monitored quantity."""
    return (lambda a, b: a > b, -np.inf) if _mode(mode, monitor) == 'max' else (lambda a, b: a < b, np.inf)


class ModelCheckpoint(Callback):
    """``filepath`` may contain ``{epoch:02d}`` (1-based like Keras) and any key of ``logs``, e.g. ``{val_loss:.4f}``."""

    def __init__(self, filepath, monitor='val_loss', verbose=0, save_best_only=False, save_weights_only=False, mode='auto', period=1):
        self.filepath, self.monitor, self.verbose = str(filepath), monitor, verbose
        self.save_best_only, self.save_weights_only, self.period = save_best_only, save_weights_only, int(period)
        self.op, self.best = _better(mode, monitor)
        self.epochs_since_last_save = 0

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if self.save_best_only:
            cur = logs.get(self.monitor)
            if cur is None or not self.op(cur, self.best):
                return
            self.best = cur
        if self.verbose:
            print('Epoch %05d: saving model to %s' % (epoch + 1, path))
        (self.model.save_weights if self.save_weights_only else self.model.save)(path)


class LearningRateScheduler(Callback):
    def __init__(self, schedule, verbose=0):
        self.schedule, self.verbose = schedule, verbose

    def on_epoch_begin(self, epoch, logs=None):
        lr = self.schedule(epoch)
        if not isinstance(lr, (float, np.floating)):
            raise ValueError('The output of the "schedule" function should be float.')
        self.model.optimizer.lr = float(lr)
        if self.verbose:
            print('Epoch %05d: LearningRateScheduler setting learning rate to %s.' % (epoch + 1, lr))


class TerminateOnNaN(Callback):
    def on_batch_end(self, batch, logs=None):
        loss = (logs or {}).get('loss')
        if loss is not None and not np.isfinite(loss):
            print('Batch %d: Invalid loss, terminating training' % batch)
            self.model.stop_training = True


class CSVLogger(Callback):
    def __init__(self, filename, separator=',', append=False):
        self.filename, self.sep, self.append = str(filename), separator, append
        self.keys, self.file, self.writer = None, None, None

    def on_train_begin(self, logs=None):
        self.append_header = not (self.append and os.path.exists(self.filename) and os.path.getsize(self.filename) > 0)
        self.file = open(self.filename, 'a' if self.append else 'w', newline='')

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        if self.keys is None:
            self.keys = sorted(logs)
            self.writer = csv.writer(self.file, delimiter=self.sep)
            if self.append_header:
                self.writer.writerow(['epoch'] + self.keys)
        self.writer.writerow([epoch] + [logs.get(k, 'NA') for k in self.keys])
        self.file.flush()

    def on_train_end(self, logs=None):
        if self.file:
            self.file.close()
            self.file = None


class EarlyStopping(Callback):
    def __init__(self, monitor='val_loss', min_delta=0, patience=0, verbose=0, mode='auto'):
        self.monitor, self.patience, self.verbose = monitor, int(patience), verbose
        self.maximise = _mode(mode, monitor) == 'max'
        self.min_delta = abs(float(min_delta))
        self.best = -np.inf if self.maximise else np.inf
        self.wait, self.stopped_epoch = 0, 0

    def on_train_begin(self, logs=None):
        self.wait = 0
        self.best = -np.inf if self.maximise else np.inf

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        improved = cur - self.min_delta > self.best if self.maximise else cur + self.min_delta < self.best
        if improved:
            self.best, self.wait = cur, 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.model.stop_training = True
                if self.verbose:
                    print('Epoch %05d: early stopping' % (epoch + 1))


class ReduceLROnPlateau(Callback):
    def __init__(self, monitor='val_loss', factor=0.1, patience=10, verbose=0, mode='auto', epsilon=1e-4, cooldown=0, min_lr=0):
        if factor >= 1.0:
            raise ValueError('ReduceLROnPlateau does not support a factor >= 1.0.')
        self.monitor, self.factor, self.patience, self.verbose = monitor, float(factor), int(patience), verbose
        self.epsilon, self.cooldown, self.min_lr = float(epsilon), int(cooldown), float(min_lr)
        self.maximise = _mode(mode, monitor) == 'max'
        self.best = -np.inf if self.maximise else np.inf
        self.wait, self.cooldown_counter = 0, 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        improved = cur > self.best + self.epsilon if self.maximise else cur < self.best - self.epsilon
        if improved:
            self.best, self.wait = cur, 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old = float(self.model.optimizer.lr)
                if old > self.min_lr:
                    self.model.optimizer.lr = max(old * self.factor, self.min_lr)
                    if self.verbose:
                        print('Epoch %05d: ReduceLROnPlateau reducing learning rate to %s.' % (epoch + 1, self.model.optimizer.lr))
                    self.cooldown_counter, self.wait = self.cooldown, 0
