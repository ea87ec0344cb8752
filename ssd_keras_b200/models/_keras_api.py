"""The training side of the Keras ``Model`` surface the reference's notebooks use (``ssd300_training.ipynb:153-173, 437-448``,
``ssd7_training.ipynb:153-156, 330-340``): ``compile(optimizer, loss)``, ``train_on_batch``, ``test_on_batch``,
``fit_generator(generator, steps_per_epoch, epochs, callbacks, validation_data, validation_steps, initial_epoch)`` and
``evaluate_generator`` -- thin host code over ``SSDTrainer`` (encode-free: the generator yields ``(images, y_encoded)`` exactly as
the reference's ``DataGenerator.generate(..., label_encoder=SSDInputEncoder, returns={'processed_images', 'encoded_labels'})``)."""
import numpy as np


class History:
    """keras.callbacks.History: ``history`` maps a metric name to its per-epoch values, ``epoch`` lists the epochs run."""

    def __init__(self):
        self.epoch, self.history = [], {}


def _call(callbacks, name, *args):
    for cb in callbacks:
        fn = getattr(cb, name, None)
        if callable(fn):
            fn(*args)


class KerasTrainingMixin:
    optimizer = None
    _compiled = None
    _fit_trainer = None
    stop_training = False

    def compile(self, optimizer, loss=None, **kwargs):
        """``optimizer``: ``ssd_keras_b200.optimizers.SGD`` / ``Adam`` (or any object with their attributes, or 'sgd' / 'adam');
        ``loss``: ``SSDLoss(...).compute_loss`` (the bound method, as in the reference) or the ``SSDLoss`` object."""
        from .. import optimizers
        from ..keras_loss_function.keras_ssd_loss import SSDLoss
        if kwargs:
            raise TypeError('compile() got unsupported arguments: %s' % sorted(kwargs))
        if isinstance(optimizer, str):
            if optimizer.lower() not in ('sgd', 'adam'):
                raise ValueError("optimizer must be 'sgd' or 'adam' (the two the reference trains with), got %r" % (optimizer,))
            optimizer = optimizers.SGD() if optimizer.lower() == 'sgd' else optimizers.Adam()
        kind = getattr(optimizer, 'kind', type(optimizer).__name__.lower())
        if kind not in ('sgd', 'adam') or not hasattr(optimizer, 'lr'):
            raise ValueError('unsupported optimizer %r: SGD (lr, momentum) and Adam (lr, beta_1, beta_2, epsilon) are implemented' % (optimizer,))
        owner = getattr(loss, '__self__', loss)
        if not isinstance(owner, SSDLoss):
            raise ValueError('loss must be SSDLoss(...).compute_loss (keras_loss_function/keras_ssd_loss.py:98): the training step '
                             'computes that loss and its gradient on the device')
        self.optimizer, self._compiled = optimizer, (kind, owner)
        self._fit_trainer = None

    def _trainer_for(self, batch):
        from ..training import SSDTrainer
        if self._compiled is None:
            raise RuntimeError('You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.')
        kind, lossobj = self._compiled
        t = self._fit_trainer
        if t is None or t.batch != batch:
            if t is not None:
                raise ValueError('this model has been training on batches of %d images; a batch of %d would need its own optimizer '
                                 'state (the reference generator yields constant-size batches)' % (t.batch, batch))
            o = self.optimizer
            t = SSDTrainer(self, batch, lr=o.lr, momentum=getattr(o, 'momentum', 0.0), neg_pos_ratio=lossobj.neg_pos_ratio,
                           n_neg_min=lossobj.n_neg_min, alpha=lossobj.alpha, optimizer=kind, beta_1=getattr(o, 'beta_1', 0.9),
                           beta_2=getattr(o, 'beta_2', 0.999), epsilon=getattr(o, 'epsilon', 1e-8))
            self._fit_trainer = t
        t.lr = float(self.optimizer.lr)                  # a callback (or the caller) may have changed the learning rate
        return t

    @staticmethod
    def _to_cuda(a):
        import torch
        if torch.is_tensor(a):
            return a.to(device='cuda', dtype=torch.float32).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()

    def train_on_batch(self, x, y):
        """One optimizer step on ``(x, y)``; returns the scalar Keras reports: the mean over the batch of ``compute_loss``."""
        if self._compiled is None:
            raise RuntimeError('You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.')
        xd, yd = self._to_cuda(x), self._to_cuda(y)
        loss = self._trainer_for(int(xd.shape[0])).train_on_batch(xd, yd)
        return float(loss.mean().item())

    def test_on_batch(self, x, y):
        """The same scalar without an update, in the inference phase (moving BatchNormalization statistics), like Keras."""
        if self._compiled is None:
            raise RuntimeError('You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.')
        xd, yd = self._to_cuda(x), self._to_cuda(y)
        if self.decoder is not None:
            raise ValueError("test_on_batch needs the raw predictions: build the model with mode='training'")
        return float(self._compiled[1].compute_loss(yd, self.forward_device(xd)).mean().item())

    def evaluate_generator(self, generator, steps):
        it = iter(generator)
        tot, n = 0.0, 0
        for _ in range(int(steps)):
            x, y = next(it)[:2]
            tot += self.test_on_batch(x, y) * len(x)
            n += len(x)
        return tot / max(n, 1)

    def fit_generator(self, generator, steps_per_epoch, epochs=1, verbose=0, callbacks=None, validation_data=None,
                      validation_steps=None, initial_epoch=0):
        """Keras' loop: ``epochs - initial_epoch`` epochs of ``steps_per_epoch`` batches from ``generator`` (an endless iterator of
        ``(images, y_encoded)``), the epoch's mean loss in ``history['loss']``, the validation loss in ``history['val_loss']``.
        Callbacks are duck-typed: ``set_model``, ``on_train_begin/end(logs)``, ``on_epoch_begin/end(epoch, logs)``,
        ``on_batch_end(batch, logs)`` are called where present; a callback may set ``model.optimizer.lr`` or ``model.stop_training``."""
        callbacks = list(callbacks or [])
        hist = History()
        _call(callbacks, 'set_model', self)
        self.stop_training = False
        _call(callbacks, 'on_train_begin', {})
        it = iter(generator)
        for epoch in range(int(initial_epoch), int(epochs)):
            _call(callbacks, 'on_epoch_begin', epoch, {})
            tot, n = 0.0, 0
            for step in range(int(steps_per_epoch)):
                x, y = next(it)[:2]
                l = self.train_on_batch(x, y)
                tot += l * len(x)
                n += len(x)
                _call(callbacks, 'on_batch_end', step, {'loss': l, 'size': len(x)})
                if not np.isfinite(l):
                    self.stop_training = any(type(cb).__name__ == 'TerminateOnNaN' for cb in callbacks) or self.stop_training
                if self.stop_training:
                    break
            logs = {'loss': tot / max(n, 1)}
            if validation_data is not None:
                if validation_steps is None:
                    raise ValueError('`validation_steps` must be given when `validation_data` is a generator')
                logs['val_loss'] = self.evaluate_generator(validation_data, validation_steps)
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            if verbose:
                print('Epoch %d/%d - ' % (epoch + 1, epochs) + ' - '.join('%s: %.4f' % kv for kv in logs.items()), flush=True)
            _call(callbacks, 'on_epoch_end', epoch, logs)
            if self.stop_training:
                break
        _call(callbacks, 'on_train_end', {})
        self.history = hist
        return hist
