"""The two optimizer configurations the reference's notebooks pass to ``model.compile`` (``keras.optimizers.SGD`` in
``ssd300_training.ipynb:169`` / ``ssd512_training.ipynb``, ``keras.optimizers.Adam`` in ``ssd7_training.ipynb:153``), as plain
parameter holders with Keras' argument names and defaults.  The update itself runs on the device
(``ssdk_train_apply`` / ``ssdk_train_apply_adam`` in csrc/train.cu); options those kernels do not implement raise here."""


class SGD:
    def __init__(self, lr=0.01, momentum=0.0, decay=0.0, nesterov=False):
        if nesterov:
            raise ValueError('SGD(nesterov=True) is not implemented (the reference trains with nesterov=False)')
        if decay:
            raise ValueError('SGD(decay != 0) is not implemented: change `optimizer.lr` between batches instead '
                             '(the reference uses a LearningRateScheduler callback, ssd300_training.ipynb:404-413)')
        self.lr, self.momentum, self.decay, self.nesterov = float(lr), float(momentum), 0.0, False
        self.kind = 'sgd'


class Adam:
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-08, decay=0.0):
        if decay:
            raise ValueError('Adam(decay != 0) is not implemented: change `optimizer.lr` between batches instead')
        self.lr, self.beta_1, self.beta_2, self.epsilon, self.decay = float(lr), float(beta_1), float(beta_2), float(epsilon), 0.0
        self.kind = 'adam'
