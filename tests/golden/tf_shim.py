"""A NumPy stand-in for the handful of TensorFlow 1.x / Keras 2.x symbols the reference's loss and layer code calls, so that
THE REFERENCE'S OWN SOURCE (keras_loss_function/keras_ssd_loss.py, keras_layers/keras_layer_*.py) can be executed eagerly in
this container, where TensorFlow cannot be installed, to produce golden vectors (make_tf_golden.py).

Test infrastructure only.  What this pins: everything the reference's Python code does around the primitives -- masks,
normalisation, thresholds, class loop order, padding, top-k, neutral boxes, the n_neg_min / count_nonzero logic.  What it
assumes: the semantics of the primitives themselves, restated here from TensorFlow's documentation:
  * float32 arithmetic for float tensors, int32 for Python ints (tf.constant defaults);
  * tf.nn.top_k: descending values, ties -> lower index first;
  * tf.image.non_max_suppression: greedy by descending score (ties -> lower index), a candidate is dropped iff its IoU with an
    already selected box is > iou_threshold, boxes with non-positive area have IoU 0, stops at max_output_size;
  * tf.to_int32 truncates toward zero; tf.pad / tf.gather / tf.boolean_mask / tf.scatter_nd as documented;
  * K.l2_normalize(x, axis) = x * rsqrt(max(sum(x^2, axis), 1e-12)).
"""
import collections
import sys
import types

import numpy as np

F = np.float32


def _c(x):
    """Python scalars take TensorFlow's default dtypes."""
    if isinstance(x, (bool, np.bool_)):
        return np.bool_(x)
    if isinstance(x, int):
        return np.int32(x)
    if isinstance(x, float):
        return F(x)
    a = np.asarray(x)
    if a.dtype == np.float64:
        a = a.astype(F)
    if a.dtype == np.int64:
        a = a.astype(np.int32)
    return a


TopK = collections.namedtuple('TopK', ['values', 'indices'])


def _nms(boxes, scores, max_output_size, iou_threshold=0.5, name=None):
    boxes = np.asarray(boxes, dtype=F)
    scores = np.asarray(scores, dtype=F)
    n = boxes.shape[0]
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))
    y0 = np.minimum(boxes[:, 0], boxes[:, 2]); y1 = np.maximum(boxes[:, 0], boxes[:, 2])
    x0 = np.minimum(boxes[:, 1], boxes[:, 3]); x1 = np.maximum(boxes[:, 1], boxes[:, 3])
    area = ((y1 - y0) * (x1 - x0)).astype(F)
    keep = []
    thr = F(iou_threshold)
    for i in order:
        if len(keep) >= int(max_output_size):
            break
        ok = True
        for j in keep:
            if area[i] <= 0 or area[j] <= 0:
                continue
            ih = max(F(min(y1[i], y1[j]) - max(y0[i], y0[j])), F(0))
            iw = max(F(min(x1[i], x1[j]) - max(x0[i], x0[j])), F(0))
            inter = F(ih * iw)
            if F(inter / F(F(area[i] + area[j]) - inter)) > thr:
                ok = False
                break
        if ok:
            keep.append(int(i))
    return np.asarray(keep, dtype=np.int32)


def _top_k(x, k=1, sorted=True, name=None):
    x = np.asarray(x)
    idx = np.lexsort((np.arange(x.shape[0]), -x.astype(np.float64)))[:int(k)].astype(np.int32)
    return TopK(x[idx], idx)


def _scatter_nd(indices, updates, shape, name=None):
    out = np.zeros(tuple(int(s) for s in np.atleast_1d(shape)), dtype=np.asarray(updates).dtype)
    np.add.at(out, tuple(np.asarray(indices).T), updates)
    return out


def _constant(value, dtype=None, shape=None, name=None):
    v = _c(value)
    if dtype is not None:
        v = np.asarray(v).astype(dtype)
    if shape is not None:
        v = np.full(shape, v, dtype=np.asarray(v).dtype)
    return v


def _map_fn(fn, elems, dtype=None, parallel_iterations=None, back_prop=True, swap_memory=False, infer_shape=True, name=None):
    return np.stack([np.asarray(fn(e)) for e in elems], axis=0)


def _pad(tensor, paddings, mode='CONSTANT', constant_values=0.0, name=None):
    pw = [(int(a), int(b)) for a, b in paddings]
    return np.pad(np.asarray(tensor), pw, mode='constant', constant_values=constant_values)


def _reduce(fn):
    def red(x, axis=None, keepdims=False, name=None, keep_dims=None):      # keep_dims: the TensorFlow 1.x spelling
        x = np.asarray(x)
        keepdims = bool(keep_dims) if keep_dims is not None else keepdims
        return fn(x, axis=axis, keepdims=keepdims).astype(x.dtype) if x.dtype.kind == 'f' else fn(x, axis=axis, keepdims=keepdims)
    return red


def make_tf():
    tf = types.ModuleType('tensorflow')
    tf.float32, tf.int32, tf.int64, tf.bool = np.float32, np.int32, np.int64, np.bool_
    tf.constant = _constant
    tf.abs = lambda x, name=None: np.abs(x)
    tf.exp = lambda x, name=None: np.exp(np.asarray(x, dtype=F)).astype(F)
    tf.log = lambda x, name=None: np.log(np.asarray(x, dtype=F)).astype(F)
    tf.less = lambda a, b, name=None: np.less(a, _c(b))
    tf.equal = lambda a, b, name=None: np.equal(a, b)
    tf.not_equal = lambda a, b, name=None: np.not_equal(a, b)
    tf.greater_equal = lambda a, b, name=None: np.greater_equal(a, b)
    tf.where = lambda c, a, b, name=None: np.where(c, a, b).astype(np.asarray(a).dtype)
    tf.maximum = lambda a, b, name=None: np.maximum(_c(a), _c(b))
    tf.minimum = lambda a, b, name=None: np.minimum(_c(a), _c(b))
    tf.reduce_sum = _reduce(np.sum)
    tf.reduce_max = _reduce(np.max)
    tf.argmax = lambda x, axis=None, name=None: np.argmax(x, axis=axis).astype(np.int64)
    tf.count_nonzero = lambda x, axis=None, dtype=np.int64, name=None: np.asarray(np.count_nonzero(x, axis=axis)).astype(dtype)
    tf.to_float = lambda x, name=None: np.asarray(x).astype(F)
    tf.to_int32 = lambda x, name=None: np.trunc(np.asarray(x)).astype(np.int32)
    tf.shape = lambda x, name=None: np.asarray(np.asarray(x).shape, dtype=np.int32)
    tf.size = lambda x, name=None: np.int32(np.asarray(x).size)
    tf.zeros = lambda shape, dtype=F, name=None: np.zeros([int(s) for s in np.atleast_1d(shape)], dtype=dtype)
    tf.ones_like = lambda x, dtype=None, name=None: np.ones_like(x, dtype=dtype)
    tf.fill = lambda dims, value, name=None: np.full([int(s) for s in dims], value, dtype=np.asarray(value).dtype)
    tf.range = lambda *a, **k: np.arange(*[int(v) for v in a], dtype=np.int32)
    tf.reshape = lambda tensor, shape, name=None: np.reshape(tensor, [int(s) for s in np.atleast_1d(shape)])
    tf.expand_dims = lambda input, axis=None, name=None: np.expand_dims(input, axis)
    tf.concat = lambda values, axis, name=None: np.concatenate([np.asarray(v) for v in values], axis=axis)
    tf.gather = lambda params, indices, axis=0, name=None: np.take(params, np.asarray(indices, dtype=np.int64), axis=axis)
    tf.boolean_mask = lambda tensor, mask, name=None: np.asarray(tensor)[np.asarray(mask, dtype=bool)]
    tf.pad = _pad
    tf.cond = lambda pred, true_fn, false_fn, name=None: true_fn() if bool(np.asarray(pred)) else false_fn()
    tf.map_fn = _map_fn
    tf.scatter_nd = _scatter_nd
    tf.nn = types.SimpleNamespace(top_k=_top_k)
    tf.image = types.SimpleNamespace(non_max_suppression=_nms)
    return tf


class _T(np.ndarray):
    """ndarray that accepts attributes (Keras attaches `_keras_shape` to tensors)."""


def keras_tensor(a):
    t = np.asarray(a, dtype=F).view(_T)
    t._keras_shape = tuple(t.shape)
    return t


def make_keras():
    keras = types.ModuleType('keras')
    K = types.ModuleType('keras.backend')
    K.backend = lambda: 'tensorflow'
    K.image_dim_ordering = lambda: 'tf'
    K.variable = lambda value, dtype=None, name=None: np.asarray(value, dtype=F)
    K.constant = lambda value, dtype=None, shape=None, name=None: np.asarray(value, dtype=dtype or F)
    K.shape = lambda x: np.asarray(np.asarray(x).shape, dtype=np.int32)
    K.tile = lambda x, n: np.tile(x, [int(v) for v in n])
    K.expand_dims = lambda x, axis=-1: np.expand_dims(x, axis)
    K.concatenate = lambda tensors, axis=-1: np.concatenate(tensors, axis=axis)
    K.stack = lambda xs, axis=0: np.stack([np.asarray(v) for v in xs], axis=axis)

    def l2_normalize(x, axis=None):
        x = np.asarray(x, dtype=F)
        ss = np.sum(x * x, axis=axis, keepdims=True, dtype=F)
        return (x * (F(1) / np.sqrt(np.maximum(ss, F(1e-12))))).astype(F)
    K.l2_normalize = l2_normalize

    engine = types.ModuleType('keras.engine')
    topology = types.ModuleType('keras.engine.topology')

    class InputSpec(object):
        def __init__(self, **kwargs):
            self.__dict__.update(kwargs)

    class Layer(object):
        def __init__(self, **kwargs):
            self.name = kwargs.get('name', self.__class__.__name__.lower())
            self.built = False
            self.trainable_weights = []

        def build(self, input_shape):
            self.built = True

        def get_config(self):
            return {}

        def __call__(self, x, **kw):
            if not self.built:
                self.build(tuple(np.asarray(x).shape))
                W = STATE.get('weights') or {}
                if hasattr(self, 'gamma') and (self.name + '/gamma') in W:       # L2Normalization: trained scale by layer name
                    self.gamma = np.asarray(W[self.name + '/gamma'], dtype=F)
            return keras_tensor(np.asarray(self.call(x, **kw), dtype=F))

    topology.InputSpec, topology.Layer = InputSpec, Layer
    engine.topology = topology
    keras.backend, keras.engine = K, engine
    return {'keras': keras, 'keras.backend': K, 'keras.engine': engine, 'keras.engine.topology': topology}


# ---------------------------------------------------------------------------------------------------------------
# Eager stand-ins for the Keras functional-API layers the reference's model builders use (models/keras_ssd300.py etc.):
# `Input` returns the concrete image batch stored in STATE['input'], every layer call computes its output immediately
# (float32, NHWC; convolutions through torch on the CPU), weights come from STATE['weights'] by Keras layer name.
# ---------------------------------------------------------------------------------------------------------------
STATE = {'input': None, 'weights': None}


def _same_pad(size, k, s, d=1):
    """TensorFlow 'SAME': total padding so that out = ceil(size / s); the extra pixel goes to the end."""
    ke = (k - 1) * d + 1
    out = -(-size // s)
    total = max((out - 1) * s + ke - size, 0)
    return total // 2, total - total // 2


def _pair2(v):
    return (int(v), int(v)) if np.isscalar(v) else (int(v[0]), int(v[1]))


def make_keras_layers():
    import torch
    import torch.nn.functional as Fn
    L = types.ModuleType('keras.layers')

    class _Base(object):
        def __init__(self, name=None, **kw):
            self.name = name

    def Input(shape=None, **kw):
        x = keras_tensor(STATE['input'])
        assert tuple(x.shape[1:]) == tuple(shape), (x.shape, shape)
        return x

    class Lambda(_Base):
        def __init__(self, function, output_shape=None, name=None, **kw):
            _Base.__init__(self, name); self.fn = function

        def __call__(self, x):
            return keras_tensor(np.asarray(self.fn(np.asarray(x)), dtype=F))   # Keras casts NumPy constants to the tensor's dtype

    class Conv2D(_Base):
        def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', dilation_rate=(1, 1), activation=None,
                     kernel_initializer=None, kernel_regularizer=None, name=None, **kw):
            _Base.__init__(self, name)
            self.filters, self.k, self.s, self.d = filters, _pair2(kernel_size), _pair2(strides), _pair2(dilation_rate)
            self.padding, self.activation = padding, activation

        def __call__(self, x):
            w = np.asarray(STATE['weights'][self.name + '/kernel'], dtype=F)      # HWIO
            b = np.asarray(STATE['weights'][self.name + '/bias'], dtype=F)
            assert w.shape[:2] == self.k and w.shape[3] == self.filters, (self.name, w.shape)
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(x))).permute(0, 3, 1, 2)
            if self.padding == 'same':
                pt, pb = _same_pad(t.shape[2], self.k[0], self.s[0], self.d[0])
                pl, pr = _same_pad(t.shape[3], self.k[1], self.s[1], self.d[1])
                t = Fn.pad(t, (pl, pr, pt, pb))
            else:
                assert self.padding == 'valid'
            y = Fn.conv2d(t, torch.from_numpy(w).permute(3, 2, 0, 1).contiguous(), torch.from_numpy(b), stride=self.s, dilation=self.d)
            if self.activation == 'relu':
                y = torch.relu(y)
            else:
                assert self.activation is None, self.activation
            return keras_tensor(y.permute(0, 2, 3, 1).contiguous().numpy())

    class MaxPooling2D(_Base):
        def __init__(self, pool_size=(2, 2), strides=None, padding='valid', name=None, **kw):
            _Base.__init__(self, name)
            self.k = _pair2(pool_size); self.s = _pair2(strides if strides is not None else pool_size); self.padding = padding

        def __call__(self, x):
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(x))).permute(0, 3, 1, 2)
            if self.padding == 'same':
                pt, pb = _same_pad(t.shape[2], self.k[0], self.s[0])
                pl, pr = _same_pad(t.shape[3], self.k[1], self.s[1])
                t = Fn.pad(t, (pl, pr, pt, pb), value=float('-inf'))
            y = Fn.max_pool2d(t, self.k, self.s)
            return keras_tensor(y.permute(0, 2, 3, 1).contiguous().numpy())

    class ZeroPadding2D(_Base):
        def __init__(self, padding=(1, 1), name=None, **kw):
            _Base.__init__(self, name)
            p = padding
            self.p = ((p, p), (p, p)) if np.isscalar(p) else tuple((q, q) if np.isscalar(q) else tuple(q) for q in p)

        def __call__(self, x):
            (t, b), (l, r) = self.p
            return keras_tensor(np.pad(np.asarray(x), ((0, 0), (t, b), (l, r), (0, 0))))

    class Reshape(_Base):
        def __init__(self, target_shape, name=None, **kw):
            _Base.__init__(self, name); self.shape = tuple(target_shape)

        def __call__(self, x):
            x = np.asarray(x)
            return keras_tensor(x.reshape((x.shape[0],) + self.shape))

    class Concatenate(_Base):
        def __init__(self, axis=-1, name=None, **kw):
            _Base.__init__(self, name); self.axis = axis

        def __call__(self, xs):
            return keras_tensor(np.concatenate([np.asarray(v, dtype=F) for v in xs], axis=self.axis))

    class Activation(_Base):
        def __init__(self, activation, name=None, **kw):
            _Base.__init__(self, name); self.a = activation

        def __call__(self, x):
            assert self.a == 'softmax'
            return keras_tensor(torch.softmax(torch.from_numpy(np.ascontiguousarray(np.asarray(x))), dim=-1).numpy())

    class BatchNormalization(_Base):            # inference phase: moving statistics, Keras default epsilon 1e-3
        def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, name=None, **kw):
            _Base.__init__(self, name); self.eps = epsilon; assert axis in (3, -1)

        def __call__(self, x):
            W = STATE['weights']; n = self.name
            g, b = np.asarray(W[n + '/gamma'], F), np.asarray(W[n + '/beta'], F)
            mu, var = np.asarray(W[n + '/moving_mean'], F), np.asarray(W[n + '/moving_variance'], F)
            return keras_tensor((np.asarray(x) - mu) / np.sqrt(var + F(self.eps)) * g + b)

    class ELU(_Base):
        def __init__(self, alpha=1.0, name=None, **kw):
            _Base.__init__(self, name); self.alpha = F(alpha)

        def __call__(self, x):
            x = np.asarray(x)
            return keras_tensor(np.where(x > 0, x, self.alpha * np.expm1(np.minimum(x, F(0)))))

    for k, v in dict(Input=Input, Lambda=Lambda, Conv2D=Conv2D, MaxPooling2D=MaxPooling2D, ZeroPadding2D=ZeroPadding2D, Reshape=Reshape,
                     Concatenate=Concatenate, Activation=Activation, BatchNormalization=BatchNormalization, ELU=ELU).items():
        setattr(L, k, v)
    models = types.ModuleType('keras.models')

    class Model(object):
        def __init__(self, inputs=None, outputs=None, **kw):
            self.inputs, self.output = inputs, outputs
    models.Model = Model
    reg = types.ModuleType('keras.regularizers')
    reg.l2 = lambda v=0.01: ('l2', v)
    return {'keras.layers': L, 'keras.models': models, 'keras.regularizers': reg}


def install():
    """Register the stand-ins as `tensorflow` / `keras` (refuses to shadow real installations)."""
    for name in ('tensorflow', 'keras'):
        if name in sys.modules and not getattr(sys.modules[name], '_ssd_b200_shim', False):
            raise RuntimeError('%s is really installed: use it instead of the shim' % name)
    tf = make_tf()
    tf._ssd_b200_shim = True
    sys.modules['tensorflow'] = tf
    mods = make_keras()
    mods.update(make_keras_layers())
    for k, v in mods.items():
        v._ssd_b200_shim = True
        sys.modules[k] = v
    for sub in ('layers', 'models', 'regularizers'):
        setattr(mods['keras'], sub, mods['keras.' + sub])
    return tf
