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
            return self.call(x, **kw)

    topology.InputSpec, topology.Layer = InputSpec, Layer
    engine.topology = topology
    keras.backend, keras.engine = K, engine
    return {'keras': keras, 'keras.backend': K, 'keras.engine': engine, 'keras.engine.topology': topology}


def install():
    """Register the stand-ins as `tensorflow` / `keras` (refuses to shadow real installations)."""
    for name in ('tensorflow', 'keras'):
        if name in sys.modules and not getattr(sys.modules[name], '_ssd_b200_shim', False):
            raise RuntimeError('%s is really installed: use it instead of the shim' % name)
    tf = make_tf()
    tf._ssd_b200_shim = True
    sys.modules['tensorflow'] = tf
    for k, v in make_keras().items():
        v._ssd_b200_shim = True
        sys.modules[k] = v
    return tf
