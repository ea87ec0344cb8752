"""Stand-alone layer calls on CUDA tensors: the Conv2D / MaxPooling2D forward that the model graphs are built from
(``ssdk_conv2d_fwd`` / ``ssdk_maxpool`` in include/ssdk.h; reference: every ``Conv2D`` / ``MaxPooling2D`` of
models/keras_ssd300.py:274-335).  Each call builds, runs and frees a one-layer plan -- meant for tests and interop, the model
builders are the steady-state path."""
import ctypes as C

import numpy as np

from . import _ffi
from .models._graph import same_pad, tf_same_pool_pad

_ACT = {None: _ffi.ACT_NONE, 'linear': _ffi.ACT_NONE, 'relu': _ffi.ACT_RELU, 'elu': _ffi.ACT_ELU}


def _cuda_f32(x):
    import torch
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if x.dim() != 4:
        raise ValueError('expected a (batch, height, width, channels) tensor, got shape %s' % (tuple(x.shape),))
    return x.to(device='cuda', dtype=torch.float32).contiguous()


def conv2d(x, kernel, bias=None, strides=1, padding='same', dilation_rate=1, activation=None, precision='bf16x3'):
    """Keras ``Conv2D(filters, (kh, kw), strides, padding, dilation_rate, activation)`` forward.
    x: (B,H,W,Cin) float32 (CUDA tensor or ndarray); kernel: HWIO ndarray (kh,kw,Cin,Cout); padding: 'same' (stride 1, odd
    kernels), 'valid', or explicit (top, left, bottom, right).  Returns a CUDA tensor (B,Ho,Wo,Cout)."""
    import torch
    x = _cuda_f32(x)
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    if k.ndim != 4 or k.shape[2] != x.shape[3]:
        raise ValueError('kernel must be HWIO with %d input channels, got shape %s' % (x.shape[3], k.shape))
    kh, kw, cin, cout = k.shape
    if activation not in _ACT:
        raise ValueError('unsupported activation %r' % (activation,))
    if padding == 'same':
        if strides != 1 or kh % 2 == 0 or kw % 2 == 0:
            raise ValueError("padding='same' is offered for stride 1 and odd kernels; pass explicit (top, left, bottom, right) pads otherwise")
        pt, pl, pb, pr = same_pad(kh, dilation_rate)[0], same_pad(kw, dilation_rate)[0], same_pad(kh, dilation_rate)[0], same_pad(kw, dilation_rate)[0]
    elif padding == 'valid':
        pt = pl = pb = pr = 0
    else:
        pt, pl, pb, pr = [int(v) for v in padding]
    B, H, W, _ = x.shape
    ho = (H + pt + pb - dilation_rate * (kh - 1) - 1) // strides + 1
    wo = (W + pl + pr - dilation_rate * (kw - 1) - 1) // strides + 1
    if ho <= 0 or wo <= 0:
        raise ValueError('empty convolution output')
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    if b is not None and b.shape != (cout,):
        raise ValueError('bias must have shape (%d,)' % cout)
    y = torch.empty((B, ho, wo, cout), dtype=torch.float32, device=x.device)
    _ffi.check(_ffi.lib().ssdk_conv2d_fwd(_ffi.context(x.device.index), _ffi.dptr(x), B, H, W, cin, _ffi.np_ptr(k, C.c_float),
                                          None if b is None else _ffi.np_ptr(b, C.c_float), cout, kh, kw, int(strides), int(dilation_rate),
                                          pt, pl, pb, pr, _ACT[activation], 0 if precision == 'bf16x3' else 1, _ffi.dptr(y), _ffi.stream_ptr()))
    return y


def max_pool2d(x, pool_size=(2, 2), strides=None, padding='same'):
    """Keras ``MaxPooling2D(pool_size, strides, padding)`` forward with TensorFlow's 'same' rule (extra padding at the end).
    x: (B,H,W,C) float32.  Returns a CUDA tensor (B,Ho,Wo,C)."""
    import torch
    x = _cuda_f32(x)
    kh, kw = (pool_size, pool_size) if np.isscalar(pool_size) else pool_size
    s = kh if strides is None else (strides if np.isscalar(strides) else strides[0])
    B, H, W, Cc = x.shape
    if padding == 'same':
        (pt, pb), (pl, pr) = tf_same_pool_pad(H, kh, s), tf_same_pool_pad(W, kw, s)
    elif padding == 'valid':
        pt = pl = pb = pr = 0
    else:
        pt, pl, pb, pr = [int(v) for v in padding]
    ho, wo = (H + pt + pb - kh) // s + 1, (W + pl + pr - kw) // s + 1
    if ho <= 0 or wo <= 0:
        raise ValueError('empty pooling output')
    y = torch.empty((B, ho, wo, Cc), dtype=torch.float32, device=x.device)
    _ffi.check(_ffi.lib().ssdk_maxpool(_ffi.context(x.device.index), _ffi.dptr(x), B, H, W, Cc, int(kh), int(kw), int(s), pt, pl, pb, pr,
                                       _ffi.dptr(y), _ffi.stream_ptr()))
    return y
