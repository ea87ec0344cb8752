"""``convert_coordinates`` and ``iou`` with the reference signatures (``bounding_box_utils/bounding_box_utils.py:24-87,
283-383``).  ``iou`` runs on the GPU (``ssdk_iou``, float64, same operation order as NumPy, including the quirk that
the intersection ignores ``border_pixels``); ``convert_coordinates`` is host-side index shuffling, as in the reference."""

import numpy as np

from .. import _ffi

_CONV = ('minmax2centroids', 'centroids2minmax', 'corners2centroids', 'centroids2corners', 'minmax2corners', 'corners2minmax')


def convert_coordinates(tensor, start_index, conversion, border_pixels='half'):
    if conversion not in _CONV:
        raise ValueError("Unexpected conversion value. Supported values are 'minmax2centroids', 'centroids2minmax', "
                         "'corners2centroids', 'centroids2corners', 'minmax2corners', and 'corners2minmax'.")
    d = _ffi.BORDER_D[border_pixels]
    t = np.asarray(tensor)
    out = np.array(t, dtype=np.float64, copy=True)
    i = start_index
    s0, s1, s2, s3 = t[..., i], t[..., i + 1], t[..., i + 2], t[..., i + 3]
    if conversion == 'minmax2centroids':
        out[..., i], out[..., i + 1], out[..., i + 2], out[..., i + 3] = (s0 + s1) / 2.0, (s2 + s3) / 2.0, s1 - s0 + d, s3 - s2 + d
    elif conversion == 'centroids2minmax':
        out[..., i], out[..., i + 1], out[..., i + 2], out[..., i + 3] = s0 - s2 / 2.0, s0 + s2 / 2.0, s1 - s3 / 2.0, s1 + s3 / 2.0
    elif conversion == 'corners2centroids':
        out[..., i], out[..., i + 1], out[..., i + 2], out[..., i + 3] = (s0 + s2) / 2.0, (s1 + s3) / 2.0, s2 - s0 + d, s3 - s1 + d
    elif conversion == 'centroids2corners':
        out[..., i], out[..., i + 1], out[..., i + 2], out[..., i + 3] = s0 - s2 / 2.0, s1 - s3 / 2.0, s0 + s2 / 2.0, s1 + s3 / 2.0
    else:
        out[..., i + 1], out[..., i + 2] = s2, s1
    return out


def convert_coordinates2(tensor, start_index, conversion):
    """Reference :89-116: the matrix-product form of the centroids <-> minmax conversion (host side, like the reference)."""
    if conversion == 'minmax2centroids':
        M = np.array([[0.5, 0., -1., 0.], [0.5, 0., 1., 0.], [0., 0.5, 0., -1.], [0., 0.5, 0., 1.]])
    elif conversion == 'centroids2minmax':
        M = np.array([[1., 1., 0., 0.], [0., 0., 1., 1.], [-0.5, 0.5, 0., 0.], [0., 0., -0.5, 0.5]])
    else:
        raise ValueError("Unexpected conversion value. Supported values are 'minmax2centroids' and 'centroids2minmax'.")
    out = np.array(tensor, dtype=np.float64, copy=True)
    i = start_index
    out[..., i:i + 4] = np.dot(out[..., i:i + 4], M)
    return out


def iou(boxes1, boxes2, coords='centroids', mode='outer_product', border_pixels='half'):
    import torch
    b1, b2 = np.asarray(boxes1, dtype=np.float64), np.asarray(boxes2, dtype=np.float64)
    if b1.ndim > 2:
        raise ValueError("boxes1 must have rank either 1 or 2, but has rank {}.".format(b1.ndim))
    if b2.ndim > 2:
        raise ValueError("boxes2 must have rank either 1 or 2, but has rank {}.".format(b2.ndim))
    if b1.ndim == 1:
        b1 = b1[None]
    if b2.ndim == 1:
        b2 = b2[None]
    if not (b1.shape[1] == b2.shape[1] == 4):
        raise ValueError("All boxes must consist of 4 coordinates, but the boxes in `boxes1` and `boxes2` have {} and {} "
                         "coordinates, respectively.".format(b1.shape[1], b2.shape[1]))
    if mode not in ('outer_product', 'element-wise'):
        raise ValueError("`mode` must be one of 'outer_product' and 'element-wise', but got '{}'.".format(mode))
    if coords not in _ffi.COORDS:
        raise ValueError("Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.")
    m, n = b1.shape[0], b2.shape[0]
    d1 = torch.from_numpy(np.ascontiguousarray(b1)).cuda()
    d2 = torch.from_numpy(np.ascontiguousarray(b2)).cuda()
    elem = mode == 'element-wise'
    out = torch.empty((max(m, n),) if elem else (m, n), dtype=torch.float64, device='cuda')
    _ffi.check(_ffi.lib().ssdk_iou(_ffi.context(), _ffi.dptr(d1), m, _ffi.dptr(d2), n, _ffi.COORDS[coords],
                                   _ffi.BORDER_D[border_pixels], 1 if elem else 0, _ffi.dptr(out), _ffi.stream_ptr()))
    return out.cpu().numpy()


def _intersection_t(b1, b2, coords, elementwise, d):
    """(m,4), (n,4) float64 tensors (any device) in 'corners' or 'minmax' order -> (m,n) or (max(m,n),) intersection areas
    (reference ``intersection_area_``, bounding_box_utils.py:226-280)."""
    import torch
    if coords == 'corners':
        lo, hi = [0, 1], [2, 3]
    else:                                                   # minmax: xmin, xmax, ymin, ymax
        lo, hi = [0, 2], [1, 3]
    if elementwise:
        mn = torch.maximum(b1[:, lo], b2[:, lo]); mx = torch.minimum(b1[:, hi], b2[:, hi])
        side = torch.clamp(mx - mn + d, min=0)
        return side[:, 0] * side[:, 1]
    mn = torch.maximum(b1[:, None, lo], b2[None, :, lo]); mx = torch.minimum(b1[:, None, hi], b2[None, :, hi])
    side = torch.clamp(mx - mn + d, min=0)
    return side[:, :, 0] * side[:, :, 1]


def intersection_area_(boxes1, boxes2, coords='corners', mode='outer_product', border_pixels='half'):
    """Reference :226-280 (no checks).  A few float64 tensor operations on the GPU; the encoder itself never calls this --
    its IoUs are computed inside ``ssdk_encode``."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('ssd_keras_b200.bounding_box_utils needs a CUDA device (there is no CPU fallback)')
    b1 = torch.from_numpy(np.ascontiguousarray(np.asarray(boxes1, dtype=np.float64))).cuda()
    b2 = torch.from_numpy(np.ascontiguousarray(np.asarray(boxes2, dtype=np.float64))).cuda()
    return _intersection_t(b1, b2, coords, mode == 'element-wise', float(_ffi.BORDER_D[border_pixels])).cpu().numpy()


def intersection_area(boxes1, boxes2, coords='centroids', mode='outer_product', border_pixels='half'):
    """Reference :89-196: the checks and conversions of ``iou``, then ``intersection_area_``."""
    b1, b2 = np.asarray(boxes1, dtype=np.float64), np.asarray(boxes2, dtype=np.float64)
    if b1.ndim > 2:
        raise ValueError("boxes1 must have rank either 1 or 2, but has rank {}.".format(b1.ndim))
    if b2.ndim > 2:
        raise ValueError("boxes2 must have rank either 1 or 2, but has rank {}.".format(b2.ndim))
    if b1.ndim == 1:
        b1 = b1[None]
    if b2.ndim == 1:
        b2 = b2[None]
    if not (b1.shape[1] == b2.shape[1] == 4):
        raise ValueError("All boxes must consist of 4 coordinates, but the boxes in `boxes1` and `boxes2` have {} and {} "
                         "coordinates, respectively.".format(b1.shape[1], b2.shape[1]))
    if mode not in ('outer_product', 'element-wise'):
        raise ValueError("`mode` must be one of 'outer_product' and 'element-wise', but got '{}'.".format(mode))
    if coords == 'centroids':                                # :136-139
        b1 = convert_coordinates(b1, start_index=0, conversion='centroids2corners')
        b2 = convert_coordinates(b2, start_index=0, conversion='centroids2corners')
        coords = 'corners'
    elif coords not in ('minmax', 'corners'):
        raise ValueError("Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.")
    return intersection_area_(b1, b2, coords=coords, mode=mode, border_pixels=border_pixels)
