"""Box-format conversion and IoU (oracle, float64 NumPy).

Restates ``bounding_box_utils/bounding_box_utils.py``:
  * ``convert_coordinates``  :24-87
  * ``intersection_area_``   :226-280
  * ``iou``                  :283-383
Pinned against the real reference by tests/golden/make_golden.py.
"""
import numpy as np

_BORDER_D = {'half': 0, 'include': 1, 'exclude': -1}

_CONVERSIONS = ('minmax2centroids', 'centroids2minmax', 'corners2centroids',
                'centroids2corners', 'minmax2corners', 'corners2minmax')


def convert_coordinates(tensor, start_index, conversion, border_pixels='half'):
    """bounding_box_utils.py:24-87.

    The output container is always float64 (``np.copy(t).astype(np.float)``, :60)
    but the arithmetic on the right-hand side runs in the INPUT dtype and is only
    then stored into the f64 container -- a float32 input keeps f32-rounded values.
    """
    if conversion not in _CONVERSIONS:
        raise ValueError("Unexpected conversion value. Supported values are 'minmax2centroids', "
                         "'centroids2minmax', 'corners2centroids', 'centroids2corners', "
                         "'minmax2corners', and 'corners2minmax'.")
    d = _BORDER_D[border_pixels]
    src = np.asarray(tensor)
    out = np.array(src, dtype=np.float64, copy=True)
    i = start_index
    a, b, c, e = src[..., i], src[..., i + 1], src[..., i + 2], src[..., i + 3]
    if conversion == 'minmax2centroids':          # (xmin,xmax,ymin,ymax) -> (cx,cy,w,h)
        out[..., i] = (a + b) / 2.0
        out[..., i + 1] = (c + e) / 2.0
        out[..., i + 2] = b - a + d
        out[..., i + 3] = e - c + d
    elif conversion == 'centroids2minmax':        # (cx,cy,w,h) -> (xmin,xmax,ymin,ymax)
        out[..., i] = a - c / 2.0
        out[..., i + 1] = a + c / 2.0
        out[..., i + 2] = b - e / 2.0
        out[..., i + 3] = b + e / 2.0
    elif conversion == 'corners2centroids':       # (xmin,ymin,xmax,ymax) -> (cx,cy,w,h)
        out[..., i] = (a + c) / 2.0
        out[..., i + 1] = (b + e) / 2.0
        out[..., i + 2] = c - a + d
        out[..., i + 3] = e - b + d
    elif conversion == 'centroids2corners':       # (cx,cy,w,h) -> (xmin,ymin,xmax,ymax)
        out[..., i] = a - c / 2.0
        out[..., i + 1] = b - e / 2.0
        out[..., i + 2] = a + c / 2.0
        out[..., i + 3] = b + e / 2.0
    else:                                         # swap the two middle slots
        out[..., i + 1] = c
        out[..., i + 2] = b
    return out


def _axes(coords):
    if coords == 'corners':
        return 0, 1, 2, 3      # xmin, ymin, xmax, ymax
    if coords == 'minmax':
        return 0, 2, 1, 3      # xmin, ymin, xmax, ymax positions in (xmin,xmax,ymin,ymax)
    raise ValueError("Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.")


def iou(boxes1, boxes2, coords='centroids', mode='outer_product', border_pixels='half'):
    """bounding_box_utils.py:283-383 (+ intersection_area_ :226-280).

    Quirk kept on purpose: ``iou`` calls
    ``intersection_area_`` WITHOUT forwarding ``border_pixels`` (:345), so the
    intersection always uses d=0 while the two box areas use the requested d.
    No epsilon: 0/0 gives NaN (:383).
    """
    b1 = np.asarray(boxes1)
    b2 = np.asarray(boxes2)
    if b1.ndim > 2:
        raise ValueError("boxes1 must have rank either 1 or 2, but has rank {}.".format(b1.ndim))
    if b2.ndim > 2:
        raise ValueError("boxes2 must have rank either 1 or 2, but has rank {}.".format(b2.ndim))
    if b1.ndim == 1:
        b1 = b1[None, :]
    if b2.ndim == 1:
        b2 = b2[None, :]
    if not (b1.shape[1] == b2.shape[1] == 4):
        raise ValueError("All boxes must consist of 4 coordinates, but the boxes in `boxes1` and `boxes2` "
                         "have {} and {} coordinates, respectively.".format(b1.shape[1], b2.shape[1]))
    if mode not in ('outer_product', 'element-wise'):
        raise ValueError("`mode` must be one of 'outer_product' and 'element-wise', but got '{}'.".format(mode))
    if coords == 'centroids':
        b1 = convert_coordinates(b1, 0, 'centroids2corners')
        b2 = convert_coordinates(b2, 0, 'centroids2corners')
        coords = 'corners'
    x0, y0, x1, y1 = _axes(coords)
    d = _BORDER_D[border_pixels]

    if mode == 'outer_product':
        p = b1[:, None, :]
        q = b2[None, :, :]
    else:
        p, q = b1, b2
    # intersection: d is NOT applied here (reference quirk, :345)
    iw = np.maximum(0, np.minimum(p[..., x1], q[..., x1]) - np.maximum(p[..., x0], q[..., x0]))
    ih = np.maximum(0, np.minimum(p[..., y1], q[..., y1]) - np.maximum(p[..., y0], q[..., y0]))
    inter = iw * ih
    area_p = (p[..., x1] - p[..., x0] + d) * (p[..., y1] - p[..., y0] + d)
    area_q = (q[..., x1] - q[..., x0] + d) * (q[..., y1] - q[..., y0] + d)
    with np.errstate(divide='ignore', invalid='ignore'):
        return inter / (area_p + area_q - inter)
