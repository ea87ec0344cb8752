"""Anchor ("prior") box generation (oracle, float64 NumPy).

Restates ``SSDInputEncoder.generate_anchor_boxes_for_layer``
(``ssd_encoder_decoder/ssd_input_encoder.py:420-548``), which the
``AnchorBoxes`` Keras layer duplicates (``keras_layers/keras_layer_AnchorBoxes.py:150-253``,
cast to float32 at :252).  Pinned against the real encoder by make_golden.py.
"""
import numpy as np

from .boxes import convert_coordinates


def boxes_per_cell(aspect_ratios, two_boxes_for_ar1):
    """ssd_input_encoder.py:238-252 / models/keras_ssd300.py:220-232."""
    return len(aspect_ratios) + (1 if (1 in aspect_ratios) and two_boxes_for_ar1 else 0)


def _pair(v, default):
    if v is None:
        return default, default
    if isinstance(v, (list, tuple)) and len(v) == 2:
        return v[0], v[1]
    return v, v


def anchor_boxes_for_layer(img_height, img_width, feature_map_size, aspect_ratios, this_scale, next_scale,
                           two_boxes_for_ar1=True, this_steps=None, this_offsets=None, clip_boxes=False,
                           coords='centroids', normalize_coords=True):
    """-> (H, W, n_boxes, 4) float64, ssd_input_encoder.py:456-543."""
    fm_h, fm_w = int(feature_map_size[0]), int(feature_map_size[1])
    size = min(img_height, img_width)                                        # :459
    wh = []
    for ar in aspect_ratios:                                                 # :462-474
        if ar == 1:
            wh.append((this_scale * size, this_scale * size))
            if two_boxes_for_ar1:
                s = np.sqrt(this_scale * next_scale) * size
                wh.append((s, s))
        else:
            wh.append((this_scale * size * np.sqrt(ar), this_scale * size / np.sqrt(ar)))
    wh = np.array(wh, dtype=np.float64)
    n_boxes = len(wh)

    if this_steps is None:                                                   # :480-482
        step_h, step_w = img_height / fm_h, img_width / fm_w
    else:
        step_h, step_w = _pair(this_steps, None)
    off_h, off_w = _pair(this_offsets, 0.5)                                  # :490-500

    cy = np.linspace(off_h * step_h, (off_h + fm_h - 1) * step_h, fm_h)      # :503
    cx = np.linspace(off_w * step_w, (off_w + fm_w - 1) * step_w, fm_w)      # :504
    t = np.zeros((fm_h, fm_w, n_boxes, 4))
    t[..., 0] = cx[None, :, None]
    t[..., 1] = cy[:, None, None]
    t[..., 2] = wh[:, 0]
    t[..., 3] = wh[:, 1]
    t = convert_coordinates(t, 0, 'centroids2corners')                       # :519
    if clip_boxes:                                                           # :522-530
        xs = t[..., [0, 2]]
        xs[xs >= img_width] = img_width - 1
        xs[xs < 0] = 0
        t[..., [0, 2]] = xs
        ys = t[..., [1, 3]]
        ys[ys >= img_height] = img_height - 1
        ys[ys < 0] = 0
        t[..., [1, 3]] = ys
    if normalize_coords:                                                     # :533-535
        t[..., [0, 2]] /= img_width
        t[..., [1, 3]] /= img_height
    if coords == 'centroids':                                                # :540-543
        t = convert_coordinates(t, 0, 'corners2centroids', border_pixels='half')
    elif coords == 'minmax':
        t = convert_coordinates(t, 0, 'corners2minmax', border_pixels='half')
    return t


def resolve_scales(n_layers, min_scale, max_scale, scales):
    """ssd_input_encoder.py:193-196."""
    if scales is None:
        return np.linspace(min_scale, max_scale, n_layers + 1)
    return np.array(scales, dtype=np.float64)


def all_anchors(img_height, img_width, predictor_sizes, scales, aspect_ratios_per_layer,
                two_boxes_for_ar1=True, steps=None, offsets=None, clip_boxes=False,
                coords='centroids', normalize_coords=True):
    """Concatenate every layer's anchors in model order -> (P, 4) float64.

    Prior index = ((y*W + x)*n_boxes + b) within a layer; layers in order
    (ssd_input_encoder.py:590-596, models/keras_ssd300.py:363-402).
    """
    predictor_sizes = np.array(predictor_sizes)
    if predictor_sizes.ndim == 1:
        predictor_sizes = predictor_sizes[None, :]
    n = len(predictor_sizes)
    steps = steps if steps is not None else [None] * n
    offsets = offsets if offsets is not None else [None] * n
    parts = []
    for i in range(n):
        parts.append(anchor_boxes_for_layer(img_height, img_width, predictor_sizes[i], aspect_ratios_per_layer[i],
                                            scales[i], scales[i + 1], two_boxes_for_ar1, steps[i], offsets[i],
                                            clip_boxes, coords, normalize_coords).reshape(-1, 4))
    return np.concatenate(parts, axis=0)
