"""Ground-truth encoder (oracle, float64 NumPy).

Restates ``SSDInputEncoder.__init__`` / ``__call__`` / ``generate_encoding_template``
(``ssd_encoder_decoder/ssd_input_encoder.py:36-275, 277-418, 550-611``).
Pinned against the real encoder by tests/golden/make_golden.py.
"""
import numpy as np

from .anchors import all_anchors, boxes_per_cell, resolve_scales
from .boxes import convert_coordinates, iou
from .matching import match_bipartite_greedy, match_multi


class DegenerateBoxError(Exception):
    """ssd_input_encoder.py:613-617."""


class OracleEncoder:
    """Same constructor arguments as the reference ``SSDInputEncoder`` (:36-57)."""

    def __init__(self, img_height, img_width, n_classes, predictor_sizes, min_scale=0.1, max_scale=0.9,
                 scales=None, aspect_ratios_global=(0.5, 1.0, 2.0), aspect_ratios_per_layer=None,
                 two_boxes_for_ar1=True, steps=None, offsets=None, clip_boxes=False,
                 variances=(0.1, 0.1, 0.2, 0.2), matching_type='multi', pos_iou_threshold=0.5,
                 neg_iou_limit=0.3, border_pixels='half', coords='centroids', normalize_coords=True,
                 background_id=0):
        predictor_sizes = np.array(predictor_sizes)
        if predictor_sizes.ndim == 1:
            predictor_sizes = predictor_sizes[None, :]
        n_layers = predictor_sizes.shape[0]
        self.img_height, self.img_width = img_height, img_width
        self.n_classes = n_classes + 1                                        # :188
        self.predictor_sizes = predictor_sizes
        self.scales = resolve_scales(n_layers, min_scale, max_scale, scales)
        if aspect_ratios_per_layer is None:
            self.aspect_ratios = [list(aspect_ratios_global)] * n_layers
        else:
            self.aspect_ratios = [list(a) for a in aspect_ratios_per_layer]
        self.two_boxes_for_ar1 = two_boxes_for_ar1
        self.steps, self.offsets = steps, offsets
        self.clip_boxes = clip_boxes
        self.variances = np.array(variances, dtype=np.float64)
        self.matching_type = matching_type
        self.pos_iou_threshold = pos_iou_threshold
        self.neg_iou_limit = neg_iou_limit
        self.border_pixels = border_pixels
        self.coords = coords
        self.normalize_coords = normalize_coords
        self.background_id = background_id
        self.n_boxes = [boxes_per_cell(a, two_boxes_for_ar1) for a in self.aspect_ratios]
        self.anchors = all_anchors(img_height, img_width, predictor_sizes, self.scales, self.aspect_ratios,
                                   two_boxes_for_ar1, steps, offsets, clip_boxes, coords, normalize_coords)

    def template(self, batch_size):
        """generate_encoding_template, :550-611: [zeros(C) | anchors | anchors | variances]."""
        P = self.anchors.shape[0]
        row = np.concatenate([np.zeros((P, self.n_classes)), self.anchors, self.anchors,
                              np.broadcast_to(self.variances, (P, 4))], axis=1)
        return np.tile(row[None], (batch_size, 1, 1))

    def __call__(self, ground_truth_labels, diagnostics=False, return_matches=False):
        """:277-418.  ``return_matches`` additionally returns, per image, the per-anchor
        matched gt index (-1 = none) and a neutral flag -- the integer outputs the CUDA
        kernel is compared against bit-exactly."""
        B = len(ground_truth_labels)
        C = self.n_classes
        y = self.template(B)
        y[:, :, self.background_id] = 1
        P = y.shape[1]
        eye = np.eye(C)
        match_idx = np.full((B, P), -1, dtype=np.int32)
        neutral = np.zeros((B, P), dtype=bool)
        for i in range(B):
            gt = np.asarray(ground_truth_labels[i])
            if gt.size == 0:                                                  # :329
                continue
            lab = gt.astype(np.float64)
            if np.any(lab[:, 3] - lab[:, 1] <= 0) or np.any(lab[:, 4] - lab[:, 2] <= 0):   # :333-336
                raise DegenerateBoxError("SSDInputEncoder detected degenerate ground truth bounding boxes for "
                                         "batch item {} with bounding boxes {}, ".format(i, lab))
            if self.normalize_coords:                                         # :339-341
                lab[:, [2, 4]] /= self.img_height
                lab[:, [1, 3]] /= self.img_width
            if self.coords == 'centroids':                                    # :344-347
                lab = convert_coordinates(lab, 1, 'corners2centroids', border_pixels=self.border_pixels)
            elif self.coords == 'minmax':
                lab = convert_coordinates(lab, 1, 'corners2minmax')
            onehot = np.concatenate([eye[lab[:, 0].astype(int)], lab[:, 1:5]], axis=-1)      # :349-350
            sim = iou(lab[:, 1:5], y[i, :, -12:-8], coords=self.coords, mode='outer_product',
                      border_pixels=self.border_pixels)                      # :354
            bip = match_bipartite_greedy(sim)                                 # :360
            y[i, bip, :-8] = onehot                                           # :363 (duplicates: last wins)
            match_idx[i, bip] = np.arange(len(bip))
            sim[:, bip] = 0                                                   # :366
            if self.matching_type == 'multi':                                 # :371-381
                g_idx, a_idx = match_multi(sim, self.pos_iou_threshold)
                y[i, a_idx, :-8] = onehot[g_idx]
                match_idx[i, a_idx] = g_idx
                sim[:, a_idx] = 0
            neut = np.nonzero(np.amax(sim, axis=0) >= self.neg_iou_limit)[0]  # :388-390
            y[i, neut, self.background_id] = 0
            neutral[i, neut] = True
        if self.coords == 'centroids':                                        # :396-400
            y[:, :, [-12, -11]] -= y[:, :, [-8, -7]]
            y[:, :, [-12, -11]] /= y[:, :, [-6, -5]] * y[:, :, [-4, -3]]
            y[:, :, [-10, -9]] /= y[:, :, [-6, -5]]
            y[:, :, [-10, -9]] = np.log(y[:, :, [-10, -9]]) / y[:, :, [-2, -1]]
        elif self.coords == 'corners':                                        # :401-405
            y[:, :, -12:-8] -= y[:, :, -8:-4]
            y[:, :, [-12, -10]] /= (y[:, :, -6] - y[:, :, -8])[..., None]
            y[:, :, [-11, -9]] /= (y[:, :, -5] - y[:, :, -7])[..., None]
            y[:, :, -12:-8] /= y[:, :, -4:]
        elif self.coords == 'minmax':                                         # :406-410
            y[:, :, -12:-8] -= y[:, :, -8:-4]
            y[:, :, [-12, -11]] /= (y[:, :, -7] - y[:, :, -8])[..., None]
            y[:, :, [-10, -9]] /= (y[:, :, -5] - y[:, :, -6])[..., None]
            y[:, :, -12:-8] /= y[:, :, -4:]
        out = [y]
        if diagnostics:                                                       # :412-416
            y2 = np.copy(y)
            y2[:, :, -12:-8] = 0
            out.append(y2)
        if return_matches:
            out += [match_idx, neutral]
        return out[0] if len(out) == 1 else tuple(out)
