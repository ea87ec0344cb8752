"""Batch assembly for the encoder on the device.

In the reference, ``DataGenerator.generate`` (``data_generator/object_detection_2d_data_generator.py:830``) runs the augmentation
chain per image on the host -- image ops and, for every geometric op, the matching arithmetic on the label array -- filters
degenerate boxes (``:1095-1112``) and hands the list of label arrays to ``SSDInputEncoder.__call__`` (``:1146-1151``).
Here the label half runs on the GPU: ``assemble_batch_device`` uploads the ragged labels once (one pinned copy), applies a
per-image list of box operations (``ssdk_assemble_batch``: the label arithmetic of the reference's ``CropPad`` / ``Flip`` /
``Resize`` / ``BoxFilter``, with the parameters the caller's random augmentation logic picked) and leaves the packed
``(sum G_i, 5)`` rows + ``(B+1,)`` offsets on the device, which ``SSDInputEncoder.encode_device_offsets`` consumes without a
host round trip."""
import ctypes as C

import numpy as np

from .. import _ffi


def crop_pad(patch_ymin, patch_xmin, patch_height, patch_width, center_point_filter=False, clip_boxes=True):
    """Label arithmetic of ``CropPad`` (patch_sampling_ops.py:312-330).  ``SSDExpand``: negative origin, no filter, no clip;
    ``SSDRandomCrop``: ``center_point_filter=True, clip_boxes=True``."""
    return (_ffi.BOXOP_CROP_PAD, (1 if center_point_filter else 0) | (2 if clip_boxes else 0), float(patch_ymin), float(patch_xmin),
            float(patch_height), float(patch_width))


def flip(img_size, dim='horizontal'):
    """``Flip`` (geometric_ops.py:186,194): ``img_size`` is the image width (horizontal) or height (vertical)."""
    if dim not in ('horizontal', 'vertical'):
        raise ValueError("`dim` can be one of 'horizontal' and 'vertical'.")
    return (_ffi.BOXOP_FLIP_H if dim == 'horizontal' else _ffi.BOXOP_FLIP_V, 0, float(img_size), 0.0, 0.0, 0.0)


def resize(in_height, in_width, out_height, out_width, drop_degenerate=True):
    """``Resize`` (geometric_ops.py:88-100) with its degenerate-box ``BoxFilter``."""
    return (_ffi.BOXOP_RESIZE, 1 if drop_degenerate else 0, float(in_height), float(in_width), float(out_height), float(out_width))


def box_filter(check_degenerate=True, min_area=None):
    """``BoxFilter`` without the overlap test (validation_utils.py:155-165); also ``degenerate_box_handling='remove'``."""
    return (_ffi.BOXOP_FILTER, (1 if check_degenerate else 0) | (2 if min_area is not None else 0), float(min_area or 0.0), 0.0, 0.0, 0.0)


def assemble_batch_device(labels_list, ops_per_image=None):
    """``labels_list``: B arrays ``(k_i, 5)`` ``[class_id, xmin, ymin, xmax, ymax]`` (any numeric dtype, possibly empty).
    ``ops_per_image``: optional list of B lists of operations built with ``crop_pad`` / ``flip`` / ``resize`` / ``box_filter``.
    Returns ``(gt_dev (N,5) float32, offsets_dev (B+1,) int32, stats_dev int32[2] = (boxes left, largest image), total_upper,
    max_upper)`` -- CUDA tensors; the two upper bounds are host integers (box counts before filtering)."""
    import torch
    rows, offs = [], [0]
    for g in labels_list:
        g = np.asarray(g.detach().cpu().numpy() if hasattr(g, 'detach') else g, dtype=np.float64).reshape(-1, 5) if np.size(g) else np.zeros((0, 5), np.float64)
        rows.append(g)
        offs.append(offs[-1] + g.shape[0])
    B = len(rows)
    total = offs[-1]
    max_g = max([r.shape[0] for r in rows] + [0])
    flat = np.concatenate(rows, axis=0) if total else np.zeros((1, 5), np.float64)      # float64 like the reference's label arrays
    gt_in = torch.from_numpy(np.ascontiguousarray(flat)).pin_memory().cuda(non_blocking=True)
    offs_in = torch.from_numpy(np.asarray(offs, dtype=np.int32)).pin_memory().cuda(non_blocking=True)
    max_ops = max([len(o) for o in ops_per_image] + [0]) if ops_per_image else 0
    ops_dev = None
    if max_ops:
        if len(ops_per_image) != B:
            raise ValueError('ops_per_image must have one list per batch item')
        arr = (_ffi.BoxOp * (B * max_ops))()
        for b, lst in enumerate(ops_per_image):
            for i, o in enumerate(lst):
                e = arr[b * max_ops + i]
                e.op, e.flags, e.a0, e.a1, e.a2, e.a3 = o
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        ops_dev = host.pin_memory().cuda(non_blocking=True)
    gt_out = torch.empty((max(total, 1), 5), dtype=torch.float32, device='cuda')
    offs_out = torch.empty((B + 1,), dtype=torch.int32, device='cuda')
    stats = torch.empty((2,), dtype=torch.int32, device='cuda')
    _ffi.check(_ffi.lib().ssdk_assemble_batch(_ffi.context(), _ffi.dptr(gt_in), 1, _ffi.dptr(offs_in), B, int(total), _ffi.dptr(ops_dev),
                                              int(max_ops), _ffi.dptr(gt_out), _ffi.dptr(offs_out), _ffi.dptr(stats), _ffi.stream_ptr()))
    return gt_out, offs_out, stats, int(total), int(max_g)


def encode_batch_device(encoder, labels_list, ops_per_image=None, out=None):
    """``label_encoder(batch_y)`` of the reference's generator (:1146-1151) without leaving the device: box operations, packing
    and ``SSDInputEncoder`` -> float32 CUDA tensor ``(B, P, C+12)``."""
    gt, offs, _, total, max_g = assemble_batch_device(labels_list, ops_per_image)
    return encoder.encode_device_offsets(gt, offs, total, max_g, out=out)
