"""Deterministic synthetic inputs shared by the oracle, the tests and bench.py.

Everything is generated on the host from fixed seeds so that the oracle and the CUDA path see
byte-identical tensors.  Not part of the reference; test/bench infrastructure.
"""
import numpy as np


def synth_gt(seed, B, G, W, H, ncls):
    """per image xmin=U(0,.8W), ymin=U(0,.8H), w=U(10,W/2), h=U(10,H/2), clipped to the
    image, class in [1,ncls]; rows [cls,xmin,ymin,xmax,ymax] float32, never degenerate."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(B):
        xmin = rng.uniform(0, 0.8 * W, G)
        ymin = rng.uniform(0, 0.8 * H, G)
        w = rng.uniform(10, W / 2, G)
        h = rng.uniform(10, H / 2, G)
        xmax = np.minimum(xmin + w, W - 1)
        ymax = np.minimum(ymin + h, H - 1)
        c = rng.integers(1, ncls + 1, G)
        out.append(np.stack([c, xmin, ymin, xmax, ymax], axis=1).astype(np.float32))
    return out


def synth_images(seed, B, H, W):
    """uint8-valued U[0,255] images as float32 (B,H,W,3)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(B, H, W, 3), dtype=np.uint8).astype(np.float32)


def synth_weights(seed, shapes, bias_scale=0.0):
    """he_normal kernels (N(0, sqrt(2/fan_in)), HWIO) and zero (or small) biases for ``shapes``
    = {layer: (kh,kw,cin,cout)}.  Returns {'layer/kernel': f32, 'layer/bias': f32}."""
    rng = np.random.default_rng(seed)
    w = {}
    for name in sorted(shapes):
        kh, kw, cin, cout = shapes[name]
        std = np.sqrt(2.0 / (kh * kw * cin))
        w[name + '/kernel'] = (rng.standard_normal((kh, kw, cin, cout)) * std).astype(np.float32)
        w[name + '/bias'] = (rng.standard_normal(cout) * bias_scale).astype(np.float32)
    return w


def synth_y_pred(seed, B, anchors, n_classes_total, variances=(0.1, 0.1, 0.2, 0.2), sharp=4.0, loc_scale=1.0):
    """A plausible prediction tensor (B,P,C+12) float32: softmax of random logits (``sharp`` controls
    how peaked), random offsets, the given anchors (P,4) and variances."""
    rng = np.random.default_rng(seed)
    P = anchors.shape[0]
    logits = rng.standard_normal((B, P, n_classes_total)) * sharp
    logits -= logits.max(axis=-1, keepdims=True)
    p = np.exp(logits)
    p /= p.sum(axis=-1, keepdims=True)
    loc = rng.standard_normal((B, P, 4)) * loc_scale
    y = np.concatenate([p, loc, np.broadcast_to(anchors[None], (B, P, 4)),
                        np.broadcast_to(np.asarray(variances)[None, None], (B, P, 4))], axis=-1)
    return y.astype(np.float32)


def synth_y_true_pred_for_loss(seed, encoder, B, G, ncls, sharp=2.0):
    """(y_true float32 from the oracle encoder on synth_gt, y_pred from synth_y_pred)."""
    gt = synth_gt(seed, B, G, encoder.img_width, encoder.img_height, ncls)
    y_true = encoder(gt).astype(np.float32)
    y_pred = synth_y_pred(seed + 1, B, encoder.anchors, encoder.n_classes, encoder.variances, sharp=sharp, loc_scale=0.7)
    return y_true, y_pred
