"""SSD multibox loss (oracle, float32 NumPy).  Not checkable against a running TensorFlow (not installable offline); PINNED
against the reference's own compute_loss source executed over a NumPy stand-in for the TensorFlow primitives it calls
(tests/golden/make_tf_golden.py, tests/test_oracle_tf_shim_golden.py: seven cases at 2e-6).

Restates ``keras_loss_function/keras_ssd_loss.py``:
  * ``smooth_L1_loss`` :53-75, ``log_loss`` :77-96, ``compute_loss`` :98-211.
``tf.nn.top_k`` tie rule restated: among equal values the lower flat index wins.
The gradient treats the hard-negative mask as a constant (it is built from
integer indices through ``tf.scatter_nd``, :185-188).
"""
import numpy as np

F = np.float32


def _per_box_losses(y_true, y_pred):
    yt = np.asarray(y_true, dtype=F)
    yp = np.asarray(y_pred, dtype=F)
    cls = -np.sum(yt[:, :, :-12] * np.log(np.maximum(yp[:, :, :-12], F(1e-15))), axis=-1, dtype=F)   # :93-95
    d = yt[:, :, -12:-8] - yp[:, :, -12:-8]
    ad = np.abs(d)
    loc = np.sum(np.where(ad < 1.0, F(0.5) * d * d, ad - F(0.5)), axis=-1, dtype=F)                 # :72-75
    return yt, yp, cls.astype(F), loc.astype(F)


def hard_negative_mask(cls, neg, n_positive, neg_pos_ratio, n_neg_min):
    """:154-188 -> (mask (B,P) float32, k)."""
    neg_all = (cls * neg).astype(F)
    n_neg_losses = int(np.count_nonzero(neg_all))                               # :154
    k = min(max(int(neg_pos_ratio) * int(n_positive), int(n_neg_min)), n_neg_losses)               # :166
    mask = np.zeros(neg_all.size, dtype=F)
    if n_neg_losses > 0 and k > 0:
        flat = neg_all.reshape(-1)
        order = np.lexsort((np.arange(flat.size), -flat.astype(np.float64)))[:k]                   # top_k, ties -> lower index
        mask[order] = 1
    return mask.reshape(neg_all.shape), k


def ssd_loss(y_true, y_pred, neg_pos_ratio=3, n_neg_min=0, alpha=1.0, return_parts=False):
    """compute_loss, :98-211 -> (B,) float32."""
    yt, yp, cls, loc = _per_box_losses(y_true, y_pred)
    B = yt.shape[0]
    neg = yt[:, :, 0]                                                           # :139
    pos = np.max(yt[:, :, 1:-12], axis=-1).astype(F)                            # :140
    n_pos = F(np.sum(pos, dtype=np.float64))                                    # :143
    pos_cls = np.sum(cls * pos, axis=-1, dtype=np.float64)                      # :148
    mask, k = hard_negative_mask(cls, neg, int(n_pos), neg_pos_ratio, n_neg_min)
    neg_cls = np.sum(cls * mask, axis=-1, dtype=np.float64)                     # :190
    loc_pos = np.sum(loc * pos, axis=-1, dtype=np.float64)                      # :202
    total = (pos_cls + neg_cls + float(alpha) * loc_pos) / max(1.0, float(n_pos))   # :204
    total = (total * B).astype(F)                                               # :209
    if return_parts:
        return total, dict(n_positive=int(n_pos), k=k, mask=mask, cls=cls, loc=loc, pos=pos)
    return total


def ssd_loss_grad(y_true, y_pred, neg_pos_ratio=3, n_neg_min=0, alpha=1.0, upstream=None):
    """d(sum_b upstream[b] * loss[b]) / d y_pred, mask held constant.  ``upstream`` defaults to 1/B
    (Keras averages the (B,) loss vector over the batch)."""
    yt, yp, cls, loc = _per_box_losses(y_true, y_pred)
    B = yt.shape[0]
    if upstream is None:
        upstream = np.full((B,), 1.0 / B)
    upstream = np.asarray(upstream, dtype=np.float64)
    neg = yt[:, :, 0]
    pos = np.max(yt[:, :, 1:-12], axis=-1).astype(F)
    n_pos = float(np.sum(pos, dtype=np.float64))
    mask, _ = hard_negative_mask(cls, neg, int(n_pos), neg_pos_ratio, n_neg_min)
    scale = (upstream * B / max(1.0, n_pos))[:, None]                           # (B,1)
    w_cls = (pos + mask).astype(np.float64) * scale                             # weight of cls loss per box
    w_loc = pos.astype(np.float64) * scale * float(alpha)
    g = np.zeros(yp.shape, dtype=np.float64)
    p = yp[:, :, :-12].astype(np.float64)
    dlog = np.where(p >= 1e-15, 1.0 / np.maximum(p, 1e-15), 0.0)               # d log(max(p,eps)) / dp
    g[:, :, :-12] = -yt[:, :, :-12].astype(np.float64) * dlog * w_cls[..., None]
    d = yp[:, :, -12:-8].astype(np.float64) - yt[:, :, -12:-8].astype(np.float64)   # = -(y_true - y_pred)
    g[:, :, -12:-8] = np.where(np.abs(d) < 1.0, d, np.sign(d)) * w_loc[..., None]
    return g.astype(F)
