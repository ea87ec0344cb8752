"""GT<->anchor matching (oracle).

Restates ``ssd_encoder_decoder/matching_utils.py``:
  * ``match_bipartite_greedy`` :22-79
  * ``match_multi``            :81-116
Pinned against the real reference by tests/golden/make_golden.py.
"""
import numpy as np


def match_bipartite_greedy(weight_matrix):
    """matching_utils.py:22-79.

    G rounds; each round takes the globally best (gt, anchor) pair -- found as
    row-argmax then argmax over rows, both first-index on ties -- records it and
    zeroes that row and column.  Rows already matched stay in the competition with
    value 0 (quirk): when every remaining overlap is 0 the round
    re-selects gt 0 / anchor 0 and overwrites ``matches[0]``.
    """
    w = np.array(weight_matrix, copy=True)
    n_gt = w.shape[0]
    matches = np.zeros(n_gt, dtype=int)
    rows = np.arange(n_gt)
    for _ in range(n_gt):
        best_anchor_per_gt = w.argmax(axis=1)
        best_val_per_gt = w[rows, best_anchor_per_gt]
        g = int(best_val_per_gt.argmax())
        a = int(best_anchor_per_gt[g])
        matches[g] = a
        w[g, :] = 0
        w[:, a] = 0
    return matches


def match_multi(weight_matrix, threshold):
    """matching_utils.py:81-116: per-anchor best gt (first index), kept where >= threshold."""
    w = np.asarray(weight_matrix)
    best_gt = w.argmax(axis=0)
    best_val = w[best_gt, np.arange(w.shape[1])]
    anchors = np.nonzero(best_val >= threshold)[0]
    return best_gt[anchors], anchors
