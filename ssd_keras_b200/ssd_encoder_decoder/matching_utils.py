"""``match_bipartite_greedy`` / ``match_multi`` with the reference signatures (``ssd_encoder_decoder/matching_utils.py:22-116``)
for callers that hold a similarity matrix of their own.

The encoder does NOT go through these: ``ssdk_encode`` fuses IoU, both matchings and the target encoding into its own kernels
(``csrc/encode.cu``) and never materialises the (G, P) matrix.  These stand-alone versions exist for API completeness; they
run as a handful of tensor operations on the GPU (the matrix is uploaded, the G greedy rounds are G small arg-max launches)
and reproduce NumPy's first-index tie rule, which ``torch.argmax`` documents."""
import numpy as np


def _bipartite_t(w):
    """w: (G, P) float64 tensor on any device -> (G,) int64 tensor (matching_utils.py:63-77)."""
    import torch
    w = w.clone()
    G = w.shape[0]
    matches = torch.zeros((G,), dtype=torch.int64, device=w.device)
    rows = torch.arange(G, device=w.device)
    for _ in range(G):
        anchor_indices = torch.argmax(w, dim=1)                 # :67  first maximal index per row
        overlaps = w[rows, anchor_indices]                      # :68
        g = torch.argmax(overlaps)                              # :69  first maximal row
        a = anchor_indices[g]
        matches[g] = a                                          # :72
        w[g, :] = 0                                             # :76
        w[:, a] = 0                                             # :77
    return matches


def _multi_t(w, threshold):
    """-> (gt indices, anchor indices) int64 tensors, anchors ascending (matching_utils.py:107-116)."""
    import torch
    gt = torch.argmax(w, dim=0)                                 # :109
    overlaps = w[gt, torch.arange(w.shape[1], device=w.device)]
    met = torch.nonzero(overlaps >= threshold).reshape(-1)      # :113
    return gt[met], met


def _to_gpu(weight_matrix):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('ssd_keras_b200.matching_utils needs a CUDA device (there is no CPU fallback)')
    return torch.from_numpy(np.ascontiguousarray(np.asarray(weight_matrix, dtype=np.float64))).cuda()


def match_bipartite_greedy(weight_matrix):
    return _bipartite_t(_to_gpu(weight_matrix)).cpu().numpy()


def match_multi(weight_matrix, threshold):
    g, a = _multi_t(_to_gpu(weight_matrix), float(threshold))
    return g.cpu().numpy(), a.cpu().numpy()
