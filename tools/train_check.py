#!/usr/bin/env python
"""Debug harness for the backward pass / SGD step: small graphs, every gradient compared with torch autograd (float64)
through oracle/graph.py.  Each case runs in its own process.
Usage (GPU box):  python tools/train_check.py [--case N]
"""
import argparse
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)

# name, HxW, batch, layers: ('conv', name, cout, k, stride, dil, pad, act) | ('pool', name, k, stride, pad) | ('l2norm', name)
# | ('head', name, src)
CASES = [
    ('one_conv_head', 12, 2, [('conv', 'c1', 64, 3, 1, 1, 1, 1), ('conv', 'c2', 64, 3, 1, 1, 1, 1), ('head', 'h1', 'c2')]),
    ('pool_128_1x1', 16, 2, [('conv', 'c1', 64, 3, 1, 1, 1, 1), ('conv', 'c2', 64, 3, 1, 1, 1, 1), ('pool', 'p1', 2, 2, 0),
                             ('conv', 'c3', 128, 3, 1, 1, 1, 1), ('conv', 'c4', 64, 1, 1, 1, 0, 1), ('head', 'h1', 'c4')]),
    ('two_heads_l2norm', 19, 2, [('conv', 'c1', 64, 3, 1, 1, 1, 1), ('conv', 'c2', 128, 3, 1, 1, 1, 1), ('l2norm', 'n2', 'c2'),
                                 ('head', 'h1', 'n2'), ('pool', 'p1', 2, 2, 1), ('conv', 'c3', 256, 3, 1, 1, 1, 1), ('head', 'h2', 'c3')]),
    ('stride2_dil_valid', 21, 2, [('conv', 'c1', 64, 3, 1, 1, 1, 1), ('pool', 'p5', 3, 1, 1), ('conv', 'c2', 128, 3, 1, 3, 3, 1),
                                  ('conv', 'c3', 64, 1, 1, 1, 0, 1), ('conv', 'c4', 128, 3, 2, 1, 1, 1), ('head', 'h1', 'c4'),
                                  ('conv', 'c5', 64, 3, 1, 1, 0, 1), ('head', 'h2', 'c5')]),
]


def build(case):
    from ssd_keras_b200 import _ffi
    from ssd_keras_b200.models._graph import SSDModel, Spec, same_pad
    name, hw, B, layers = case
    specs = [Spec('input', _ffi.OP_INPUT, params={'mean': [127.5] * 3, 'stddev': [64.0] * 3, 'swap': [2, 1, 0]})]
    prev = 'input'
    n_heads = 0
    for l in layers:
        if l[0] == 'conv':
            _, n, cout, k, s, d, p, act = l
            specs.append(Spec(n, _ffi.OP_CONV, prev, cout=cout, k=(k, k), stride=s, dilation=d, pad=(p, p, p, p), act=act)); prev = n
        elif l[0] == 'pool':
            _, n, k, s, p = l
            pad = (p, p, p, p) if k == 3 else (0, 0, p, p)          # 2x2/s2 'same' pads at the end only
            specs.append(Spec(n, _ffi.OP_MAXPOOL, prev, k=(k, k), stride=s, pad=pad)); prev = n
        elif l[0] == 'l2norm':
            specs.append(Spec(l[1], _ffi.OP_L2NORM, l[2]))
        else:
            specs.append(Spec(l[1], _ffi.OP_HEAD, l[2], k=(3, 3), pad=same_pad(3), n_boxes=3, params={'conf_name': l[1] + 'c', 'loc_name': l[1] + 'l'}))
            n_heads += 1
    n_cls = 5
    anchor_cfg = dict(scales=list(np.linspace(0.2, 0.9, n_heads + 1)), aspect_ratios_per_layer=[[1.0, 2.0]] * n_heads,
                      two_boxes_for_ar1=True, steps=None, offsets=None, clip_boxes=False, coords='centroids', normalize_coords=True)
    m = SSDModel(specs, hw, hw, 3, n_cls, anchor_cfg, [0.1, 0.1, 0.2, 0.2], 'training', {}, precision='bf16x3', seed=3)
    rng = np.random.default_rng(7)
    w = m.get_weights()
    for k_ in w:
        if k_.endswith('/bias'):
            w[k_] = (rng.standard_normal(w[k_].shape) * 0.1).astype(np.float32)
        if k_.endswith('/gamma'):
            w[k_] = rng.uniform(5, 15, w[k_].shape).astype(np.float32)
    m.set_weights(w)
    return m, w, n_cls


def small_gt(seed, B, G, hw, ncls):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(B):
        x0 = rng.uniform(0, 0.5 * hw, G); y0 = rng.uniform(0, 0.5 * hw, G)
        w = rng.uniform(0.25 * hw, 0.5 * hw, G); h = rng.uniform(0.25 * hw, 0.5 * hw, G)
        c = rng.integers(1, ncls + 1, G)
        out.append(np.stack([c, x0, y0, np.minimum(x0 + w, hw - 1), np.minimum(y0 + h, hw - 1)], axis=1).astype(np.float32))
    return out


def run_case(i):
    import torch
    from oracle import graph as og
    from oracle.encoder import OracleEncoder
    from ssd_keras_b200.training import SSDTrainer
    case = CASES[i]
    name, hw, B, _ = case
    m, w, n_cls = build(case)
    rng = np.random.default_rng(int(os.environ.get('SSDK_CHECK_SEED', '11')))      # (another seed = another set of activations near 0)
    x = rng.integers(0, 256, size=(B, hw, hw, 3)).astype(np.float32)
    enc = OracleEncoder(hw, hw, n_cls - 1, m.predictor_sizes, scales=m.anchor_cfg['scales'], aspect_ratios_per_layer=m.anchor_cfg['aspect_ratios_per_layer'],
                        variances=[0.1, 0.1, 0.2, 0.2], pos_iou_threshold=0.3, neg_iou_limit=0.2)
    assert np.array_equal(enc.anchors, m.anchors)
    gt = small_gt(5, B, 3, hw, n_cls - 1)
    y_true = enc(gt).astype(np.float32)
    lr, mom, l2 = 1e-3, 0.9, 5e-4
    tr = SSDTrainer(m, B, lr=lr, momentum=mom, l2_regularization=l2)
    xd, ytd = torch.from_numpy(x).cuda(), torch.from_numpy(y_true).cuda()
    loss, y_pred = tr.forward_backward(xd, ytd)
    torch.cuda.synchronize()
    grads = tr.gradients()
    # oracle: float64 autograd
    params = og.make_params(m.specs, w, dtype=torch.float64)
    yp, _ = og.forward(m.specs, params, x, n_cls, m.anchors, [0.1, 0.1, 0.2, 0.2], dtype=torch.float64)
    lvec = og.ssd_loss_torch(y_true, yp)
    lvec.mean().backward()
    ok = True
    lerr = float(np.abs(loss.cpu().numpy() - lvec.detach().numpy()).max() / (np.abs(lvec.detach().numpy()).max() + 1e-30))
    print('  case %d %-18s loss %s rel err %.2e   n_pos %d' % (i, name, np.round(lvec.detach().numpy(), 4), lerr, int((y_true[:, :, 1:n_cls].max(-1) > 0).sum())), flush=True)
    ok &= lerr < 1e-4
    for k_ in sorted(grads):
        ref = params[k_].grad.numpy()
        got = grads[k_]
        scale = np.abs(ref).max() + 1e-30
        err = np.abs(got - ref).max() / scale
        flag = '' if err < 2e-3 else '   <-- BAD'
        print('  case %d %-18s grad %-12s shape %-18s max|ref| %.3e  rel err %.2e%s' % (i, name, k_, str(got.shape), scale, err, flag), flush=True)
        ok &= err < 2e-3
    # one SGD step, then a second forward/backward on the updated weights
    tr.apply(1.0)
    torch.cuda.synchronize()
    new_w = tr.get_weights()
    ref_w, ref_v = og.sgd_step(w, {k_: params[k_].grad.numpy() for k_ in w}, {}, lr, mom, l2)
    werr = max(np.abs(new_w[k_] - ref_w[k_]).max() / (np.abs(ref_w[k_]).max() + 1e-30) for k_ in w)
    print('  case %d %-18s weights after SGD step: max rel err %.2e' % (i, name, werr), flush=True)
    # the bar the gradient bar implies: an error of 2e-3 max|g| moves a weight by lr * 2e-3 max|g| ~ 2e-5 of max|w| in these graphs.
    # Measured 1e-6 ... 2e-6 when no ReLU mask differs from the float64 run and 1.06e-5 in case 3 with the fused weight operand:
    # ONE element of c2's mask differs there (c2/bias off by 4.6e-4 = that element's dY, c2/kernel by 4.6e-4 x its input 5.9)
    ok &= werr < 2e-5
    loss2, _ = tr.forward_backward(xd, ytd)
    params2 = og.make_params(m.specs, ref_w, dtype=torch.float64)
    yp2, _ = og.forward(m.specs, params2, x, n_cls, m.anchors, [0.1, 0.1, 0.2, 0.2], dtype=torch.float64)
    l2vec = og.ssd_loss_torch(y_true, yp2).detach().numpy()
    l2err = float(np.abs(loss2.cpu().numpy() - l2vec).max() / (np.abs(l2vec).max() + 1e-30))
    print('  case %d %-18s loss after the step (re-packed weights): %s rel err %.2e' % (i, name, np.round(l2vec, 4), l2err), flush=True)
    ok &= l2err < 1e-4
    print('CASE %d %s: %s' % (i, name, 'OK' if ok else 'FAIL'), flush=True)
    return 0 if ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', type=int, default=-1)
    args = ap.parse_args()
    if args.case >= 0:
        sys.exit(run_case(args.case))
    import __graft_entry__
    __graft_entry__.build()
    bad = 0
    for i in range(len(CASES)):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--case', str(i)], capture_output=True, text=True, timeout=300)
            print((r.stdout + r.stderr)[-4000:], flush=True)
            bad += (r.returncode != 0)
        except subprocess.TimeoutExpired:
            print('CASE %d TIMEOUT' % i, flush=True)
            bad += 1
    print('train_check: %d/%d cases failed' % (bad, len(CASES)))
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
