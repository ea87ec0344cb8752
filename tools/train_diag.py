#!/usr/bin/env python
"""Per-tensor gradient error of the SSD300 training step against float64 autograd (GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', 'tests')))
import torch
import __graft_entry__
__graft_entry__.build()
import test_gpu_train as T
from oracle import graph as og
from ssd_keras_b200.training import SSDTrainer
B = 2
m, w, x, y_true = T._ssd300(B)
tr = SSDTrainer(m, B, lr=1e-3, momentum=0.9, l2_regularization=5e-4)
xd, ytd = torch.from_numpy(x).cuda(), torch.from_numpy(y_true).cuda()
loss, yp_dev = tr.forward_backward(xd, ytd)
torch.cuda.synchronize()
grads = tr.gradients()
for dt in (torch.float64, torch.float32):
    params = og.make_params(m.specs, w, dtype=dt)
    yp, _ = og.forward(m.specs, params, x, 21, m.anchors, [0.1, 0.1, 0.2, 0.2], dtype=dt)
    lvec = og.ssd_loss_torch(y_true, yp)
    lvec.mean().backward()
    if dt == torch.float64:
        ref = {k: params[k].grad.numpy().copy() for k in w}
        yp64 = yp.detach().numpy()
        l64 = lvec.detach().numpy()
    else:
        ref32 = {k: params[k].grad.numpy().astype(np.float64) for k in w}
        l32 = lvec.detach().numpy()
print('loss ours', loss.cpu().numpy(), 'f64', l64, 'f32 torch', l32)
ypd = yp_dev.cpu().numpy()
print('y_pred max abs err conf %.3e loc %.3e (max |loc| %.3e)' % (np.abs(ypd[..., :21] - yp64[..., :21]).max(), np.abs(ypd[..., 21:25] - yp64[..., 21:25]).max(), np.abs(yp64[..., 21:25]).max()))
names = [s.name for s in m.specs]
for s in m.specs:
    for k in sorted(grads):
        if k.split('/')[0] in (s.name, s.params.get('conf_name') if s.params else None, s.params.get('loc_name') if s.params else None):
            sc = np.abs(ref[k]).max() + 1e-30
            print('%-28s max|ref| %.3e  ours-f64 %.2e  torch32-f64 %.2e  mean signed (ours-f64)/max %.2e' % (k, sc, np.abs(grads[k] - ref[k]).max() / sc, np.abs(ref32[k] - ref[k]).max() / sc, (grads[k] - ref[k]).mean() / sc))
for k in ('conv6_1/bias', 'fc7/bias', 'conv5_3/bias', 'conv1_1/bias', 'conv6_1/kernel'):
    e = np.abs(grads[k] - ref[k]).ravel() / (np.abs(ref[k]).max() + 1e-30)
    srt = np.sort(e)[::-1]
    print('%-16s top errors %s  median %.2e  90%% %.2e' % (k, np.array2string(srt[:6], precision=2), np.median(e), np.quantile(e, 0.9)))
