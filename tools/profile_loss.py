#!/usr/bin/env python
"""Profiled region: SSDLoss forward (and forward+backward) at SSD300 B=32 -- for ncu captures and a CUDA-event timing."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, ROOT)
import __graft_entry__; __graft_entry__.build()
import bench
from oracle import synth
from oracle.model import SSD300_AR
from ssd_keras_b200 import _ffi
from ssd_keras_b200.keras_loss_function.keras_ssd_loss import SSDLoss
from ssd_keras_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
ps = [(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)]
enc = SSDInputEncoder(300, 300, 20, ps, scales=bench.SC300, aspect_ratios_per_layer=SSD300_AR, steps=[8, 16, 32, 64, 100, 300],
                      offsets=[0.5] * 6, pos_iou_threshold=0.5, neg_iou_limit=0.5)
gt = synth.synth_gt(2, 32, 8, 300, 300, 20)
offs = np.cumsum([0] + [g.shape[0] for g in gt]).astype(np.int32)
y_true = enc.encode_device(torch.from_numpy(np.concatenate(gt)).cuda(), offs)
y_pred = torch.from_numpy(synth.synth_y_pred(3, 32, enc.anchors, 21, sharp=2.0)).cuda()
L = SSDLoss()
loss = torch.empty((32,), dtype=torch.float32, device='cuda'); grad = torch.empty_like(y_pred)
ctx = _ffi.context()
def fwd(): _ffi.check(_ffi.lib().ssdk_ssd_loss_fwd(ctx, _ffi.dptr(y_true), _ffi.dptr(y_pred), 32, 8732, 21, 3, 0, 1.0, _ffi.dptr(loss), _ffi.dptr(None), _ffi.stream_ptr()))
def both(): _ffi.check(_ffi.lib().ssdk_ssd_loss_fwd_bwd(ctx, _ffi.dptr(y_true), _ffi.dptr(y_pred), 32, 8732, 21, 3, 0, 1.0, _ffi.dptr(None), _ffi.dptr(loss), _ffi.dptr(None), _ffi.dptr(grad), _ffi.stream_ptr()))
print('loss fwd     : %.1f us' % (bench._time_cuda(fwd, iters=10, warm=5, inner=20) * 1e3))
print('loss fwd+bwd : %.1f us' % (bench._time_cuda(both, iters=10, warm=5, inner=20) * 1e3))
torch.cuda.profiler.start()
fwd(); both()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
