#!/usr/bin/env python
"""Profiled region: five SSD300 / VOC batch-32 encodes (config 3) -- for an ncu launch list or --set full capture."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, ROOT)
import __graft_entry__; __graft_entry__.build()
import bench
from oracle import synth
from oracle.model import SSD300_AR
from ssd_keras_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
ps = [(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)]
enc = SSDInputEncoder(300, 300, 20, ps, scales=bench.SC300, aspect_ratios_per_layer=SSD300_AR, steps=[8, 16, 32, 64, 100, 300],
                      offsets=[0.5] * 6, pos_iou_threshold=0.5, neg_iou_limit=0.5)
gt = synth.synth_gt(2, 32, 8, 300, 300, 20)
offs = np.cumsum([0] + [g.shape[0] for g in gt]).astype(np.int32)
gdev = torch.from_numpy(np.concatenate(gt)).cuda()
for _ in range(3): enc.encode_device(gdev, offs)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); enc.encode_device(gdev, offs); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print('encode SSD300 B=32: median %.1f us, min %.1f us (CUDA events around the call)' % (np.median(ts) * 1e3, np.min(ts) * 1e3))
torch.cuda.profiler.start()
for _ in range(5): enc.encode_device(gdev, offs)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
