#!/usr/bin/env python
"""Per-layer conv timing of one SSD300 B=32 forward (CUDA events around every conv launch)."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, ROOT)
import bench

from oracle import synth

from ssd_keras_b200.models.keras_ssd300 import ssd_300
prec = 'bf16' if 'fast' in sys.argv else 'bf16x3'
model = ssd_300((300, 300, 3), 20, mode='training', scales=bench.SC300, precision=prec)
model.set_weights(bench._weights())
x = torch.from_numpy(synth.synth_images(0, 32, 300, 300)).cuda()
for _ in range(3): model.forward_device(x)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); model.forward_device(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
model.set_timing(32, True); model.forward_device(x); torch.cuda.synchronize()
print('forward ms (median of 5): %.3f   conv kernels ms: %.3f' % (float(np.median(ts)), model.last_conv_ms(32)))
