#!/usr/bin/env python
"""Profiled region: one config-5 style encode (P=1e5, G=128, B from argv) and one SSD300 B=32 encode."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, ROOT)
import __graft_entry__; __graft_entry__.build()
from oracle import synth
from ssd_keras_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
Bm = int(sys.argv[1]) if len(sys.argv) > 1 else 64
encm = SSDInputEncoder(1000, 1600, 20, [(125, 200)], scales=[0.1, 0.2], aspect_ratios_global=[0.5, 1.0, 2.0], pos_iou_threshold=0.5, neg_iou_limit=0.5)
gtm = synth.synth_gt(4, Bm, 128, 1600, 1000, 20)
offm = np.cumsum([0] + [g.shape[0] for g in gtm]).astype(np.int32)
gm = torch.from_numpy(np.concatenate(gtm)).cuda()
for _ in range(2): encm.encode_device(gm, offm)
torch.cuda.synchronize(); torch.cuda.profiler.start()
encm.encode_device(gm, offm)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
