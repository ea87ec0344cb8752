#!/usr/bin/env python
"""Time the SSD300 training step (BASELINE config 3: batch 32, 21 classes): forward, backward, update.
Usage (GPU box): python tools/train_bench.py [--batch 32] [--steps 5]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    args = ap.parse_args()
    import torch
    import __graft_entry__
    __graft_entry__.build()
    from oracle import synth
    from ssd_keras_b200.models.keras_ssd300 import ssd_300
    from ssd_keras_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
    from ssd_keras_b200.training import SSDTrainer
    B = args.batch
    sc = [0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05]
    m = ssd_300((300, 300, 3), 20, mode='training', scales=sc)
    enc = SSDInputEncoder(300, 300, 20, m.predictor_sizes, scales=sc, aspect_ratios_per_layer=m.anchor_cfg['aspect_ratios_per_layer'],
                          steps=[8, 16, 32, 64, 100, 300], variances=[0.1, 0.1, 0.2, 0.2])
    x = torch.from_numpy(synth.synth_images(0, B, 300, 300)).cuda()
    y_true = torch.from_numpy(enc(synth.synth_gt(1, B, 8, 300, 300, 20)).astype(np.float32)).cuda()
    tr = SSDTrainer(m, B, lr=1e-4, momentum=0.9)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t = np.zeros(3)
    for i in range(args.warmup + args.steps):
        ev[0].record()
        yp = m.forward_device(x, training=True)
        ev[1].record()
        loss = torch.empty((B,), dtype=torch.float32, device='cuda')
        from ssd_keras_b200 import _ffi
        _ffi.check(_ffi.lib().ssdk_train_backward(tr.handle, _ffi.dptr(y_true), _ffi.dptr(yp), 3, 0, 1.0, _ffi.dptr(loss), _ffi.stream_ptr()))
        ev[2].record()
        tr.apply(1.0)
        ev[3].record()
        torch.cuda.synchronize()
        if i >= args.warmup:
            t += [ev[j].elapsed_time(ev[j + 1]) for j in range(3)]
        print('step %d loss %.4f' % (i, float(loss.mean().item())), flush=True)
    t /= args.steps
    print('B=%d forward %.2f ms  backward %.2f ms  update %.2f ms  total %.2f ms  -> %.1f images/s' % (B, t[0], t[1], t[2], t.sum(), B / t.sum() * 1e3))
    print('memory allocated by torch: %.2f GB; device used %.2f GB' % (torch.cuda.memory_allocated() / 2**30, (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2**30))


if __name__ == '__main__':
    main()
