#!/usr/bin/env python
"""One profiled SSD300 B=32 inference step (+ one pass of each micro-benchmark kernel) for ncu.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_step.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'step'
    import __graft_entry__
    __graft_entry__.build()
    from oracle import synth
    from ssd_keras_b200.models.keras_ssd300 import ssd_300
    prec = 'bf16' if 'fast' in sys.argv else 'bf16x3'
    if what == 'step':
        model = ssd_300((300, 300, 3), 20, mode='inference', scales=bench.SC300, precision=prec)
        model.set_weights(bench._weights())
        x = torch.from_numpy(synth.synth_images(0, 32, 300, 300)).cuda()
        for _ in range(2):
            model.predict_device(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        model.predict_device(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    else:
        peaks, _ = bench._peaks()
        torch.cuda.profiler.start()
        print(bench.micro_benchmarks(peaks))
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


if __name__ == '__main__':
    main()
