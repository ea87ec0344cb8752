#!/usr/bin/env python
"""Where does an end-to-end step spend its time?  H2D bandwidth of a pinned batch, the per-step loop with a full sync, and
SSDModel.predict_stream with host timestamps per phase."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, ROOT)
import bench
from oracle import synth
from ssd_keras_b200.models.keras_ssd300 import ssd_300

model = ssd_300((300, 300, 3), 20, mode='inference', scales=bench.SC300)
model.set_weights(bench._weights())
host = [torch.from_numpy(synth.synth_images(i, 32, 300, 300)).pin_memory() for i in range(4)]
print('pinned:', [h.is_pinned() for h in host])
dev = [h.cuda() for h in host]
for i in range(3):
    model.predict_device(dev[i])
torch.cuda.synchronize()

# (a) H2D alone
up = torch.cuda.Stream()
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(up):
        a.record(up)
        xs = [host[i % 4].cuda(non_blocking=True) for i in range(8)]
        b.record(up)
    torch.cuda.synchronize()
    print('H2D alone: %.3f ms per 34.6 MB batch (%.1f GB/s)' % (a.elapsed_time(b) / 8, 8 * 34.56e-3 / (a.elapsed_time(b) * 1e-3)))
del xs

# (b) device-resident loop
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(10):
    model.predict_device(dev[i % 4])
b.record(); torch.cuda.synchronize()
print('device-resident: %.3f ms/step' % (a.elapsed_time(b) / 10))

# (c) predict_stream, with host timestamps
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    stamps = []
    for r in model.predict_stream(host[i % 4] for i in range(10)):
        stamps.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    print('predict_stream: %.3f ms/step; yields at' % ((time.perf_counter() - t0) * 100), ' '.join('%.1f' % (s * 1e3) for s in stamps))

# (d) the same loop written out, timing the host side of every phase
main = torch.cuda.current_stream()
def upload(hb):
    with torch.cuda.stream(up):
        x = hb.cuda(non_blocking=True)
        ev = torch.cuda.Event(); ev.record(up)
    return x, ev
for variant in ('sync_prev', 'sync_now'):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows = []
    nxt = upload(host[0]); pending = None
    out_pin = [torch.empty((32, 200, 6)).pin_memory() for _ in range(3)]
    for i in range(10):
        t = [time.perf_counter()]
        x, ev = nxt
        nxt = upload(host[(i + 1) % 4]); t.append(time.perf_counter())
        main.wait_event(ev); x.record_stream(main)
        out = model.predict_device(x); t.append(time.perf_counter())
        out_pin[i % 3].copy_(out, non_blocking=True)
        done = torch.cuda.Event(); done.record(main); t.append(time.perf_counter())
        if variant == 'sync_now':
            done.synchronize()
        elif pending is not None:
            pending.synchronize()
        pending = done; t.append(time.perf_counter())
        rows.append(' '.join('%.2f' % ((b - a) * 1e3) for a, b in zip(t[:-1], t[1:])))
    torch.cuda.synchronize()
    print('%s: %.3f ms/step; per step [upload | launch | d2h | wait] ms:' % (variant, (time.perf_counter() - t0) * 100))
    for r in rows:
        print('   ', r)
