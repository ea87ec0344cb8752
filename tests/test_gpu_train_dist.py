"""2-GPU data-parallel training step: each rank runs forward/backward on its shard of the batch,
ONE NCCL all-reduce sums the flat gradient buffer, every rank applies the same update.  Checked against a single-GPU step
on the full batch.  Note the reference's loss normalises by the number of positives of the batch it sees
(keras_ssd_loss.py:143), so the data-parallel gradient is mean_r(grad_r) of replica-local losses, exactly what
multi-GPU Keras replicas compute -- the single-GPU comparison is therefore made on that quantity."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, 'tests'))
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
import test_gpu_train as T
from ssd_keras_b200.training import SSDTrainer
B = 4
m, w, x, y_true = T._ssd300(B, seed=7)
lo, hi = rank * B // world, (rank + 1) * B // world
tr = SSDTrainer(m, hi - lo, lr=1e-3, momentum=0.9, l2_regularization=5e-4)
xd, yd = torch.from_numpy(x[lo:hi]).cuda(), torch.from_numpy(y_true[lo:hi]).cuda()
loss = tr.train_on_batch(xd, yd)
torch.cuda.synchronize()
g = tr.grad.clone()                     # summed over ranks by train_on_batch
wts = tr.get_weights()
# every rank must hold identical gradients and identical updated weights
g0 = g.clone(); dist.broadcast(g0, 0)
assert torch.equal(g, g0), 'ranks disagree on the reduced gradient'
if rank == 0:
    # reference: the two shards on ONE GPU, gradients summed by hand, update with scale 1/world
    acc = None
    for r in range(world):
        a, b = r * B // world, (r + 1) * B // world
        m2, _, _, _ = T._ssd300(B, seed=7)
        t2 = SSDTrainer(m2, b - a, lr=1e-3, momentum=0.9, l2_regularization=5e-4)
        t2.forward_backward(torch.from_numpy(x[a:b]).cuda(), torch.from_numpy(y_true[a:b]).cuda())
        acc = t2.grad.clone() if acc is None else acc + t2.grad
        if r == 0:
            keep = t2
    err = float((acc - g).abs().max() / acc.abs().max())
    print('dist-train: max gradient deviation %%.3e' %% err)
    assert err < 1e-5, err                # fp32 atomics inside wgrad reorder sums from run to run
    np.savez(os.path.join(%(root)r, 'gpurun_out', 'dist_train_ok.npz'), err=err)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_train_step(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29631', str(script)], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert 'dist-train: max gradient deviation' in r.stdout
