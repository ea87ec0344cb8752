"""world_size-2 gloo tests of the N>1 host logic (batch sharding, all-gather of decoded boxes, gradient all-reduce of the
training step), runnable without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ssd_keras_b200 import distributed as D


def test_shard_bounds_cover_batch():
    for n in (1, 2, 7, 16, 32, 33, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        full = torch.from_numpy(rng.standard_normal((n_images, 5, 6)).astype(np.float32))   # "decoded boxes" of the whole batch
        lo, hi = D.shard_bounds(n_images, rank, world)
        local = full[lo:hi].clone()
        if n_images % world == 0:
            got = D.all_gather_detections(local)
        else:
            got = D.all_gather_ragged(local)
        ok = bool(torch.equal(got, full))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t.item()) == float(world)
        # gradient exchange of the training step: one all-reduce of the flat buffer, update scaled by 1 / world
        g = torch.arange(10, dtype=torch.float32) * (rank + 1)
        scale = D.all_reduce_gradients_(g)
        ok = ok and scale == 1.0 / world and bool(torch.equal(g, torch.arange(10, dtype=torch.float32) * sum(range(1, world + 1))))
        # bucketed, overlapped exchange: layers are "differentiated" top down, each bucket is reduced as soon as it is complete
        sizes = [0, 5, 0, 7, 3, 0, 9, 4]                                   # parameters per graph layer (0: pool / input)
        first, o = [], 0
        for sz in sizes:
            first.append(o if sz else None); o += sz
        buckets = D.plan_buckets(first, sizes, bucket_bytes=40)            # 10 floats per bucket
        flat = torch.zeros(o)
        done = []

        def produce(hi, lo):                                               # fills the gradients of layers hi .. lo only
            for i in range(hi, lo - 1, -1):
                if sizes[i]:
                    flat[first[i]:first[i] + sizes[i]] = float(rank + 1) * (i + 1)
            done.append((hi, lo))
        n_coll = D.all_reduce_buckets_(flat, buckets, produce)
        exp = torch.cat([torch.full((sz,), float(sum(range(1, world + 1))) * (i + 1)) for i, sz in enumerate(sizes) if sz])
        ok = ok and bool(torch.equal(flat, exp)) and n_coll == len([b for b in buckets if b[3]])
        ok = ok and done[0][0] == len(sizes) - 1 and done[-1][1] == 0 and all(a[1] == b[0] + 1 for a, b in zip(done, done[1:]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_images', [8, 7])
def test_all_gather_two_ranks_gloo(n_images):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_gradient_exchange_is_identity_without_process_group():
    g = torch.ones(4)
    assert D.all_reduce_gradients_(g) == 1.0 and bool(torch.equal(g, torch.ones(4)))


def test_plan_buckets_partitions_the_layers():
    sizes = [0, 1728, 36864, 0, 73728, 147456, 0, 2359296, 4718592, 1000, 0, 12]
    first, o = [], 0
    for sz in sizes:
        first.append(o if sz else None); o += sz
    for bb in (1, 4 * 100000, 4 * 3000000, 1 << 40):
        b = D.plan_buckets(first, sizes, bb)
        assert b[0][0] == len(sizes) - 1 and b[-1][1] == 0
        assert all(x[1] == y[0] + 1 for x, y in zip(b, b[1:]))                      # consecutive, top down, no gaps
        assert sum(x[3] for x in b) == o
        for hi, lo, off, cnt in b:                                                 # the span is exactly the layers' parameters
            owned = [i for i in range(lo, hi + 1) if sizes[i]]
            if owned:
                assert off == first[owned[0]] and cnt == sum(sizes[i] for i in owned)
        if bb == 1 << 40:
            assert len(b) == 1
