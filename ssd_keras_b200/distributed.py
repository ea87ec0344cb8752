"""Multi-GPU plumbing for the one place the path shards: by image.

One process per GPU under torchrun / torch.distributed.  The forward pass, the encoder and the decoders are independent per
image, so ranks exchange nothing on the data path; the only collective in inference is an all-gather of the fixed-size
``(B_local, top_k, 6)`` decoded boxes (plus counts) so that every rank (or rank 0) sees the whole batch.
Works with the 'nccl' backend on CUDA tensors and with 'gloo' on CPU tensors (used by the CPU tests).
"""


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced split of ``n_items`` images: the first ``n_items % world`` ranks get one extra."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(device_index=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world)."""
    import os
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if torch.cuda.is_available():
            dev = torch.device('cuda', device_index if device_index is not None else int(os.environ.get('LOCAL_RANK', '0')))
            dist.init_process_group('nccl', device_id=dev)
        else:
            dist.init_process_group('gloo')
    return rank, world


def all_gather_detections(local, group=None):
    """``local``: (B_local, top_k, 6) tensor, the same shape on every rank -> (world * B_local, top_k, 6), rank-major
    (= global image order when the batch was split with ``shard_bounds`` into equal shards)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def all_gather_ragged(local_rows, group=None):
    """All-gather for unequal shards: pads to the largest B_local, gathers, and strips the padding."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_rows
    world = dist.get_world_size(group)
    n = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=local_rows.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    out = torch.empty((world * mx,) + tuple(pad.shape[1:]), dtype=pad.dtype, device=pad.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)], dim=0)


def all_reduce_gradients_(flat_grad, group=None):
    """Data-parallel gradient exchange of the training step: ONE all-reduce (sum) over the flat fp32
    gradient buffer, in place.  Returns the factor the optimiser must scale the summed gradient with (1 / world size), so that
    the update equals the mean of the replicas' gradients -- what Keras multi-GPU replicas of the reference's loss compute."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1.0
    dist.all_reduce(flat_grad, group=group)
    return 1.0 / dist.get_world_size(group)
