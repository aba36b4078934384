"""Data parallelism for the Wave-U-Net engine: one process per GPU, torch.distributed for the plumbing.

The reference is single-device.  Training minibatches are independent through the whole network
(no batch-norm), so a step shards the batch over the ranks, every rank runs forward/backward on its shard with
the gradient pre-scaled by 1/world (the loss is a mean over the batch, /root/reference/Training.py:62), and ONE
all-reduce (sum) of the flat fp32 gradient buffer makes every replica apply the same Adam update.  Inference
windows (/root/reference/Evaluate.py:125-139) are independent: contiguous ranges per rank, no collective.

Everything here works on CPU tensors with the gloo backend too (tests/test_parallel_cpu.py).
"""
import torch


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank, world_size):
    """Contiguous [lo, hi) of `n_items` for `rank`; sizes differ by at most one (first ranks get the extra)."""
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensor, rank, world_size, dim=0):
    lo, hi = shard_range(tensor.shape[dim], rank, world_size)
    return tensor.narrow(dim, lo, hi - lo)


def grad_scale_for(local_batch, global_batch):
    """Factor each rank applies to its local mean-loss gradient so that the SUM over ranks is the gradient of the
    global mean loss (ranks may hold different numbers of examples)."""
    return float(local_batch) / float(global_batch)


def allreduce_gradients(flat_grads, group=None):
    """The single collective of a training step: in-place sum of the flat gradient buffer over all ranks."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return flat_grads


def broadcast_parameters(flat_params, src=0, group=None):
    """Make every replica start from rank `src`'s variables (the DP invariant)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat_params, src=src, group=group)
    return flat_params


def max_over_ranks(value, device="cpu", group=None):
    """Timing reduction for benchmarks: the slowest rank defines the step time."""
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_window_predictions(local_preds, n_windows_total, group=None):
    """Inference: every rank predicted a contiguous range of windows [K, n_local, T_out, C]; returns the full
    [K, n_windows_total, T_out, C] on every rank (all_gather of padded shards)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_preds
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    K, _, T, C = local_preds.shape
    max_local = (n_windows_total + ws - 1) // ws
    pad = torch.zeros((K, max_local, T, C), dtype=local_preds.dtype, device=local_preds.device)
    pad[:, :local_preds.shape[1]] = local_preds
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r in range(ws):
        lo, hi = shard_range(n_windows_total, r, ws)
        out.append(parts[r][:, :hi - lo])
    return torch.cat(out, dim=1)
