"""Batch-sharded data parallelism (reference: mimic/main_mimic.py:44-48,65-67, utils/utils.py:179-185,
run_epochs.py:245-247, dataio/utils.py:120-123): one process per GPU, per-rank batch = global // world,
BatchNorm statistics and mixture selection per rank, gradients mean-reduced.

The path shards by samples and has exactly one exchange step — the gradient all-reduce — so that is the only
collective.  All 420 gradients live in ONE flat fp32 buffer: the reduction is a handful of large bucketed
`all_reduce` calls (NCCL over NVLink/NVSwitch on GPUs; the same host logic runs on gloo for the CPU tests), and the
1/world scale is folded into the Adam kernel instead of a separate pass.
"""
import torch
import torch.distributed as dist


def shard_batch(batch, rank, world):
    """contiguous shard of every tensor in a {name: tensor} batch — what DistributedSampler + per-rank
    batch_size //= world amounts to for one step (dataio/utils.py:120-123, main_mimic.py:48)"""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def bucket_bounds(numel, bucket_elems):
    """[start, end) element ranges covering `numel` in buckets of at most `bucket_elems` (256-B aligned)"""
    bucket_elems = max(64, bucket_elems // 64 * 64)
    return [(s, min(numel, s + bucket_elems)) for s in range(0, numel, bucket_elems)]


class FlatGradAllReduce:
    """Sum-all-reduce of a flat gradient buffer in large buckets.  Callable: `allreduce(flat_grads)`.

    `grad_scale` (= 1 / world) is what the optimizer must multiply gradients with to obtain the DDP mean."""

    def __init__(self, group=None, bucket_mb=256):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_elems = int(bucket_mb * (1 << 20) // 4)
        self.grad_scale = 1.0 / self.world

    def __call__(self, flat):
        if self.world == 1:
            return flat
        for s, e in bucket_bounds(flat.numel(), self.bucket_elems):
            dist.all_reduce(flat[s:e], op=dist.ReduceOp.SUM, group=self.group)
        return flat


def broadcast_flat(flat, src=0, group=None):
    """rank-0 parameters to every rank (what DDP's constructor does)"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src, group=group)
    return flat
