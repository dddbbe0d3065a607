"""Image-level sharding of predict()/score() over one-process-per-GPU ranks.

The path shards by image: no collective on the data path, one
all-reduce of the int64 confusion matrix at the end of score() and one all-gather of the label
maps at the end of predict().  The helpers work on any torch.distributed backend (NCCL on the
GPU box, gloo in the CPU tests).
"""
import numpy as np
import torch


def dist_or_none():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def rank_world():
    dist = dist_or_none()
    if dist is None:
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


def rows_for_rank(first_index, count, rank, world):
    """Rows of a `count`-image batch whose first image has global index `first_index` that
    belong to `rank`: image i goes to rank i % world."""
    return [i for i in range(count) if (first_index + i) % world == rank]


def allreduce_sum_(tensor):
    """In-place SUM all-reduce (no-op for a single process)."""
    dist = dist_or_none()
    if dist is not None:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def gather_interleaved(local, like_device):
    """all_gather the per-rank results of a round-robin image shard and restore the global
    image order.  `local` may be None on ranks that received no image."""
    dist = dist_or_none()
    if dist is None:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    count = torch.tensor([0 if local is None else local.shape[0]], dtype=torch.int64,
                         device=like_device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    counts = [int(c.item()) for c in counts]
    src = int(np.argmax(counts))
    meta = torch.zeros(8, dtype=torch.int64, device=like_device)
    codes = [torch.int64, torch.float32, torch.uint8, torch.int32, torch.float64]
    if rank == src and local is not None:
        meta[0] = local.dim()
        for i, s in enumerate(local.shape[1:]):
            meta[1 + i] = s
        meta[7] = codes.index(local.dtype)
    dist.broadcast(meta, src)
    rest = tuple(int(v) for v in meta[1:int(meta[0])])
    dtype = codes[int(meta[7])]
    pad = max(counts)
    buf = torch.zeros((pad,) + rest, dtype=dtype, device=like_device)
    if local is not None and local.shape[0]:
        buf[:local.shape[0]] = local
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    total = sum(counts)
    out = torch.zeros((total,) + rest, dtype=dtype, device=like_device)
    for r in range(world):
        if counts[r]:
            out[r:total:world] = bufs[r][:counts[r]]
    return out


def samples_for_rank(num_samples, rank, world):
    """How many of `num_samples` MC-dropout samples `rank` draws when the samples (not the
    images) are split over the ranks: the first `num_samples % world` ranks take one more."""
    return num_samples // world + (1 if rank < num_samples % world else 0)


def combine_moments_(mean, var, local_samples, extra=()):
    """Merge per-rank population moments of MC-dropout samples into the moments of all samples
    ("split the T MC samples across GPUs and allreduce first/second
    moments").  mean / var: this rank's population mean and variance over its `local_samples`
    samples (tensors of equal shape, modified in place); `extra`: further tensors whose
    sample-weighted average is wanted (e.g. a mean-over-classes variance is NOT one of them - it
    is derived from `var` afterwards).  Returns the total number of samples.

        mean = sum_r t_r mean_r / T,   var = sum_r t_r (var_r + mean_r^2) / T - mean^2
    """
    dist = dist_or_none()
    if dist is None:
        return local_samples
    t = float(local_samples)
    second = (var + mean * mean) * t
    mean.mul_(t)
    total = torch.tensor([t], dtype=torch.float64, device=mean.device)
    dist.all_reduce(mean)
    dist.all_reduce(second)
    dist.all_reduce(total)
    total_t = float(total.item())
    mean.div_(total_t)
    var.copy_(second / total_t - mean * mean)
    var.clamp_(min=0)
    for tensor in extra:
        tensor.mul_(t)
        dist.all_reduce(tensor)
        tensor.div_(total_t)
    return int(round(total_t))
