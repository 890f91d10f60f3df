"""Host-side plumbing of the sample-sharded (data-parallel) calibration: which samples a rank owns and the two
exchanges of the path.  No kernels here -- backend NCCL on GPUs, gloo in the CPU tests.

The reference has no distributed code (its only multi-GPU strategy is the layer hopping of
mobilellm/utils/parallel_utils.py:136-198); sharding by calibration sample is equivalent to the reference run with
`--batch_size world_size` because the MSE loss is a batch mean (algorithm.py:459,532-533) and running min/max are
associative (ptq/generate_act_range.py:55-69)."""
import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n, rank=None, world=None):
    """Round-robin ownership of n independent units (samples / micro-batches): i % world == rank."""
    if rank is None:
        rank, world = rank_world()
    return [i for i in range(n) if i % world == rank]


def allreduce_ranges(packed):
    """packed: float32 [n, 2] running (min, max) per statistic.  One MAX all-reduce over [-min, max]; the result is
    bit-identical on every rank and to a single-process pass over all samples."""
    rank, world = rank_world()
    if world == 1:
        return packed
    packed = packed.clone()
    packed[:, 0].neg_()
    dist.all_reduce(packed, op=dist.ReduceOp.MAX)
    packed[:, 0].neg_()
    return packed


def allreduce_grads(params, world=None):
    """SUM of the learnable-scalar gradients (LET / LWC / LRL, 0.3-1.0 M fp32) in one flat buffer, then the batch mean."""
    if world is None:
        _, world = rank_world()
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
