"""Multi-GPU plumbing for the MSM path: one process per GPU (torch.distributed), points sharded
evenly by index range, one Jacobian partial per rank, a tiny all-gather (144 B / 288 B per rank
over NCCL/NVLink) and a final addition on rank 0.  No bulk data ever crosses NVLink.
The reference has no multi-device path at all (it takes devices[0], src/gpu.rs:233-234).
"""
from __future__ import annotations


def shard_range(n, rank, world):
    """contiguous even split: rank g owns [g·n/G, (g+1)·n/G)"""
    return n * rank // world, n * (rank + 1) // world


def gather_partials(partial, world, dist=None, out=None):
    """all-gather one partial (1-D int64 tensor of 18/36 limbs) from every rank → (world, limbs).
    Works with any backend (nccl on GPUs; gloo in the CPU tests).  `out`: a preallocated
    (world, limbs) tensor (bench.py's timed loop allocates nothing)."""
    import torch

    if world == 1:
        return partial.reshape(1, -1)
    if dist is None:
        import torch.distributed as dist
    if out is None:
        out = torch.empty((world, partial.numel()), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out.view(-1), partial.contiguous())
    return out


def combine_on_device(group, gathered, stream=0, out=None):
    """Σ of the gathered partials by the on-device final-addition kernel → 1-D tensor."""
    import torch

    from . import msm

    if out is None:
        out = torch.zeros(gathered.shape[1], dtype=gathered.dtype, device=gathered.device)
    msm.sum_partials_device(group, gathered.data_ptr(), gathered.shape[0], out.data_ptr(), stream)
    return out
