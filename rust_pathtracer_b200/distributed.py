"""Sample-split multi-GPU driver: one process per GPU, each rank traces a disjoint
range of global sample indices for the WHOLE image into its own accumulators, and a single NCCL
sum-reduce over NVLink combines the float4 (sum r, g, b, count) buffers on rank 0.

The counter RNG is keyed on (pixel, global sample index), so the image is independent of the
number of ranks up to f32 summation order.  There is no other exchange on the path: the reduce
moves W*H*16 bytes once per batch (132.7 MB at 4K, < 1 ms on NVLink 5), against hundreds of
milliseconds of tracing, so it is left to NCCL rather than fused into the render kernel
(DESIGN.md "Multi-GPU").  torch.distributed is plumbing only; with backend "gloo" the same code
reduces CPU tensors (tests/test_distributed.py).
"""
from __future__ import annotations

from typing import Tuple


def split_samples(total_spp: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous share of `total_spp` samples for `rank`: (first sample offset, count).
    The first total_spp % world_size ranks get one extra sample."""
    if world_size < 1 or not (0 <= rank < world_size) or total_spp < 0:
        raise ValueError("bad sample split arguments")
    q, r = divmod(total_spp, world_size)
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def reduce_accumulators(accum, dst: int = 0, group=None):
    """Sum the per-rank accumulator tensors onto `dst` (in place there); no-op without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def resolve_mean(accum):
    """(sum r, g, b, count) -> running-mean RGBA image with alpha 1 where samples landed."""
    import torch
    a = accum.view(-1, 4)
    cnt = a[:, 3:4]
    out = torch.where(cnt > 0, a / cnt.clamp(min=1), torch.zeros_like(a))
    return out.reshape(accum.shape)


class DistributedTracer:
    """A Tracer per rank whose accumulators live in a torch CUDA tensor so that NCCL can reduce them."""

    def __init__(self, scene, width: int, height: int, device=None, **tracer_kw):
        import torch
        import torch.distributed as dist
        from .prelude import Tracer
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.width, self.height = width, height
        self.tracer = Tracer.new(scene, device=self.device.index or 0, **tracer_kw)
        self.accum = torch.zeros(width * height * 4, dtype=torch.float32, device=self.device)
        self.tracer.bind_accumulator(self.accum.data_ptr(), width, height)
        self.tracer.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.samples_done = 0

    def render(self, total_spp: int) -> None:
        """Trace this rank's share of the next `total_spp` global samples (asynchronous)."""
        base, count = split_samples(total_spp, self.world, self.rank)
        if count:
            self.tracer.render_samples(count, self.samples_done + base)
        self.samples_done += total_spp

    def reduce(self, dst: int = 0):
        """One NCCL sum-reduce of the accumulators; returns the reduced copy on `dst` (None elsewhere)."""
        out = self.accum.clone()
        reduce_accumulators(out, dst)
        return out if self.rank == dst else None

    def close(self):
        self.tracer.close()
