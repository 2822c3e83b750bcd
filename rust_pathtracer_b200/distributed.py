"""Sample-split multi-GPU driver: one process per GPU, each rank traces a disjoint
range of global sample indices for the WHOLE image, and the per-rank float4 (sum r, g, b, count)
partial sums are brought together on rank 0.

The counter RNG is keyed on (pixel, global sample index), so the image is independent of the
number of ranks up to f32 summation order.  Two gathers (DESIGN.md "Multi-GPU"):

* "peer" (default where CUDA IPC + peer access work): the transfer is FUSED INTO THE RENDER KERNEL — each
  rank's kernel stores every pixel's partial sum of the step straight into a slot buffer in rank 0's
  memory over NVLink (`ptb_peer_*`); per step only a stream-ordered barrier and one summing kernel on
  rank 0 remain on the critical path.
* "nccl" (the north-star baseline, and the fallback): ONE NCCL sum-reduce of the accumulators per step,
  in place onto rank 0; the other ranks hand their partial sums off and start the next step from zero.

torch.distributed is plumbing only; with backend "gloo" the same code reduces CPU tensors
(tests/test_distributed.py).
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
    """A Tracer per rank.  Two ways to bring the per-rank partial sums together on rank 0:

    * gather="peer" (default when it can be set up): every rank's render kernel stores each pixel's partial sum of the step
      straight into a slot buffer in rank 0's GPU memory over NVLink (CUDA IPC + peer stores, `ptb_peer_*`); per step only
      a stream-ordered barrier and one summing kernel on rank 0 remain.  The slots are summed in rank order: deterministic.
    * gather="nccl": the baseline — accumulators live in a torch CUDA tensor per rank and ONE NCCL sum-reduce per step
      combines them on rank 0 (also the fallback when IPC / peer access is unavailable).
    """

    def __init__(self, scene, width: int, height: int, device=None, gather: str = "auto", **tracer_kw):
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
        self.gather = "nccl"
        self.parity = 0
        if gather not in ("auto", "peer", "nccl"):
            raise ValueError("gather must be auto, peer or nccl")
        if self.world > 1 and gather in ("auto", "peer") and self.tracer.precision == "f32":
            ok = self._setup_peer()
            if not ok and gather == "peer":
                raise RuntimeError("peer gather unavailable (CUDA IPC / peer access)")
            self.gather = "peer" if ok else "nccl"

    def _setup_peer(self) -> bool:
        """Root creates the slot buffers and broadcasts the IPC handle; everybody maps them; all ranks must succeed."""
        import torch
        import torch.distributed as dist
        handle = [None]
        ok = 1
        if self.rank == 0:
            try:
                handle[0] = self.tracer.peer_slots_create(self.world)
            except Exception:
                ok = 0
        dist.broadcast_object_list(handle, src=0, device=self.device)
        if self.rank != 0:
            try:
                if handle[0] is None:
                    raise RuntimeError("root could not create the slots")
                self.tracer.peer_slots_open(handle[0], self.world)
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            try:
                self.tracer.peer_slots_close()
            except Exception:
                pass
            return False
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)
        return True

    def render(self, total_spp: int) -> None:
        """Trace this rank's share of the next `total_spp` global samples (asynchronous)."""
        base, count = split_samples(total_spp, self.world, self.rank)
        if self.gather == "peer":
            self.tracer.peer_set_target(self.rank, self.parity)
            self.tracer.render_samples(count, self.samples_done + base)     # count == 0 clears the slot
        elif count:
            self.tracer.render_samples(count, self.samples_done + base)
        self.samples_done += total_spp

    def reduce(self, dst: int = 0):
        """Bring the step's partial sums together on `dst`; returns the cumulative (sum r, g, b, count) there, None elsewhere."""
        import torch.distributed as dist
        if self.gather == "peer":
            if dst != 0:
                raise ValueError("the peer gather lands on rank 0")
            # stream-ordered barrier: completes on a rank's stream only after every rank's render kernel — and with it its
            # stores into rank 0's slots — has completed
            dist.all_reduce(self._token)
            if self.rank == 0:
                self.tracer.peer_sum(self.parity)
            self.parity ^= 1
            return self.accum if self.rank == 0 else None
        # in place: `dst` keeps the running total (its own cumulative sum + everybody's partial sums of this step); the
        # other ranks have handed their partial sums off and start the next step from zero — no copy of the 132.7 MB image
        reduce_accumulators(self.accum, dst)
        if self.rank != dst:
            self.accum.zero_()
            return None
        return self.accum

    def step(self, total_spp: int, host_buffer=None):
        """One progressive step through the host API: trace this rank's share, gather on rank 0 and — there — start the
        asynchronous download of the running-mean image into `host_buffer.pixels` (ptb_download_async: the copy overlaps the
        next step's tracing; `tracer.wait_download()` before reading).  All device work goes through the C ABI."""
        self.render(total_spp)
        self.reduce(0)
        if self.rank == 0 and host_buffer is not None:
            host_buffer.frames = self.samples_done
            host_buffer._tracer = self.tracer
            self.tracer.download_async(host_buffer)

    def close(self):
        if self.gather == "peer":
            import torch.distributed as dist
            dist.barrier()          # nobody unmaps / frees while a peer may still write
            if self.rank != 0:
                self.tracer.peer_slots_close()
            dist.barrier()
        self.tracer.close()
