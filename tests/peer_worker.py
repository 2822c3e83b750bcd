"""Worker of tests/test_distributed.py::test_peer_gather_matches_nccl_reduce — launched under torchrun with one rank per
GPU.  Renders the same steps with the NCCL-reduce gather and with the peer-memory gather and compares on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import rust_pathtracer_b200 as rp
from rust_pathtracer_b200.distributed import DistributedTracer, resolve_mean

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
W, H, STEPS, SPP = 320, 180, 3, 5        # 5 spp over `world` ranks: uneven shares; with world > 5 some ranks trace nothing
scene = rp.AnalyticalScene.new()
out = {}
for integ in (rp._abi.PTB_INTEGRATOR_WAVEFRONT, rp._abi.PTB_INTEGRATOR_FUSED, rp._abi.PTB_INTEGRATOR_STREAM):
    for gather in ("nccl", "peer"):
        dt = DistributedTracer(scene, W, H, device=torch.device("cuda", local), gather=gather, integrator=integ)
        assert dt.gather == gather
        res = None
        for _ in range(STEPS):
            dt.render(SPP)
            res = dt.reduce(0)
        torch.cuda.synchronize()
        if rank == 0:
            out[(integ, gather)] = resolve_mean(res).cpu().numpy().copy()
        dt.close()
if rank == 0:
    for integ in (rp._abi.PTB_INTEGRATOR_WAVEFRONT, rp._abi.PTB_INTEGRATOR_FUSED, rp._abi.PTB_INTEGRATOR_STREAM):
        # the single-GPU image of the SAME integrator: the same samples through the same code, so only the f32 summation
        # order differs (the wavefront integrator shades this scene from the host-evaluated material table, the other two
        # from device arithmetic; a few pixels per 100 000 take a different branch between those two)
        single = rp.Tracer.new(scene, device=local, integrator=integ)
        buf = rp.ColorBuffer.new(W, H)
        single.render_spp(buf, STEPS * SPP)
        single.close()
        a, b = out[(integ, "nccl")], out[(integ, "peer")]
        assert np.all(b.reshape(-1, 4)[:, 3] == 1.0)
        # (not bit-equal: NCCL reduces the ranks' CUMULATIVE sums, the peer gather adds each step's partial sums in rank order)
        assert np.allclose(a, b, rtol=5e-6, atol=1e-7), float(np.abs(a - b).max())
        assert np.allclose(b, buf.pixels, rtol=1e-4, atol=1e-6), float(np.abs(b - buf.pixels).max())
    print("PEER_OK", world, flush=True)
dist.barrier()
dist.destroy_process_group()
