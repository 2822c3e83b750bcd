"""CPU tests of the sample-split multi-GPU host logic with world_size-2 gloo process groups.
The per-rank "renderer" here is the oracle (this is a test), the split / reduce / resolve code is the
product's (rust_pathtracer_b200/distributed.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_samples_partitions_exactly(rp):
    from rust_pathtracer_b200.distributed import split_samples
    for total in (0, 1, 7, 128, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            parts = [split_samples(total, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == total
            cur = 0
            for base, count in parts:
                assert base == cur
                cur += count
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        split_samples(8, 2, 2)


def _worker(rank, world, port, W, H, spp, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import rust_pathtracer_b200 as rp
    from rust_pathtracer_b200.distributed import reduce_accumulators, resolve_mean, split_samples
    from oracle import pyoracle as po
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, count = split_samples(spp, world, rank)
    sc = po.OracleScene(rp.AnalyticalScene.new().device_export())
    # per-rank accumulators (sum r,g,b,count) for samples [base, base+count), one sample at a time
    acc = np.zeros(W * H * 4, np.float64)
    for s in range(base, base + count):
        px, _, _, _ = sc.render(W, H, 1, sample_base=s, threads=2)
        acc += px.astype(np.float64)
    t = torch.from_numpy(acc)
    reduce_accumulators(t, dst=0)
    if rank == 0:
        q.put(resolve_mean(t).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_sample_split_reduce_matches_single_rank(po, rp):
    import torch.multiprocessing as mp
    W, H, spp, world = 48, 32, 5, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, spp, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, frames, _, _ = po.OracleScene(rp.AnalyticalScene.new().device_export()).render(W, H, spp)
    assert frames == spp
    assert np.allclose(got.reshape(-1, 4)[:, :3], ref.reshape(-1, 4)[:, :3], rtol=1e-5, atol=1e-7)
    assert np.all(got.reshape(-1, 4)[:, 3] == 1.0)


def test_resolve_mean_handles_empty_pixels():
    import torch
    from rust_pathtracer_b200.distributed import resolve_mean
    a = torch.tensor([2.0, 4.0, 6.0, 2.0, 0.0, 0.0, 0.0, 0.0])
    assert resolve_mean(a).tolist() == [1.0, 2.0, 3.0, 1.0, 0.0, 0.0, 0.0, 0.0]


@pytest.mark.gpu
@pytest.mark.parametrize("integ_name", ["FUSED", "WAVEFRONT", "STREAM"])
def test_peer_slot_path_equals_local_accumulation(rp, integ_name):
    """one GPU: partial sums stored into slot buffers (ptb_peer_*) and summed by k_peer_sum == the same samples accumulated
    locally, bit for bit (the additions happen in the same order)"""
    integ = getattr(rp._abi, "PTB_INTEGRATOR_" + integ_name)
    scene = rp.AnalyticalScene.new()
    W, H = 200, 120
    plain = rp.Tracer.new(scene, integrator=integ)
    b0 = rp.ColorBuffer.new(W, H)
    plain._ensure_size(b0); plain.clear()
    plain.render_samples(3, 0); plain.render_samples(2, 3)
    plain.download(b0); plain.close()
    pt = rp.Tracer.new(scene, integrator=integ)
    b1 = rp.ColorBuffer.new(W, H)
    pt._ensure_size(b1); pt.clear()
    handle = pt.peer_slots_create(3)
    assert len(handle) == rp._abi.PTB_PEER_HANDLE_BYTES
    for parity in (0, 1):
        pt.peer_set_target(0, parity); pt.render_samples(3, 0)
        pt.peer_set_target(1, parity); pt.render_samples(2, 3)
        pt.peer_set_target(2, parity); pt.render_samples(0, 5)          # a rank without samples: its slot must read as empty
    pt.peer_set_target(0xffffffff)
    pt.peer_sum(1)
    pt.download(b1)
    assert np.array_equal(b0.pixels, b1.pixels)
    with pytest.raises(Exception):
        pt.peer_set_target(3, 0)
    pt.peer_slots_close(); pt.close()


@pytest.mark.gpu
def test_peer_gather_matches_nccl_reduce(rp):
    """>= 2 GPUs: torchrun, one rank per GPU; the peer-memory gather, the NCCL reduce and a single-GPU render agree"""
    import subprocess
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 8 else 8
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "peer_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and f"PEER_OK {world}" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
