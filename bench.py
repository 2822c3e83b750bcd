#!/usr/bin/env python
"""bench.py — headline benchmark: path Msamples/s at 4K on 1/2/4/8 B200 (BASELINE.json).

Workload (config.workload): BASELINE.json configs[2] — the reference's AnalyticalScene
(renderer/src/analytical.rs) at 3840x2160, depth 4, f32; one STEP = SPP_PER_STEP (128) samples per
pixel over the whole frame (1.06e9 path samples), sample-split across the ranks, followed by ONE
NCCL sum-reduce of the float4 accumulators onto rank 0.  The default K = 8 steps are the config's
1024 spp.  Synthetic data: the scene is the reference's own analytic demo scene (no assets exist).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # own arm (CUDA, sm_100a)
    python bench.py --impl reference [...]                         # the reference algorithm on host cores

Keys beyond the base contract: `roofline` (FP32 FMA pipe — the path has no dense contraction and
moves 32 B per pixel per launch, so neither "hbm" nor "tensor" binds it; the HBM figure is given
alongside), `cpu_baseline` (the C++ oracle port on all host cores, N=1 only), `e2e` (through
Tracer.render_spp with a page-locked host ColorBuffer: scene H2D + trace + reduce + image D2H per step).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SPP_PER_STEP = 3840, 2160, 128
METRIC, UNIT = "path_msamples_per_s_4k", "Msamples/s"
WORKLOAD = "AnalyticalScene (renderer/src/analytical.rs) 3840x2160, depth 4, f32 — BASELINE.json configs[2]"

# Algorithmic FLOP cost per call (1 FLOP = add/sub/mul/div/sqrt/min/max/abs/
# compare-select or one libm call, FMA = 2) for the demo scene; DESIGN.md "Work model".
F_K = dict(gen_ray=94, closest_hit=130, finalize=36, direct_light=127, nee_contrib=16, any_hit=45, eval_common=217,
           ev_diffuse=75, ev_reflect=103, ev_refract=100, ev_clearcoat=73, sample_common=200, lobe_diffuse=26,
           lobe_clearcoat=52, lobe_reflect=156, lobe_refract=169, background=18, glue_bounce=12, glue_sample=20)
# DRAM bytes per launch of the render kernel at 3840x2160 from the committed `ncu --set full` capture
# (profiles/r01_ncu_wavefront.md, r01-k: dram__bytes_read.sum 129.1 MB + dram__bytes_write.sum 114.1 MB; the accumulator
# read-modify-write is 265.4 MB algorithmic — part of the writes is still dirty in L2 when the kernel ends — and the sample
# blocks of the tail pixels add 44 MB of stores).  Independent of spp.
NCU_TRAFFIC_BYTES_PER_LAUNCH = 129124608 + 114142720


def flops_per_sample(c: dict) -> float:
    s = max(1, c["samples"])
    f = (F_K["gen_ray"] + F_K["glue_sample"]) * s
    f += (F_K["closest_hit"] + F_K["glue_bounce"]) * c["closest_hit"]
    f += F_K["finalize"] * (c["shade"] + c["end_emitter"])
    f += (F_K["direct_light"] + F_K["sample_common"]) * c["shade"] + F_K["nee_contrib"] * c["nee_contrib"]
    f += F_K["any_hit"] * c["any_hit"] + F_K["eval_common"] * c["eval_calls"]
    for k in ("ev_diffuse", "ev_reflect", "ev_refract", "ev_clearcoat", "lobe_diffuse", "lobe_clearcoat", "lobe_reflect", "lobe_refract"):
        f += F_K[k] * c[k]
    f += F_K["background"] * c["end_sky"]
    return f / s


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self) -> dict:
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def fp32_peak_tflops(device: int):
    path = os.path.join(ROOT, "tools", "libfp32peak.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    tf, ms, sms = C.c_double(), C.c_double(), C.c_int()
    if lib.fp32_peak(device, C.byref(tf), C.byref(ms), C.byref(sms)) != 0:
        return None
    return tf.value, sms.value


def cpu_baseline(seconds_budget: float = 12.0, threads: int = 0) -> dict:
    """The reference algorithm (C++ oracle port: Rust is not installable here) on the host cores,
    on a bounded sample of the SAME workload: whole 4K frames at 1 spp each."""
    from oracle import pyoracle as po
    import rust_pathtracer_b200 as rp
    sc = po.OracleScene(rp.AnalyticalScene.new().device_export())
    cores = threads or po.max_threads()
    px, frames, secs, ctr = sc.render(WIDTH, HEIGHT, 1, threads=cores, counters=True)      # also warms the thread pool
    n_frames = max(1, min(64, int(seconds_budget / max(secs, 1e-3))))
    px, frames, secs, _ = sc.render(WIDTH, HEIGHT, n_frames, threads=cores)
    samples = WIDTH * HEIGHT * n_frames
    return {"value": samples / secs / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_frames} frames of 1 spp at {WIDTH}x{HEIGHT} ({samples / 1e6:.1f} Msamples, {secs:.1f} s), C++ oracle port of "
                      f"tracer.rs, OpenMP schedule(dynamic,1) over rows",
            "flops_per_sample": flops_per_sample({**ctr, "end_rr": 0})}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    import rust_pathtracer_b200 as rp
    sc = po.OracleScene(rp.AnalyticalScene.new().device_export())
    cores = po.max_threads()
    px = None
    frames = 0
    for _ in range(max(0, args.warmup)):
        sc.render(WIDTH // 4, HEIGHT // 4, 1, threads=cores)          # warm-up: thread pool + caches (bounded)
    t_total = 0.0
    for _ in range(args.steps):
        px, frames, secs, _ = sc.render(WIDTH, HEIGHT, 1, pixels=px, frames=frames, threads=cores)   # one bounded step = 1 spp at 4K
        t_total += secs
    samples = WIDTH * HEIGHT * args.steps
    v = samples / t_total / 1e6
    sample = f"each step = 1 spp over the {WIDTH}x{HEIGHT} frame (a 1/{SPP_PER_STEP} sample of the 128-spp step; throughput is spp-independent)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_total / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "step": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_own(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import rust_pathtracer_b200 as rp
    from rust_pathtracer_b200.distributed import DistributedTracer, split_samples, resolve_mean

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    scene = rp.AnalyticalScene.new()
    W, H, S = args.width, args.height, args.spp_per_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FLOP model from device counters (small counted render of the same scene) -------------------
    fps = None
    rays_per_sample = None
    if rank == 0:
        ct = rp.Tracer.new(scene, device=local, collect_counters=True)
        cb = rp.ColorBuffer.new(960, 540)
        ct.render_spp(cb, 4, download=False)
        cts = ct.counters()
        fps = flops_per_sample(cts)
        rays_per_sample = (cts["closest_hit"] + cts["any_hit"]) / max(1, cts["samples"])
        ct.close()

    # ---- device-resident arm: `value` ---------------------------------------------------------------------
    integ = {"auto": rp._abi.PTB_INTEGRATOR_AUTO, "fused": rp._abi.PTB_INTEGRATOR_FUSED, "wavefront": rp._abi.PTB_INTEGRATOR_WAVEFRONT,
             "stream": rp._abi.PTB_INTEGRATOR_STREAM}[args.integrator]
    kernel_name = {"fused": "k_render_fused<float,false,false>", "stream": "k_stream_* (one kernel per stage)"}.get(args.integrator, "k_render_wavefront<COUNT=false,BVH=false,RM=true>")
    dt = DistributedTracer(scene, W, H, device=dev, integrator=integ, gather=args.gather)
    gather_used = dt.gather if world > 1 else "none (single GPU)"
    stream = torch.cuda.current_stream(dev)
    reduced = None
    for _ in range(args.warmup):
        dt.render(S)
        reduced = dt.reduce(0)
    barrier()
    dt.accum.zero_(); dt.samples_done = 0
    launches0 = dt.tracer.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record(stream)
    for _ in range(args.steps):
        dt.render(S)                       # this rank's share of the step's samples (k_render_fused)
        reduced = dt.reduce(0)             # ONE NCCL sum-reduce of the float4 accumulators per step
        if args.per_kernel_timing:
            kernel_ms.append(dt.tracer.last_render_ms())
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop() if sampler else None
    launches = dt.tracer.launch_count() - launches0
    # per-launch duration of the dominant kernel, CUDA events on the launch stream (untimed extra steps so the
    # event syncs do not perturb the timed region)
    if not kernel_ms:
        for _ in range(3):
            dt.render(S)
            kernel_ms.append(dt.tracer.last_render_ms())
    kms = float(np.mean(kernel_ms))
    samples_per_step = W * H * S
    value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e6
    image_ok = None
    if rank == 0 and reduced is not None:
        mean = resolve_mean(reduced)
        image_ok = bool(torch.isfinite(mean).all().item() and abs(float(mean.view(-1, 4)[:, 3].mean().item()) - 1.0) < 1e-6)
    dt.close()

    # ---- e2e arm: through Tracer.render_spp with HOST buffers ----------------------------------------------
    # per step: scene export H2D, this rank's share of the samples, NCCL reduce, and on rank 0 the D2H of the
    # running-mean image into a page-locked ColorBuffer.
    # The D2H of step k overlaps the tracing of step k+1: two page-locked host buffers alternate, the copy runs on a side
    # stream behind an event, and the timed region ends only when the last copy has landed.
    pinned = [torch.empty(W * H * 4, dtype=torch.float32).pin_memory() for _ in range(2)]
    host_bufs = [rp.ColorBuffer.new(W, H, storage=p_.numpy()) for p_ in pinned]
    et = DistributedTracer(scene, W, H, device=dev, integrator=integ, gather=args.gather)
    scene_bytes = et.tracer.scene_bytes
    copy_stream = torch.cuda.Stream(device=dev)
    step_no = [0]

    def e2e_step():
        main = torch.cuda.current_stream(dev)
        et.tracer.sync_scene()                                  # H2D: the step's input (the scene description)
        et.render(S)
        out = et.reduce(0)
        if rank == 0:
            mean = resolve_mean(out)
            ready = torch.cuda.Event()
            ready.record(main)
            k = step_no[0] & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                pinned[k].copy_(mean, non_blocking=True)        # D2H: the step's result into ColorBuffer.pixels
                mean.record_stream(copy_stream)
            host_bufs[k].frames = et.samples_done
        step_no[0] += 1

    def e2e_drain():
        copy_stream.synchronize()
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(min(2, args.warmup)):
        e2e_step()
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_drain()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_step * args.steps / float(e2e_s.item()) / 1e6
    et.close()

    if rank == 0:
        peak = fp32_peak_tflops(local)
        base, cnt = split_samples(S, world, 0)
        roofline = None
        if fps is not None:
            ach = fps * (W * H * cnt) / (kms * 1e-3) / 1e12
            nominal = 148 * 128 * 2 * 1.965e9 / 1e12
            pk = peak[0] if peak else nominal
            roofline = {"bound": "fp32", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk,
                        "peak_source": "measured live: tools/fp32_peak.cu FMA saturation" if peak else "nominal SMs*128*2*f_max",
                        "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH if (W, H) == (WIDTH, HEIGHT) else None, "kernel": kernel_name, "kernel_ms": kms,
                        "flops_per_sample": fps, "samples_per_launch": W * H * cnt,
                        "hbm": {"algorithmic_bytes_per_launch": W * H * 32, "achieved_gbs": W * H * 32 / (kms * 1e-3) / 1e9,
                                "peak_gbs": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                                if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0}}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": WORKLOAD if (W, H) == (WIDTH, HEIGHT) else f"AnalyticalScene {W}x{H}, depth 4, f32 (non-default size)",
                          "step": f"{S} spp over the frame, sample-split over {world} GPU(s), partial sums gathered on rank 0 (config.gather: peer = stored "
                                  f"over NVLink by the render kernel + one summing kernel; nccl = ONE NCCL sum-reduce of the float4 accumulators) "
                                  f"({args.steps} steps = {S * args.steps} spp)",
                          "integrator": kernel_name, "gather": gather_used, "l2": f"accumulators {W * H * 16 / 1e6:.1f} MB > 126 MB L2; "
                          "no other input", "image_finite_alpha_one": image_ok},
               "rays_per_s": value * 1e6 * rays_per_sample if rays_per_sample else None,   # closest_hit + any_hit calls per second, whole job
               "clocks": clocks, "gpu_launches": int(launches),
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": W * H * 16},
               "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--spp-per-step", type=int, default=SPP_PER_STEP)
    ap.add_argument("--integrator", default="auto", choices=["auto", "fused", "wavefront", "stream"])
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"],
                    help="multi-GPU: partial sums stored into rank 0's memory by the render kernel (peer) or one NCCL reduce per step (nccl)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-kernel-timing", action="store_true", help="sync after every step to time each launch (perturbs `value`)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 3) if not os.environ.get("PTB_BENCH_ALLOW_SHORT_WARMUP") else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
