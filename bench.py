#!/usr/bin/env python
"""bench.py — path Msamples/s on 1/2/4/8 B200 for the five BASELINE.json configs (default: config 3, the headline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config {1..5}]     # own arm (CUDA, sm_100a)
    python bench.py --impl reference [...]                                    # the reference algorithm on the host cores

A STEP is one pass of the hot path over the frame: `spp_per_step` samples per pixel, sample-split across the ranks, the partial
sums gathered on rank 0.  The five workloads (config.workload):

  1  AnalyticalScene 800x600, a step = 64 drop-in `Tracer::render()` calls of 1 spp (host pixels in and out every call)
  2  AnalyticalScene 1920x1080, 256 spp in one batch
  3  AnalyticalScene 3840x2160, 128 spp per step (8 steps = the config's 1024 spp)                  <- default / headline
  4  procedural sphere field: 100 000 spheres, mixed Disney materials, 64 spherical lights, sphere BVH, 3840x2160, 8 spp per step
  5  divergence stress: 4 096 rough / transmissive spheres, depth 16, Russian roulette from bounce 3, 3840x2160, 16 spp per step

`value` is device-resident (inputs in HBM, CUDA events on the launch stream, max over ranks).  `e2e` goes through the
reference-facing host API — Tracer / DistributedTracer over the C ABI — with HOST buffers: per step the scene description is
uploaded (ptb_set_scene), the samples are traced (ptb_render), and rank 0 downloads the running-mean image into a page-locked
`ColorBuffer.pixels` (ptb_download_async: k_resolve on the render stream, D2H on a side stream, overlapped with the next step's
tracing; the timed region ends when the last copy has landed).  No torch kernel runs inside a step at N = 1.

`roofline`: the path has no dense contraction and moves 32 B per PIXEL per launch, so neither "hbm" nor "tensor" binds configs
1-3 and 5: the bound is the FP32 pipe (algorithmic FLOP per sample from the device's own event counters x SURVEY App. C costs).
Config 4 is bound by BVH node traffic through L1/L2: its roofline is reported as algorithmic node + leaf bytes per second
against the L2 bandwidth.  `cpu_baseline`: the C++ oracle port on all host cores (N = 1 only).
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

UNIT = "Msamples/s"

# Algorithmic FLOP cost per call (1 FLOP = add/sub/mul/div/sqrt/min/max/abs/
# compare-select or one libm call, FMA = 2); DESIGN.md "Work model".  closest_hit / any_hit are the demo scene's linear scans
# (2 spheres + plane + 1 light); BVH scenes replace them by per-node and per-leaf-sphere costs.
F_K = dict(gen_ray=94, closest_hit=130, finalize=36, direct_light=127, nee_contrib=16, any_hit=45, eval_common=217,
           ev_diffuse=75, ev_reflect=103, ev_refract=100, ev_clearcoat=73, sample_common=200, lobe_diffuse=26,
           lobe_clearcoat=52, lobe_reflect=156, lobe_refract=169, background=18, glue_bounce=12, glue_sample=20,
           bvh_node=32, bvh_sphere=28, plane=12, light_sphere=30)


def flops_per_sample(c: dict, bvh: dict = None) -> float:
    s = max(1, c["samples"])
    f = (F_K["gen_ray"] + F_K["glue_sample"]) * s
    if bvh is None:
        f += (F_K["closest_hit"] + F_K["glue_bounce"]) * c["closest_hit"]
        f += F_K["any_hit"] * c["any_hit"]
    else:   # per ray: nodes visited and leaf spheres tested (measured by the counted pass), plus the plane and the light tests
        rays = c["closest_hit"] + c["any_hit"]
        f += F_K["glue_bounce"] * c["closest_hit"] + F_K["plane"] * rays
        f += F_K["bvh_node"] * bvh["nodes"] + F_K["bvh_sphere"] * bvh["leaf_tests"] + F_K["light_sphere"] * bvh.get("light_tests", 0)
    f += F_K["finalize"] * (c["shade"] + c["end_emitter"])
    f += (F_K["direct_light"] + F_K["sample_common"]) * c["shade"] + F_K["nee_contrib"] * c["nee_contrib"]
    f += F_K["eval_common"] * c["eval_calls"]
    for k in ("ev_diffuse", "ev_reflect", "ev_refract", "ev_clearcoat", "lobe_diffuse", "lobe_clearcoat", "lobe_reflect", "lobe_refract"):
        f += F_K[k] * c[k]
    f += F_K["background"] * c["end_sky"]
    return f / s


CONFIGS = {
    1: dict(metric="path_msamples_per_s", W=800, H=600, spp=64, drop_in=True, scene="demo",
            workload="AnalyticalScene (renderer/src/analytical.rs) 800x600, depth 4, f32, 64 drop-in Tracer::render() calls of 1 spp per step — BASELINE.json configs[0]"),
    2: dict(metric="path_msamples_per_s", W=1920, H=1080, spp=256, scene="demo",
            workload="AnalyticalScene 1920x1080, depth 4, f32, 256 spp per step — BASELINE.json configs[1]"),
    3: dict(metric="path_msamples_per_s_4k", W=3840, H=2160, spp=128, scene="demo",
            workload="AnalyticalScene (renderer/src/analytical.rs) 3840x2160, depth 4, f32 — BASELINE.json configs[2]"),
    4: dict(metric="path_msamples_per_s_4k", W=3840, H=2160, spp=8, scene="field",
            workload="procedural sphere field: 100k spheres, mixed Disney materials (metal/glass/clearcoat), 64 spherical lights, sphere BVH, "
                     "3840x2160, depth 4, f32, 8 spp per step — BASELINE.json configs[3]"),
    5: dict(metric="path_msamples_per_s_4k", W=3840, H=2160, spp=16, scene="stress", rr_start=3,
            workload="divergence stress: 4096 rough / transmissive spheres, depth 16, Russian roulette from bounce 3, 3840x2160, f32, "
                     "16 spp per step — BASELINE.json configs[4]"),
}


def make_scene(kind: str):
    import rust_pathtracer_b200 as rp
    if kind == "demo":
        return rp.AnalyticalScene.new()
    if kind == "field":
        return rp.sphere_field_scene()
    if kind == "stress":
        return rp.divergence_stress_scene(side=64, depth=16)
    raise ValueError(kind)


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


def host_cores() -> int:
    """the cores this process may run on — NOT omp_get_max_threads(), which honours the OMP_NUM_THREADS=1 that
    torch.distributed.run exports to its workers"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def oracle_throughput(cfg: dict, seconds_budget: float, steps: int = 0, warmup: int = 1):
    """The reference algorithm (C++ oracle port: Rust is not installable here) on all host cores, on a bounded sample of the SAME
    workload: whole frames of the config's size at 1 spp each.  Two RNGs: the reference's own kind — per-thread ChaCha12 drawn in
    call order (rand 0.8.5 thread_rng; tracer.rs:44) — which is the reported value, and the counter RNG the parity runs use (a
    Philox block per draw: dearer on a CPU).  Returns a dict."""
    from oracle import pyoracle as po
    sc = po.OracleScene(make_scene(cfg["scene"]).device_export())
    cores = host_cores()
    W, H = cfg["W"], cfg["H"]
    for _ in range(max(0, warmup)):
        sc.render(max(16, W // 4), max(16, H // 4), 1, threads=cores)      # warm-up: thread pool + caches (bounded)
        sc.render_chacha(max(16, W // 4), max(16, H // 4), 1, threads=cores)
    px, frames, secs, ctr = sc.render(W, H, 1, threads=cores, counters=True)
    if steps <= 0:
        steps = max(1, min(64, int(0.5 * seconds_budget / max(secs, 1e-3))))
    out = {"cores": cores, "steps": steps, "counters": ctr}
    sc.render_chacha(W, H, 1, threads=cores)                               # untimed: first full-size call of this entry point
    t_total, px, frames = 0.0, None, 0
    for _ in range(steps):
        px, frames, secs = sc.render_chacha(W, H, 1, pixels=px, frames=frames, threads=cores)
        t_total += secs
    out["chacha12"] = W * H * steps / t_total / 1e6
    out["seconds"] = t_total
    t_total, px, frames = 0.0, None, 0
    for _ in range(steps):
        px, frames, secs, _ = sc.render(W, H, 1, pixels=px, frames=frames, threads=cores)
        t_total += secs
    out["counter_rng"] = W * H * steps / t_total / 1e6
    out["seconds"] += t_total
    return out


def cpu_baseline(cfg: dict, seconds_budget: float = 12.0) -> dict:
    r = oracle_throughput(cfg, seconds_budget)
    return {"value": r["chacha12"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": f"{r['steps']} frames of 1 spp at {cfg['W']}x{cfg['H']} per RNG ({cfg['W'] * cfg['H'] * r['steps'] / 1e6:.1f} Msamples, {r['seconds']:.1f} s in all), "
                      f"C++ oracle port of tracer.rs, dyn-dispatched Scene, Material::new() per bounce, per-thread ChaCha12 in call order like rand's thread_rng, "
                      f"OpenMP schedule(dynamic,1) over rows like the reference's one-row rayon tasks",
            "counter_rng_value": r["counter_rng"],
            "counter_rng_note": "the same port on the Philox counter RNG of the parity runs (one block per draw)",
            "flops_per_sample": flops_per_sample({**r["counters"], "end_rr": 0}) if cfg["scene"] == "demo" else None}


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation of the path (oracle port, ChaCha12 like rand's thread_rng), all host
    threads; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    r = oracle_throughput(cfg, 0.0, steps=max(1, args.steps), warmup=max(0, args.warmup))
    v = r["chacha12"]
    sample = f"each step = 1 spp over the {cfg['W']}x{cfg['H']} frame (a 1/{cfg['spp']} sample of the {cfg['spp']}-spp step; throughput is spp-independent)"
    print(json.dumps({
        "impl": "reference", "metric": cfg["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cfg["W"] * cfg["H"] / (v * 1e6) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": cfg["workload"], "step": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample + "; per-thread ChaCha12 in call order",
                         "counter_rng_value": r["counter_rng"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def ncu_traffic(kernel_key: str, W: int, H: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json: one record per kernel and frame size with the git hash it was taken at); None if not captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    for rec in json.load(open(path)):
        if rec.get("kernel_key") == kernel_key and rec.get("width") == W and rec.get("height") == H:
            return int(rec["dram_bytes_read"]) + int(rec["dram_bytes_write"]), rec.get("source")
    return None, None


def run_own(args, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist
    import rust_pathtracer_b200 as rp
    from rust_pathtracer_b200.distributed import DistributedTracer, split_samples

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    scene = make_scene(cfg["scene"])
    W, H, S = args.width or cfg["W"], args.height or cfg["H"], args.spp_per_step or cfg["spp"]
    tracer_kw = dict(rr_start=cfg.get("rr_start", 0))
    A = rp._abi

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- work model from device counters (small counted render of the same scene) -------------------
    fps = rays_per_sample = bvh_stats = None
    if rank == 0:
        ct = rp.Tracer.new(scene, device=local, collect_counters=True, **tracer_kw)
        cb = rp.ColorBuffer.new(960, 540)
        ct.render_spp(cb, 4, download=False)
        cts = ct.counters()
        rays_per_sample = (cts["closest_hit"] + cts["any_hit"]) / max(1, cts["samples"])
        if cfg["scene"] != "demo" and cts.get("bvh_nodes"):
            # sphere-BVH work from the counting build of the same kernels: inner-node visits (two box tests each) and leaf sphere tests
            bvh_stats = {"nodes": cts["bvh_nodes"], "leaf_tests": cts["bvh_leaf_tests"], "nodes_per_ray": cts["bvh_nodes"] / max(1, cts["closest_hit"] + cts["any_hit"]),
                         "leaf_tests_per_ray": cts["bvh_leaf_tests"] / max(1, cts["closest_hit"] + cts["any_hit"]),
                         "note": "the light BVH's node visits (64 lights) are not counted: the FLOP figure is a lower bound"}
        fps = flops_per_sample(cts, bvh_stats) if (cfg["scene"] == "demo" or bvh_stats) else None
        ct.close()

    integ = {"auto": A.PTB_INTEGRATOR_AUTO, "fused": A.PTB_INTEGRATOR_FUSED, "wavefront": A.PTB_INTEGRATOR_WAVEFRONT,
             "stream": A.PTB_INTEGRATOR_STREAM}[args.integrator]
    kernel_names = {"fused": "k_render_fused<float,COUNT=false,BVH=false>", "fused_bvh": "k_render_fused<float,COUNT=false,BVH=true>",
                    "stream": "k_stream_* (one kernel per stage: generate, closest, shade, shadow, accumulate)",
                    "stream_bvh": "k_stream_* with BVH traversal (one kernel per stage: generate, closest, shade, shadow, accumulate)",
                    "stream_split_bvh": "k_stream_* (one kernel per stage; bounce >= 1: k_stream_trace + k_stream_finish, shadow rays k_stream_trace<ANY>)",
                    "wavefront": "k_render_wavefront<COUNT=false,BVH=false,RM=false>", "wavefront_bvh": "k_render_wavefront<COUNT=false,BVH=true,RM=false>",
                    "wavefront_rm": "k_render_wavefront<COUNT=false,BVH=false,RM=true>"}

    # ---- device-resident arm: `value` ---------------------------------------------------------------------
    dt = DistributedTracer(scene, W, H, device=dev, integrator=integ, gather=args.gather, **tracer_kw)
    gather_used = dt.gather if world > 1 else "none (single GPU)"
    stream = torch.cuda.current_stream(dev)
    for _ in range(args.warmup):
        dt.render(S)
        dt.reduce(0)
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
        dt.render(S)                       # this rank's share of the step's samples
        dt.reduce(0)                       # partial sums onto rank 0 (peer stores fused into the render kernel, or ONE NCCL reduce)
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
    kernel_key = dt.tracer.integrator_used()          # what AUTO resolved to, reported by the library (ptb_last_integrator)
    kernel_name = kernel_names.get(kernel_key, kernel_key)
    # per-launch duration of the render kernel(s): CUDA events on the launch stream around the last ptb_render, taken on 3 extra
    # UNTIMED steps after the timed region so that the event synchronisations do not perturb `value`
    if not kernel_ms:
        for _ in range(3):
            dt.render(S)
            kernel_ms.append(dt.tracer.last_render_ms())
            dt.reduce(0)
    kms = float(np.mean(kernel_ms))
    samples_per_step = W * H * S
    value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e6
    image_ok = None
    if rank == 0:
        chk = rp.ColorBuffer.new(W, H)
        dt.tracer.download(chk)
        px = chk.read_pixels().reshape(-1, 4)
        image_ok = bool(np.isfinite(px).all() and abs(float(px[:, 3].mean()) - 1.0) < 1e-6)
        if cfg["scene"] == "demo" and not image_ok:     # the reference never filters NaN (tracer.rs:105-117): say how many pixels
            image_ok = f"{int((~np.isfinite(px).all(1)).sum())} non-finite pixels (0/0 at exactly grazing clearcoat samples, as in the reference)"
    dt.close()

    # ---- multi-GPU: the image must not depend on the number of ranks -----------------------------------
    n_inv = None
    if world > 1:
        w2, h2, s2 = 512, 288, 8
        nt = DistributedTracer(scene, w2, h2, device=dev, integrator=integ, gather=args.gather, **tracer_kw)
        nt.render(s2)
        nt.reduce(0)
        torch.cuda.synchronize()
        if rank == 0:
            many = rp.ColorBuffer.new(w2, h2)
            nt.tracer.download(many)
            single = rp.Tracer.new(scene, device=local, integrator=integ, **tracer_kw)
            one = rp.ColorBuffer.new(w2, h2)
            single.render_spp(one, s2)
            single.close()
            a_, b_ = many.read_pixels().reshape(-1, 4)[:, :3].astype(np.float64), one.read_pixels().reshape(-1, 4)[:, :3].astype(np.float64)
            n_inv = float((np.abs(a_ - b_).max(1) / np.maximum(np.abs(b_).max(1), 1e-3)).max())
        nt.close()

    # ---- e2e arm: the host API over the C ABI with HOST buffers -----------------------------------------
    if cfg.get("drop_in"):
        # config 1: the reference's loop — one Tracer::render(&mut ColorBuffer) per sample, pixels in host memory
        et = rp.Tracer.new(scene, device=local, integrator=integ, **tracer_kw)
        hb = rp.ColorBuffer.new(W, H)
        for _ in range(4):
            et.render(hb)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps * S):
            et.render(hb)
        e2e_s = time.perf_counter() - t0
        # the same loop when the app edits `pixels` between calls (every call uploads the host image first)
        hb2 = rp.ColorBuffer.new(W, H)
        _ = hb2.pixels
        for _ in range(4):
            et.render(hb2)
        t0 = time.perf_counter()
        for _ in range(args.steps * S):
            et.render(hb2)
        e2e_edit_s = time.perf_counter() - t0
        scene_bytes, h2d, d2h = et.scene_bytes, 0, W * H * 16 * S
        e2e_note = {"drop_in_ms_per_call": e2e_s / (args.steps * S) * 1e3, "with_upload_every_call_ms": e2e_edit_s / (args.steps * S) * 1e3,
                    "with_upload_every_call_msamples": samples_per_step * args.steps / e2e_edit_s / 1e6}
        et.close()
    else:
        et = DistributedTracer(scene, W, H, device=dev, integrator=integ, gather=args.gather, **tracer_kw)
        scene_bytes = et.tracer.scene_bytes
        host_bufs = [rp.ColorBuffer.new(W, H) for _ in range(2)]            # page-locked by the wrapper on first use (ptb_pin_host)
        step_no = [0]

        pod = et.tracer.prepare_scene()                         # the POD arrays a Rust / C++ host holds; uploaded every step

        def e2e_step():
            et.tracer.sync_scene(pod)                           # H2D: the step's input (the scene description)
            et.step(S, host_bufs[step_no[0] & 1])               # trace + gather + (rank 0) asynchronous D2H of the mean image
            step_no[0] += 1

        def e2e_drain():
            et.tracer.wait_download()
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
        e2e_s = time.perf_counter() - t0
        h2d, d2h, e2e_note = int(scene_bytes), W * H * 16, None
        et.close()
    e2e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_step * args.steps / float(e2e_t.item()) / 1e6

    if rank == 0:
        peak = fp32_peak_tflops(local)
        base, cnt = split_samples(S, world, 0)
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        roofline = None
        traffic, traffic_src = ncu_traffic(kernel_key, W, H)
        if fps is not None:
            ach = fps * (W * H * cnt) / (kms * 1e-3) / 1e12
            nominal = 148 * 128 * 2 * 1.965e9 / 1e12
            pk = peak[0] if peak else nominal
            roofline = {"bound": "fp32", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk,
                        "peak_source": "measured live: tools/fp32_peak.cu FMA saturation (MEASURED_PEAKS.json carries no FP32 figure)" if peak
                        else "nominal SMs*128*2*f_max",
                        "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name, "kernel_ms": kms,
                        "kernel_ms_source": "mean of 3 untimed extra steps after the timed region (CUDA events around ptb_render on the launch stream)",
                        "flops_per_sample": fps, "samples_per_launch": W * H * cnt,
                        "hbm": {"algorithmic_bytes_per_launch": W * H * 32, "achieved_gbs": W * H * 32 / (kms * 1e-3) / 1e9, "peak_gbs": hbm_peak}}
            if bvh_stats:
                roofline["bvh"] = bvh_stats
        out = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": cfg["workload"] if (W, H) == (cfg["W"], cfg["H"]) else f"{cfg['workload']} — at non-default size {W}x{H}",
                          "bench_config": args.config,
                          "step": f"{S} spp over the frame, sample-split over {world} GPU(s), partial sums gathered on rank 0 (config.gather: peer = stored "
                                  f"over NVLink by the render kernel + one summing kernel; nccl = ONE NCCL sum-reduce of the float4 accumulators) "
                                  f"({args.steps} steps = {S * args.steps} spp)",
                          "integrator": kernel_name, "gather": gather_used, "l2": f"accumulators {W * H * 16 / 1e6:.1f} MB "
                          + ("> 126 MB L2" if W * H * 16 > 126e6 else "< 126 MB L2 (read and written once per launch; the path is compute-bound)") + "; no other input",
                          "image_finite_alpha_one": image_ok, "n_invariance_max_rel": n_inv,
                          "n_invariance": None if n_inv is None else "512x288, 8 spp: N ranks vs rank 0 alone, max over pixels of |a-b|/max(|b|,1e-3); bar 1e-5 "
                                                                      "(same samples, f32 summation order differs)"},
               "rays_per_s": value * 1e6 * rays_per_sample if rays_per_sample else None,   # closest_hit + any_hit calls per second, whole job
               "clocks": clocks, "gpu_launches": int(launches),
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "path": "Tracer.render(ColorBuffer) x spp (ptb_render_frame_ex_f32)" if cfg.get("drop_in") else
                               "Tracer.sync_scene (ptb_set_scene_f32) + DistributedTracer.step (ptb_render [+ gather] + ptb_download_async_f32 into ColorBuffer.pixels)"},
               "roofline": roofline}
        if e2e_note:
            out["e2e"].update(e2e_note)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(cfg)
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
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json config (1-based); 3 is the headline")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp-per-step", type=int, default=0)
    ap.add_argument("--integrator", default="auto", choices=["auto", "fused", "wavefront", "stream"])
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"],
                    help="multi-GPU: partial sums stored into rank 0's memory by the render kernel (peer) or one NCCL reduce per step (nccl)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-kernel-timing", action="store_true", help="sync after every step to time each launch (perturbs `value`)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 3) if not os.environ.get("PTB_BENCH_ALLOW_SHORT_WARMUP") else args.warmup
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_own(args, cfg)


if __name__ == "__main__":
    main()
