"""The five BASELINE.json configs on one GPU (parity-tested separately; this reports throughput).
Writes one JSON line per (config, integrator) to stdout.  Not the headline bench (bench.py)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rust_pathtracer_b200 as rp

A = rp._abi


def run(name, scene, W, H, spp, integ, reps=3, **kw):
    pt = rp.Tracer.new(scene, integrator=integ, **kw)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, max(1, min(4, spp)), download=False)
    pt.synchronize()
    best = 1e30
    for _ in range(reps):
        pt.render_spp(buf, spp, download=False)
        best = min(best, pt.last_render_ms())
    pt.close()
    # rays per sample from a counted pass at reduced size
    pc = rp.Tracer.new(scene, integrator=integ, collect_counters=True, **kw)
    cb = rp.ColorBuffer.new(max(16, W // 8), max(16, H // 8))
    pc.render_spp(cb, max(1, min(4, spp)), download=False)
    c = pc.counters()
    pc.close()
    rays = (c["closest_hit"] + c["any_hit"]) / max(1, c["samples"])
    ms = W * H * spp / best / 1e3
    return {"config": name, "integrator": {A.PTB_INTEGRATOR_FUSED: "fused", A.PTB_INTEGRATOR_WAVEFRONT: "wavefront", A.PTB_INTEGRATOR_STREAM: "stream"}[integ], "width": W, "height": H,
            "spp": spp, "kernel_ms": round(best, 3), "msamples_per_s": round(ms, 1), "rays_per_sample": round(rays, 3),
            "grays_per_s": round(ms * rays / 1e3, 3), "bounces_per_sample": round(c["closest_hit"] / max(1, c["samples"]), 3)}


def drop_in(name, scene, W, H, calls):
    """config 1: `calls` drop-in Tracer::render calls (1 spp each, host pixels round trip)"""
    pt = rp.Tracer.new(scene)
    buf = rp.ColorBuffer.new(W, H)
    pt.render(buf)
    t0 = time.perf_counter()
    for _ in range(calls):
        pt.render(buf)
    dt = time.perf_counter() - t0
    pt.close()
    return {"config": name, "integrator": "auto", "width": W, "height": H, "spp": calls, "wall_ms_per_call": round(dt / calls * 1e3, 3),
            "msamples_per_s": round(W * H * calls / dt / 1e6, 1), "note": "drop-in render(): H2D of pixels + 1 spp + D2H per call"}


def main():
    which = sys.argv[1:] or ["1", "2", "3", "4", "5"]
    demo = rp.AnalyticalScene.new()
    out = []
    if "1" in which:
        out.append(drop_in("cfg1 AnalyticalScene 800x600, 64 drop-in render() calls, f32", demo, 800, 600, 64))
        pt = rp.Tracer.new(demo, precision="f64")
        buf = rp.ColorBuffer.new(800, 600, "f64")
        pt.render_spp(buf, 2, download=False); pt.synchronize()
        pt.render_spp(buf, 16, download=False)
        ms = pt.last_render_ms()
        out.append({"config": "cfg1 AnalyticalScene 800x600 f64 (F switch), 16 spp", "integrator": "fused", "kernel_ms": round(ms, 3),
                    "msamples_per_s": round(800 * 600 * 16 / ms / 1e3, 1)})
        pt.close()
    integs = (A.PTB_INTEGRATOR_FUSED, A.PTB_INTEGRATOR_WAVEFRONT, A.PTB_INTEGRATOR_STREAM)
    if os.environ.get("PTB_INTEGRATORS"):
        integs = tuple(int(x) for x in os.environ["PTB_INTEGRATORS"].split(","))
    for integ in integs:
        if "2" in which:
            out.append(run("cfg2 AnalyticalScene 1920x1080, 256 spp", demo, 1920, 1080, 256, integ))
        if "3" in which:
            out.append(run("cfg3 AnalyticalScene 3840x2160, 128 spp (one bench step)", demo, 3840, 2160, 128, integ))
        if "4" in which:
            t0 = time.time()
            field = rp.sphere_field_scene()
            out.append(run(f"cfg4 sphere field 100k spheres + 64 lights + BVH, 3840x2160, 8 spp (scene built in {time.time() - t0:.1f}s)",
                           field, 3840, 2160, 8, integ, reps=2))
        if "5" in which:
            stress = rp.divergence_stress_scene(side=64, depth=16)
            out.append(run("cfg5 divergence stress 4096 spheres, depth 16, RR from bounce 3, 3840x2160, 16 spp", stress, 3840, 2160, 16, integ,
                           reps=2, rr_start=3))
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
