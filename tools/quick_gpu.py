"""Scratch GPU check: smoke + timing of the fused integrator (not a test, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
import rust_pathtracer_b200 as rp

g.smoke()
scene = rp.AnalyticalScene.new()
for (W, H, spp) in [(1920, 1080, 64), (3840, 2160, 32), (3840, 2160, 128)]:
    pt = rp.Tracer.new(scene)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, 4, download=False)
    pt.synchronize()
    for rep in range(3):
        pt.render_spp(buf, spp, download=False)
        ms = pt.last_render_ms()
        print(f"{W}x{H} spp={spp}: {ms:.2f} ms  -> {W*H*spp/ms/1e3:.1f} Msamples/s", flush=True)
    pt.close()
ptc = rp.Tracer.new(scene, collect_counters=True)
buf = rp.ColorBuffer.new(800, 600)
ptc.render_spp(buf, 16)
c = ptc.counters(); s = c["samples"]
print({k: round(v / s, 4) for k, v in c.items()})
img = buf.pixels.reshape(600, 800, 4)
print("mean rgb", img[..., :3].reshape(-1, 3).mean(0))
