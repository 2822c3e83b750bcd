"""Scratch: run one synthetic config once for ncu.  usage: prof_cfg.py <4|5> <integrator> WxHxSPP"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_pathtracer_b200 as rp
cfg, integ = int(sys.argv[1]), int(sys.argv[2]); W, H, spp = (int(x) for x in sys.argv[3].split("x"))
scene = rp.sphere_field_scene() if cfg == 4 else rp.divergence_stress_scene(side=64, depth=16)
pt = rp.Tracer.new(scene, integrator=integ, rr_start=3 if cfg == 5 else 0)
buf = rp.ColorBuffer.new(W, H)
pt.render_spp(buf, 1, download=False)
pt.render_spp(buf, spp, download=False)
print(pt.last_render_ms())
if os.environ.get("PTB_PROF_COUNT"):
    ct = rp.Tracer.new(scene, integrator=integ, rr_start=3 if cfg == 5 else 0, collect_counters=True)
    ct.render_spp(buf, spp, download=False)
    c = ct.counters()
    print("RAYS", c["closest_hit"] + c["any_hit"], "SAMPLES", c["samples"])
