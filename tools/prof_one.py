"""Scratch: run ONE integrator at one size for ncu.  usage: prof_one.py <integrator 1|2> WxHxSPP"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_pathtracer_b200 as rp
integ = int(sys.argv[1]); W, H, spp = (int(x) for x in sys.argv[2].split("x"))
pt = rp.Tracer.new(rp.AnalyticalScene.new(), integrator=integ)
buf = rp.ColorBuffer.new(W, H)
pt.render_spp(buf, 2, download=False)
pt.render_spp(buf, spp, download=False)
print(pt.last_render_ms())
