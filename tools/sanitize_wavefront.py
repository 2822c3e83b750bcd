"""Small renders through the shared-memory wavefront integrator for compute-sanitizer (memcheck / racecheck / initcheck):
the resolved-material instantiation (demo scene), the generic one (same scene with PTB200_NO_RESOLVED_MATERIALS=1) and the
BVH one; frame sizes where every pixel is a tail pixel, where only part of the frame is, spp not a multiple of the block count,
and 1 spp (no tail items)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_pathtracer_b200 as rp
WF = rp._abi.PTB_INTEGRATOR_WAVEFRONT
big = len(sys.argv) > 1 and sys.argv[1] == "big"
cases = [(97, 61, 11), (64, 48, 1), (160, 90, 3)] + ([(1280, 720, 2)] if big else [])
for env in (None, "1"):
    if env is None:
        os.environ.pop("PTB200_NO_RESOLVED_MATERIALS", None)
    else:
        os.environ["PTB200_NO_RESOLVED_MATERIALS"] = env
    for (w, h, s) in cases:
        pt = rp.Tracer.new(rp.AnalyticalScene.new(), integrator=WF)
        buf = rp.ColorBuffer.new(w, h)
        pt.render_spp(buf, s)
        pt.render_spp(buf, s)
        pt.close()
os.environ.pop("PTB200_NO_RESOLVED_MATERIALS", None)
pt = rp.Tracer.new(rp.divergence_stress_scene(side=8, depth=6), integrator=WF, rr_start=3)
buf = rp.ColorBuffer.new(96, 54)
pt.render_spp(buf, 3)
pt.close()
# round 2: the f64 instantiation, a signed-distance scene (generic kernel, f32 and f64) and an emissive material (the
# emission read-modify-write on the slot's radiance before shading)
pt = rp.Tracer.new(rp.AnalyticalScene.new(), integrator=WF, precision="f64")
buf = rp.ColorBuffer.new(97, 61, "f64")
pt.render_spp(buf, 5)
pt.close()
for prec in ("f32", "f64"):
    pt = rp.Tracer.new(rp.sdf_demo_scene(), integrator=WF, precision=prec)
    buf = rp.ColorBuffer.new(64, 36, prec)
    pt.render_spp(buf, 2)
    pt.close()
ex = rp.AnalyticalScene.new().device_export()
ex.materials[1].emission = rp.F3(0.5, 0.2, 0.1)
pt = rp.Tracer.new(rp.ExportedScene(ex), integrator=WF)
buf = rp.ColorBuffer.new(80, 45)
pt.render_spp(buf, 3)
pt.close()
# round 2 (later): media (generic wavefront + fused, f32 and f64) and the extended light kinds
for prec in ("f32", "f64"):
    for integ in (WF, rp._abi.PTB_INTEGRATOR_FUSED):
        for scene in (rp.media_demo_scene(depth=8), rp.lights_demo_scene()):
            pt = rp.Tracer.new(scene, integrator=integ, precision=prec)
            buf = rp.ColorBuffer.new(64, 36, prec)
            pt.render_spp(buf, 2)
            pt.close()
print("ok")
