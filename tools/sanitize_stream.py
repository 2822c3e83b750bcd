"""Small renders through the streaming integrator for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_pathtracer_b200 as rp
S = rp._abi.PTB_INTEGRATOR_STREAM
for scene, kw in ((rp.AnalyticalScene.new(), {}), (rp.lights_demo_scene(), {}), (rp.sphere_field_scene(n_spheres=20000, n_lights_side=5), {}),
                  (rp.divergence_stress_scene(side=8, depth=16), {"rr_start": 3})):
    for wave in (0, 4096):
        pt = rp.Tracer.new(scene, integrator=S, wave_paths=wave, **kw)
        buf = rp.ColorBuffer.new(96, 54)
        pt.render_spp(buf, 3)
        pt.close()
print("ok")
