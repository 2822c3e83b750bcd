"""Scratch GPU check: timing of the integrators (not a test, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rust_pathtracer_b200 as rp

scene = rp.AnalyticalScene.new()
cfgs = [(1920, 1080, 64), (3840, 2160, 64)]
if len(sys.argv) > 1:
    cfgs = [tuple(int(x) for x in sys.argv[1].split("x"))]
for integ in (rp._abi.PTB_INTEGRATOR_FUSED, rp._abi.PTB_INTEGRATOR_WAVEFRONT):
    for (W, H, spp) in cfgs:
        try:
            pt = rp.Tracer.new(scene, integrator=integ)
            buf = rp.ColorBuffer.new(W, H)
            pt.render_spp(buf, 4, download=False)
            pt.synchronize()
            best = 1e9
            for rep in range(3):
                pt.render_spp(buf, spp, download=False)
                best = min(best, pt.last_render_ms())
            print(f"integrator={integ} {W}x{H} spp={spp}: {best:.2f} ms  -> {W*H*spp/best/1e3:.1f} Msamples/s", flush=True)
            pt.close()
        except Exception as e:
            print(f"integrator={integ}: {e}")
