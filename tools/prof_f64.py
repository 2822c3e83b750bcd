"""Scratch: f64 throughput of the fused and the wavefront integrator.  usage: prof_f64.py WxHxSPP"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_pathtracer_b200 as rp
W, H, spp = (int(x) for x in sys.argv[1].split("x"))
out = {}
for name, integ in (("fused", rp._abi.PTB_INTEGRATOR_FUSED), ("wavefront", rp._abi.PTB_INTEGRATOR_WAVEFRONT)):
    pt = rp.Tracer.new(rp.AnalyticalScene.new(), precision="f64", integrator=integ)
    buf = rp.ColorBuffer.new(W, H, "f64")
    pt.render_spp(buf, 2, download=False)
    ms = []
    for _ in range(3):
        pt.render_spp(buf, spp, download=False)
        ms.append(pt.last_render_ms())
    out[name] = {"ms": min(ms), "msamples_s": W * H * spp / min(ms) / 1e3, "kernel": pt.integrator_used()}
    pt.close()
import ctypes as C
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfp32peak.so"))
tf, ms, sms = C.c_double(), C.c_double(), C.c_int()
peak = tf.value if lib.fp64_peak(0, C.byref(tf), C.byref(ms), C.byref(sms)) == 0 else None
peak = tf.value
print(json.dumps({"W": W, "H": H, "spp": spp, "fp64_peak_tflops_measured": peak, **out}))
