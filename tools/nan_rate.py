"""How often does the device path poison a pixel with NaN?  (the oracle shows ~13 NaN pixels per 8.5e9 samples
of the demo scene — an exactly grazing clearcoat sample divides 0 by 0, tracer.rs:414-418 — and the reference never filters NaN.)
Renders WxH at `spp` with independent seeds and counts non-finite pixels.  GPU only; prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rust_pathtracer_b200 as rp

W, H, spp, reps = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (3840, 2160, 1024, 2)))
scene = rp.AnalyticalScene.new()
out = []
for k in range(reps):
    pt = rp.Tracer.new(scene, seed=1000 + k)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, spp)
    ms = pt.last_render_ms()
    px = buf.read_pixels().reshape(-1, 4)
    out.append({"seed": 1000 + k, "nan_pixels": int((~np.isfinite(px).all(1)).sum()), "kernel_ms": ms, "msamples_s": W * H * spp / ms / 1e3,
                "mean_lum": float(np.nanmean(px[:, :3] @ np.array([0.212671, 0.715160, 0.072169], np.float32)))})
    pt.close()
print(json.dumps({"lib": os.environ.get("PTB200_LIB", "default"), "W": W, "H": H, "spp": spp, "samples_per_run": W * H * spp, "runs": out}))
