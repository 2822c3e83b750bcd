"""A/B harness for build variants of the streaming integrator on BASELINE configs 4 and 5 (scratch tool, not a test, not the bench).

  python tools/ab_stream.py build              # here: nvcc cross-compiles the variants into rust_pathtracer_b200/variants/
  python tools/ab_stream.py run                # on the GPU box: every variant in its own process (PTB200_LIB), times + image check
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "rust_pathtracer_b200", "variants")
sys.path.insert(0, ROOT)

VARIANTS = [
    # knobs (ptb_stream.cuh): PTB_ST_INNER_REPS, PTB_ST_EAGER_FINISH, PTB_ST_REFILL, PTB_ST_LEAF_MIN, PTB_ST_TRACE_MIN_BLOCKS,
    # PTB_ST_TRACE_PLAIN, PTB_ST_SHADE_MIN_BLOCKS, PTB_ST_SPLIT_MIN; results of round 2: profiles/r02_ab_stream.txt
    ("default", []),
    ("box_sentinel", ["-DPTB_ST_BOX_SENTINEL"]),
    ("reps1_eager", ["-DPTB_ST_INNER_REPS=1", "-DPTB_ST_EAGER_FINISH", "-DPTB_ST_REFILL=8", "-DPTB_ST_LEAF_MIN=12", "-DPTB_ST_TRACE_MIN_BLOCKS=1"]),
]


def lib_of(name, flags):
    return os.path.join(ROOT, "rust_pathtracer_b200", "libptb200.so") if not flags else os.path.join(VDIR, f"libptb200_st_{name}.so")


def build():
    import __graft_entry__ as g
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for name, flags in VARIANTS:
        if flags:
            procs.append((name, subprocess.Popen(["nvcc"] + g.NVCC_FLAGS + flags + ["-o", lib_of(name, flags), os.path.join(g.CSRC, "ptb_api.cu")])))
    for name, p in procs:
        assert p.wait() == 0, name
    g.build()


def one():
    import numpy as np
    import rust_pathtracer_b200 as rp
    out = {}
    for cfg, scene, spp, kw in ((4, rp.sphere_field_scene(), 4, {}), (5, rp.divergence_stress_scene(side=64, depth=16), 8, {"rr_start": 3})):
        pt = rp.Tracer.new(scene, integrator=rp._abi.PTB_INTEGRATOR_STREAM, **kw)
        small = rp.ColorBuffer.new(320, 180)
        pt.render_spp(small, 2)
        img = small.pixels.copy()
        buf = rp.ColorBuffer.new(3840, 2160)
        pt.render_spp(buf, 1, download=False)
        best = 1e30
        for _ in range(3):
            pt.render_spp(buf, spp, download=False)
            best = min(best, pt.last_render_ms())
        pt.close()
        ref_path = os.path.join(ROOT, "gpurun_out", f"ab_stream_ref{cfg}.npy")
        if os.path.exists(ref_path):
            ref = np.load(ref_path)
            diff = float(np.abs(img - ref).max())
        else:
            np.save(ref_path, img); diff = None
        out[f"cfg{cfg}"] = {"ms": round(best, 3), "msamples_s": round(3840 * 2160 * spp / best / 1e3, 1), "max_abs_diff_vs_first": diff}
    print(json.dumps(out))


def run():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for c in (4, 5):
        p = os.path.join(ROOT, "gpurun_out", f"ab_stream_ref{c}.npy")
        if os.path.exists(p):
            os.remove(p)
    for name, flags in VARIANTS:
        env = dict(os.environ, PTB200_LIB=lib_of(name, flags))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, capture_output=True, text=True, timeout=600)
        print(f"{name:16s} {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]}", flush=True)


if __name__ == "__main__":
    {"build": build, "run": run, "one": one}[sys.argv[1]]()
