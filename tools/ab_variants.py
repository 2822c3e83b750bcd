"""A/B harness for build variants of libptb200.so (scratch tool, not a test, not the bench).

  python tools/ab_variants.py build            # here: nvcc cross-compiles the variants into rust_pathtracer_b200/variants/
  python tools/ab_variants.py run [WxHxSPP]    # on the GPU box: times every variant in its own process, compares the images
  python tools/ab_variants.py one ...          # (internal) one variant in this process

A variant is (name, extra nvcc -D flags, env).  The library is picked through PTB200_LIB (rust_pathtracer_b200/_abi.py).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "rust_pathtracer_b200", "variants")
sys.path.insert(0, ROOT)

VARIANTS = [
    # name, -D flags, env.  The experiments of round 1 (results: profiles/r01_ab_variants.txt) used, among others:
    #   PTB_WF_THREADS_RM / PTB_WF_POOL_RM / PTB_WF_SCENE_BYTES_RM   launch shape of the resolved-material instantiation
    #   PTB_WF_THREADS / PTB_WF_POOL                                 ... of the generic / BVH instantiations
    #   PTB_CHUNK                                                    pixels reserved per hand-out atomic
    #   PTB_WF_NO_TAIL, PTB_WF_TAIL_LOG2, PTB_NO_FILM_FMA            tail items off / block count, IEEE film quotients
    #   env PTB200_NO_RESOLVED_MATERIALS=1                           generic shade path (no material table)
    #   PTB_FULL_DIV                                                 f32 quotients as div.full (`a / b`) instead of rcp + mul
    #   PTB_MUFU_SINCOS                                              sin / cos on MUFU.SIN / MUFU.COS (3.6e-7 abs) instead of the ~1 ulp minimax kernel
    #   PTB_SAT_DROPS_NAN + PTB_CONTRACT_VIEW_COSINE                  round 1's NaN behaviour (saturate drops NaN, v.z contracted)
    #   PTB_WF_ASYNC                                                 barrier-free per-key rings (ptb_wavefront_async.cuh); PTB_WF_REGEN_DEN: in-place regeneration threshold
    ("default", [], {}),
    ("no_emb", ["-DPTB_NO_EMB"], {}),
]


def lib_of(name, flags):
    if not flags:
        return os.path.join(ROOT, "rust_pathtracer_b200", "libptb200.so")
    return os.path.join(VDIR, f"libptb200_{name}.so")


def build():
    import __graft_entry__ as g
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for name, flags, _ in VARIANTS:
        if not flags:
            continue
        out = lib_of(name, flags)
        cmd = ["nvcc"] + g.NVCC_FLAGS + flags + ["-o", out, os.path.join(g.CSRC, "ptb_api.cu")]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        assert p.wait() == 0, name
    g.build()
    for name, flags, _ in VARIANTS:
        r = subprocess.run(["cuobjdump", "-res-usage", lib_of(name, flags)], capture_output=True, text=True).stdout.splitlines()
        for i, l in enumerate(r):
            if "k_render_wavefrontILb0ELb0ELb1E" in l or ("k_render_wavefrontILb0ELb0ELb0E" in l and name.startswith("generic")):
                print(name, r[i + 1].strip())


def one(W, H, spp, ref_path):
    import numpy as np
    import rust_pathtracer_b200 as rp
    scene = rp.AnalyticalScene.new()
    pt = rp.Tracer.new(scene, integrator=rp._abi.PTB_INTEGRATOR_WAVEFRONT)
    # small image for the comparison
    small = rp.ColorBuffer.new(480, 270)
    pt.render_spp(small, 8)
    img = small.pixels.reshape(-1, 4).copy()
    diff = None
    if os.path.exists(ref_path):
        ref = np.load(ref_path)
        rel = np.abs(img[:, :3] - ref[:, :3]).max(axis=1) / np.maximum(np.abs(ref[:, :3]).max(axis=1), 1e-3)
        diff = {"max_rel": float(rel.max()), "frac_gt_1e-4": float((rel > 1e-4).mean()), "frac_ne": float((rel > 0).mean())}
    else:
        np.save(ref_path, img)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, 4, download=False)
    pt.synchronize()
    times = []
    for _ in range(4):
        pt.render_spp(buf, spp, download=False)
        times.append(pt.last_render_ms())
    best = min(times)
    extra = None
    if os.environ.get("PTB_TIMING_REPORT"):
        nw = int(os.environ["PTB_TIMING_REPORT"])
        pt.reset_counters()
        pt.render_spp(buf, spp, download=False)
        pt.synchronize()
        c = pt.counters()
        s1, srt, s2, b1, b2 = c["lobe_diffuse"], c["lobe_clearcoat"], c["lobe_reflect"], c["ev_diffuse"], c["ev_clearcoat"]
        tot = s1 + srt + s2
        extra = {"s1_share": s1 / tot, "sort_share": srt / tot, "s2_share": s2 / tot, "s1_warp_busy": b1 / (nw * s1), "s2_warp_busy": b2 / (nw * s2),
                 "cta_cycles_avg": tot / 148}
    print(json.dumps({"ms": best, "msamples_s": W * H * spp / best / 1e3, "times": times, "mean": float(img[:, :3].mean()), "diff_vs_first": diff, "timing": extra}), flush=True)
    pt.close()


def run(cfg):
    W, H, spp = (int(x) for x in cfg.split("x"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    ref_path = os.path.join(ROOT, "gpurun_out", "ab_ref.npy")
    if os.path.exists(ref_path):
        os.remove(ref_path)
    rows = []
    for name, flags, env in VARIANTS:
        lib = lib_of(name, flags)
        if not os.path.exists(lib):
            print(name, "missing", flush=True)
            continue
        e = dict(os.environ, PTB200_LIB=lib, **env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(W), str(H), str(spp), ref_path], env=e, capture_output=True, text=True, timeout=300)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
        print(f"{name:20s} {line}", flush=True)
        rows.append((name, line))
    with open(os.path.join(ROOT, "gpurun_out", "ab_variants.txt"), "a") as f:
        f.write(f"# {cfg}\n")
        for name, line in rows:
            f.write(f"{name} {line}\n")


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "run"
    if mode == "build":
        build()
    elif mode == "one":
        one(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
    else:
        run(sys.argv[2] if len(sys.argv) > 2 else "3840x2160x64")
