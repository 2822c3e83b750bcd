"""The Rust side cannot be compiled in this image (no cargo / rustc), so it is checked the ways that are possible: the raw
bindings are GENERATED from include/ptb200.h (tools/gen_rust_sys.py) and must be up to date; their #[repr(C)] field orders must
equal the ctypes mirror's, which is what the GPU tests actually run through; every exported symbol must be declared; build.rs must
watch every CUDA source."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYS = os.path.join(ROOT, "rust", "ptb200-sys")


def _rust_structs():
    src = open(os.path.join(SYS, "src", "lib.rs")).read()
    out = {}
    for m in re.finditer(r"pub struct (\w+) \{\n(.*?)\n\}", src, flags=re.S):
        out[m.group(1)] = [re.match(r"\s*pub (\w+): (.+),", line).groups() for line in m.group(2).splitlines()]
    return src, out


def test_generated_bindings_are_up_to_date():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_repr_c_field_order_equals_the_ctypes_mirror(rp):
    import ctypes as C
    _, structs = _rust_structs()
    scalar = {C.c_float: "f32", C.c_double: "f64", C.c_uint32: "u32", C.c_uint64: "u64", C.c_int32: "i32"}

    def rust_ty(ct):
        if ct in scalar:
            return scalar[ct]
        if hasattr(ct, "_length_"):
            return f"[{rust_ty(ct._type_)}; {ct._length_}]"
        if hasattr(ct, "contents") or ct.__name__.startswith("LP_"):
            return None                                   # pointer: compared by name only
        return None                                       # nested struct: compared by name only

    pairs = [("ptb_config", rp._abi.Config), ("ptb_counters", rp._abi.Counters)]
    for sfx in ("f32", "f64"):
        T = rp._abi.TYPES[sfx]
        pairs += [(f"ptb_material_{sfx}", T["Material"]), (f"ptb_sphere_{sfx}", T["Sphere"]), (f"ptb_plane_{sfx}", T["Plane"]),
                  (f"ptb_light_{sfx}", T["Light"]), (f"ptb_camera_{sfx}", T["Camera"]), (f"ptb_background_{sfx}", T["Background"]),
                  (f"ptb_scene_{sfx}", T["Scene"]), (f"ptb_sdf_node_{sfx}", T["SdfNode"]), (f"ptb_sdf_{sfx}", T["Sdf"])]
    for name, ct in pairs:
        assert name in structs, name
        rf = structs[name]
        assert [f.rstrip("_") for f, _ in rf] == [n for n, _ in ct._fields_], name
        for (fname, fty), (cname, cty) in zip(rf, ct._fields_):
            want = rust_ty(cty)
            if want is not None:
                assert fty == want, (name, fname, fty, want)
    assert len(structs) == len(pairs)                     # (the opaque ptb_tracer is a one-line declaration, not matched here)

def test_every_exported_symbol_is_declared(rp):
    src, _ = _rust_structs()
    declared = set(re.findall(r"pub fn (ptb_\w+)\(", src))
    assert declared == set(rp._abi.SYMBOLS), declared ^ set(rp._abi.SYMBOLS)


def test_build_rs_watches_every_cuda_source():
    b = open(os.path.join(SYS, "build.rs")).read()
    for f in os.listdir(os.path.join(ROOT, "rust_pathtracer_b200", "csrc")):
        assert f in b, f
    assert "-ftz=true" in b and "arch=compute_100a,code=sm_100a" in b


def test_wrapper_crate_keeps_the_reference_surface():
    """names the reference's prelude exports for this path (lib.rs:24-48; scene.rs:7-25,88; buffer.rs:18-102; tracer.rs:13,22,629)"""
    w = open(os.path.join(ROOT, "rust", "rust-pathtracer", "src", "lib.rs")).read()
    for needle in ("pub trait Scene", "fn background(&self, ray: &Ray) -> F3", "fn closest_hit(&self, ray: &Ray, state: &mut State, light: &mut LightSampleRec) -> bool",
                   "fn any_hit(&self, ray: &Ray, max_dist: F) -> bool", "fn camera(&self) -> &Box<dyn Camera3D>", "fn number_of_lights(&self) -> usize",
                   "fn light_at(&self, index: usize) -> &AnalyticalLight", "fn recursion_depth(&self) -> u16", "fn as_any(&mut self) -> &mut dyn Any",
                   "fn device_export(&self) -> Option<DeviceScene>", "pub struct ColorBuffer", "pub fn to_u8_vec(&self) -> Vec<u8>",
                   "pub fn convert_to_u8(&self, frame: &mut [u8])", "pub fn convert_to_u8_at(&self, frame: &mut [u8], at: (usize, usize, usize, usize))",
                   "pub fn new(scene: Box<dyn Scene>) -> Self", "pub fn render(&mut self, buffer: &mut ColorBuffer)",
                   "pub fn scene(&mut self) -> &mut Box<dyn Scene>", 'feature = "f64"'):
        assert needle in w, needle
