"""ctypes mirror of include/ptb200.h and loader of the CUDA library.

The product path has NO CPU fallback: `load()` raises if libptb200.so is missing, and
`ptb_create` fails with PTB_E_NO_DEVICE when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

PTB_ABI_VERSION = 4

# status codes
PTB_OK, PTB_E_INVALID, PTB_E_NO_DEVICE, PTB_E_CUDA, PTB_E_NO_SCENE, PTB_E_PRECISION, PTB_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6

PTB_ALBEDO_CONSTANT, PTB_ALBEDO_CHECKER_DIR_RATIO = 0, 1
PTB_LIGHT_RECTANGULAR, PTB_LIGHT_SPHERICAL, PTB_LIGHT_DISTANT = 0, 1, 2
PTB_BG_CONSTANT, PTB_BG_GRADIENT_Y = 0, 1
PTB_SCENE_ANYHIT_IGNORES_MAX_DIST, PTB_SCENE_FORCE_BVH, PTB_SCENE_NO_BVH = 1, 2, 4
PTB_INTEGRATOR_AUTO, PTB_INTEGRATOR_FUSED, PTB_INTEGRATOR_WAVEFRONT, PTB_INTEGRATOR_STREAM = 0, 1, 2, 3
PTB_PEER_HANDLE_BYTES = 64
PTB_SCENE_EXTENDED_LIGHTS = 1 << 3
PTB_FRAME_HOST_UNCHANGED = 1
PTB_MEDIUM_NONE, PTB_MEDIUM_ABSORB, PTB_MEDIUM_SCATTER, PTB_MEDIUM_EMISSIVE = 0, 1, 2, 3
PTB_SDF_SPHERE, PTB_SDF_BOX, PTB_SDF_TORUS, PTB_SDF_PLANE = 0, 1, 2, 3
PTB_SDF_UNION, PTB_SDF_SMOOTH_UNION, PTB_SDF_SUBTRACT, PTB_SDF_INTERSECT = 16, 17, 18, 19
PTB_SDF_MAX_NODES, PTB_SDF_MAX_STACK = 16, 8
PTB_KERNEL_BVH, PTB_KERNEL_RM_TABLE, PTB_KERNEL_SPLIT, PTB_KERNEL_F64 = 1, 2, 4, 8

PTB_MAT_RGB, PTB_MAT_EMISSION, PTB_MAT_ANISOTROPIC, PTB_MAT_METALLIC = 1 << 0, 1 << 1, 1 << 2, 1 << 3
PTB_MAT_ROUGHNESS, PTB_MAT_SUBSURFACE, PTB_MAT_SPECULAR_TINT, PTB_MAT_SHEEN = 1 << 4, 1 << 5, 1 << 6, 1 << 7
PTB_MAT_SHEEN_TINT, PTB_MAT_CLEARCOAT, PTB_MAT_CLEARCOAT_GLOSS = 1 << 8, 1 << 9, 1 << 10
PTB_MAT_SPEC_TRANS, PTB_MAT_IOR, PTB_MAT_ALL = 1 << 11, 1 << 12, 0x1FFF


def _declare(real):
    class Material(C.Structure):
        _fields_ = [("rgb", real * 3), ("emission", real * 3),
                    ("anisotropic", real), ("metallic", real), ("roughness", real), ("subsurface", real),
                    ("specular_tint", real), ("sheen", real), ("sheen_tint", real), ("clearcoat", real),
                    ("clearcoat_gloss", real), ("spec_trans", real), ("ior", real),
                    ("set_mask", C.c_uint32), ("albedo_kind", C.c_uint32),
                    ("checker_a", real), ("checker_b", real), ("checker_scale", real), ("checker_offset", real),
                    ("medium_type", C.c_uint32), ("medium_density", real), ("medium_color", real * 3), ("medium_anisotropy", real)]

    class Sphere(C.Structure):
        _fields_ = [("center", real * 3), ("radius", real), ("material", C.c_uint32)]

    class Plane(C.Structure):
        _fields_ = [("point", real * 3), ("normal", real * 3), ("material", C.c_uint32)]

    class Light(C.Structure):
        _fields_ = [("position", real * 3), ("radius", real), ("emission", real * 3), ("type", C.c_uint32), ("u", real * 3), ("v", real * 3)]

    class Camera(C.Structure):
        _fields_ = [("origin", real * 3), ("center", real * 3), ("fov", real)]

    class Background(C.Structure):
        _fields_ = [("kind", C.c_uint32), ("colour_a", real * 3), ("colour_b", real * 3), ("scale", real), ("gamma", real)]

    class Scene(C.Structure):
        _fields_ = [("n_spheres", C.c_uint32), ("n_planes", C.c_uint32), ("n_materials", C.c_uint32), ("n_lights", C.c_uint32),
                    ("spheres", C.POINTER(Sphere)), ("planes", C.POINTER(Plane)),
                    ("materials", C.POINTER(Material)), ("lights", C.POINTER(Light)),
                    ("camera", Camera), ("background", Background),
                    ("depth", C.c_uint32), ("flags", C.c_uint32), ("eps", real)]

    class SdfNode(C.Structure):
        _fields_ = [("op", C.c_uint32), ("material", C.c_uint32), ("p", real * 3), ("a", real * 4)]

    class Sdf(C.Structure):
        _fields_ = [("n_nodes", C.c_uint32), ("nodes", C.POINTER(SdfNode)), ("hit_eps", real), ("max_dist", real), ("normal_h", real),
                    ("max_steps", C.c_uint32)]

    return dict(Material=Material, Sphere=Sphere, Plane=Plane, Light=Light, Camera=Camera, Background=Background, Scene=Scene,
                SdfNode=SdfNode, Sdf=Sdf)


TYPES = {"f32": _declare(C.c_float), "f64": _declare(C.c_double)}
REAL = {"f32": C.c_float, "f64": C.c_double}


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("integrator", C.c_uint32), ("seed", C.c_uint64), ("rr_start", C.c_uint32),
                ("wave_paths", C.c_uint32), ("bvh_threshold", C.c_uint32), ("collect_counters", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "samples", "closest_hit", "any_hit", "shade", "nee_contrib", "eval_calls",
        "lobe_diffuse", "lobe_clearcoat", "lobe_reflect", "lobe_refract",
        "end_sky", "end_emitter", "end_pdf", "end_depth", "end_rr",
        "ev_diffuse", "ev_clearcoat", "ev_reflect", "ev_refract", "bvh_nodes", "bvh_leaf_tests")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# every symbol include/ptb200.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = [
    "ptb_create", "ptb_destroy", "ptb_abi_version", "ptb_last_error", "ptb_device_count", "ptb_set_stream",
    "ptb_set_scene_f32", "ptb_set_scene_f64", "ptb_set_sdf_f32", "ptb_set_sdf_f64", "ptb_resize", "ptb_bind_accumulator", "ptb_clear",
    "ptb_upload_f32", "ptb_upload_f64", "ptb_download_f32", "ptb_download_f64", "ptb_denoise_f32", "ptb_denoise_f64", "ptb_frames",
    "ptb_render", "ptb_render_frame_f32", "ptb_render_frame_f64", "ptb_render_frame_ex_f32", "ptb_render_frame_ex_f64", "ptb_synchronize",
    "ptb_download_async_f32", "ptb_download_async_f64", "ptb_wait_download", "ptb_pin_host", "ptb_unpin_host",
    "ptb_peer_slots_create", "ptb_peer_slots_open", "ptb_peer_set_target", "ptb_peer_sum", "ptb_peer_slots_close",
    "ptb_convert_to_u8", "ptb_convert_to_u8_at", "ptb_convert_pixels_to_u8_f32", "ptb_convert_pixels_to_u8_f64",
    "ptb_convert_pixels_to_u8_at_f32", "ptb_convert_pixels_to_u8_at_f64", "ptb_get_counters", "ptb_reset_counters", "ptb_launch_count",
    "ptb_last_render_ms", "ptb_last_integrator",
    "ptb_test_sphere_hit_f32", "ptb_test_plane_hit_f32", "ptb_test_gen_ray_f32", "ptb_test_closest_hit_f32",
    "ptb_test_any_hit_f32", "ptb_test_sdf_eval_f32", "ptb_test_sdf_trace_f32", "ptb_test_background_f32", "ptb_test_sample_light_f32", "ptb_test_finalize_f32",
    "ptb_test_disney_eval_f32", "ptb_test_disney_sample_f32", "ptb_test_rng_f32", "ptb_test_resolved_material_f32", "ptb_test_film_quotients_f32", "ptb_test_bvh_build_f32",
]
PTB_RMAT_FLOATS = 35

LIB_NAME = "libptb200.so"
# The STRICT build of the same sources (-DPTB_IEEE -prec-div=true -prec-sqrt=true -ftz=false -fmad=false): every deliberate
# approximation of the shipped build replaced by the correctly rounded operation in the reference's operation order.  It is
# bit-identical to the reference's f32 arithmetic wherever no libm transcendental is involved (profiles/r02_function_parity.md)
# and exists for parity-critical users and for the tests; `Tracer.new(scene, strict=True)` or PTB200_STRICT=1 selects it.
STRICT_LIB_NAME = "libptb200_strict.so"
_libs = {}


def strict_default() -> bool:
    return os.environ.get("PTB200_STRICT", "0") not in ("", "0")


def lib_path(strict: bool = False) -> str:
    # PTB200_LIB: developer knob for A/B-ing two builds of the CUDA library (tools/ab_variants.py); it must exist
    if os.environ.get("PTB200_LIB") and not strict:
        return os.environ["PTB200_LIB"]
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), STRICT_LIB_NAME if strict else LIB_NAME)


class PtbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ptb200 error {code}: {msg}")
        self.code = code


def load(strict: bool = None):
    """Load libptb200.so / libptb200_strict.so (built in-tree by __graft_entry__.build()). Raises if absent — by design."""
    if strict is None:
        strict = strict_default()
    strict = bool(strict)
    if strict in _libs:
        return _libs[strict]
    path = lib_path(strict)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: the CUDA extension must be built (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback for the render path.")
    lib = C.CDLL(path)
    lib.ptb_last_error.restype = C.c_char_p
    lib.ptb_abi_version.restype = C.c_int
    lib.ptb_device_count.restype = C.c_int
    lib.ptb_destroy.restype = None
    lib.ptb_destroy.argtypes = [C.c_void_p]
    lib.ptb_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.ptb_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.ptb_set_scene_f32.argtypes = [C.c_void_p, C.POINTER(TYPES["f32"]["Scene"])]
    lib.ptb_set_scene_f64.argtypes = [C.c_void_p, C.POINTER(TYPES["f64"]["Scene"])]
    lib.ptb_resize.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    lib.ptb_bind_accumulator.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.ptb_clear.argtypes = [C.c_void_p]
    lib.ptb_upload_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ptb_upload_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ptb_download_f32.argtypes = [C.c_void_p, C.c_void_p]
    lib.ptb_download_f64.argtypes = [C.c_void_p, C.c_void_p]
    lib.ptb_frames.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.ptb_denoise_f32.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_void_p]
    lib.ptb_denoise_f64.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_void_p]
    lib.ptb_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
    lib.ptb_test_film_quotients_f32.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.ptb_test_bvh_build_f32.argtypes = [C.POINTER(TYPES["f32"]["Sphere"]), C.c_uint32] + [C.POINTER(C.c_uint32)] * 3
    lib.ptb_test_resolved_material_f32.argtypes = [C.POINTER(TYPES["f32"]["Scene"]), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    lib.ptb_peer_slots_create.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p]
    lib.ptb_peer_slots_open.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
    lib.ptb_peer_set_target.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    lib.ptb_peer_sum.argtypes = [C.c_void_p, C.c_uint32]
    lib.ptb_peer_slots_close.argtypes = [C.c_void_p]
    lib.ptb_render_frame_f32.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p]
    lib.ptb_render_frame_f64.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p]
    lib.ptb_render_frame_ex_f32.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint32]
    lib.ptb_render_frame_ex_f64.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint32]
    lib.ptb_download_async_f32.argtypes = [C.c_void_p, C.c_void_p]
    lib.ptb_download_async_f64.argtypes = [C.c_void_p, C.c_void_p]
    lib.ptb_wait_download.argtypes = [C.c_void_p]
    lib.ptb_pin_host.argtypes = [C.c_void_p, C.c_size_t]
    lib.ptb_unpin_host.argtypes = [C.c_void_p]
    lib.ptb_synchronize.argtypes = [C.c_void_p]
    lib.ptb_convert_to_u8.argtypes = [C.c_void_p, C.c_void_p]
    lib.ptb_convert_to_u8_at.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.ptb_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    lib.ptb_reset_counters.argtypes = [C.c_void_p]
    lib.ptb_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.ptb_last_render_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.ptb_last_integrator.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    vp = C.c_void_p
    lib.ptb_test_sphere_hit_f32.argtypes = [vp, C.c_size_t] + [vp] * 5
    lib.ptb_test_plane_hit_f32.argtypes = [vp, C.c_size_t] + [vp] * 5
    lib.ptb_test_gen_ray_f32.argtypes = [vp, C.c_size_t, vp, vp, C.c_float, C.c_float, vp, vp]
    lib.ptb_test_closest_hit_f32.argtypes = [vp, C.c_size_t] + [vp] * 10
    lib.ptb_test_any_hit_f32.argtypes = [vp, C.c_size_t] + [vp] * 4
    lib.ptb_test_sdf_eval_f32.argtypes = [vp, C.c_size_t] + [vp] * 3
    lib.ptb_test_sdf_trace_f32.argtypes = [vp, C.c_size_t] + [vp] * 6
    lib.ptb_set_sdf_f32.argtypes = [vp, vp]
    lib.ptb_set_sdf_f64.argtypes = [vp, vp]
    lib.ptb_test_background_f32.argtypes = [vp, C.c_size_t, vp, vp]
    lib.ptb_test_sample_light_f32.argtypes = [vp, C.c_size_t, C.c_uint32] + [vp] * 8
    lib.ptb_test_finalize_f32.argtypes = [vp, C.c_size_t, C.c_uint32] + [vp] * 11
    lib.ptb_test_disney_eval_f32.argtypes = [vp, C.c_size_t, C.c_uint32] + [vp] * 6
    lib.ptb_test_disney_sample_f32.argtypes = [vp, C.c_size_t, C.c_uint32] + [vp] * 11
    lib.ptb_test_rng_f32.argtypes = [vp, C.c_size_t, vp, vp, C.c_uint32, vp]
    for sfx in ("f32", "f64"):
        getattr(lib, f"ptb_convert_pixels_to_u8_{sfx}").argtypes = [vp, C.c_size_t, vp, vp]
        getattr(lib, f"ptb_convert_pixels_to_u8_at_{sfx}").argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp] + [C.c_uint32] * 4
    if lib.ptb_abi_version() != PTB_ABI_VERSION:
        raise RuntimeError("libptb200.so ABI version mismatch")
    _libs[strict] = lib
    return lib


def check(code: int, lib=None):
    if code != PTB_OK:
        msg = (lib or load()).ptb_last_error()
        raise PtbError(code, msg.decode() if msg else "")
