"""Host-side mirror of `rust_pathtracer::prelude::*` (rust-pathtracer/src/lib.rs:24-48) above the
C ABI of include/ptb200.h.

Same names, argument meaning and behaviour as the reference crate, with the one API addition the
device path needs: `Scene.device_export()`.  A GPU cannot call back into host
scene code, so `Tracer.new(scene)` FAILS for a scene that does not export itself — there is no CPU
render path in this package.

    scene  = AnalyticalScene.new()
    buffer = ColorBuffer.new(800, 600)
    pt     = Tracer.new(scene)
    pt.render(buffer)                 # one more sample per pixel, running mean in buffer.pixels
    buffer.convert_to_u8(frame)       # frame: bytearray / np.uint8 array of w*h*4
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _abi

# lib.rs:5-6 — the scalar switch.  "f32" (the reference's setting at this commit) or "f64".
F = "f32"
I = np.int32

_NP = {"f32": np.float32, "f64": np.float64}


class F3:
    """fx.rs:209-515 (value type; only what scene descriptions need)."""
    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    @staticmethod
    def new(x, y, z):
        return F3(x, y, z)

    @staticmethod
    def new_x(v):
        return F3(v, v, v)

    @staticmethod
    def zeros():
        return F3(0.0, 0.0, 0.0)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __repr__(self):
        return f"F3({self.x}, {self.y}, {self.z})"


def _f3(v) -> F3:
    return v if isinstance(v, F3) else F3(*v)


class MediumType:
    """material.rs:7-13"""
    NONE, ABSORB, SCATTER, EMISSIVE = _abi.PTB_MEDIUM_NONE, _abi.PTB_MEDIUM_ABSORB, _abi.PTB_MEDIUM_SCATTER, _abi.PTB_MEDIUM_EMISSIVE


@dataclass
class Medium:
    """material.rs:15-34 — the reference's tracer never reads it; semantics here: PTB_MEDIUM_* in include/ptb200.h"""
    medium_type: int = _abi.PTB_MEDIUM_NONE
    density: float = 0.0
    color: F3 = field(default_factory=F3.zeros)
    anisotropy: float = 0.0

    @staticmethod
    def new() -> "Medium":
        return Medium()


@dataclass
class Material:
    """material.rs:48-114 — defaults are Material::new()'s (rgb 1.5!, roughness 0.5, ior 1.45).

    `set_mask` lists which fields the owning primitive's closest_hit branch assigns
    (see PTB_MAT_* in include/ptb200.h); `None` = derive from the fields passed to the constructor
    via `Material.assigning(...)`, PTB_MAT_ALL when built directly.
    """
    rgb: F3 = field(default_factory=lambda: F3(1.5, 1.5, 1.5))
    emission: F3 = field(default_factory=F3.zeros)
    anisotropic: float = 0.0
    metallic: float = 0.0
    roughness: float = 0.5
    subsurface: float = 0.0
    specular_tint: float = 0.0
    sheen: float = 0.0
    sheen_tint: float = 0.0
    clearcoat: float = 0.0
    clearcoat_gloss: float = 0.0
    spec_trans: float = 0.0
    ior: float = 1.45
    set_mask: int = _abi.PTB_MAT_ALL
    albedo_kind: int = _abi.PTB_ALBEDO_CONSTANT
    checker_a: float = 0.25
    checker_b: float = 0.1
    checker_scale: float = 0.5
    checker_offset: float = 100.0
    medium: "Medium" = field(default_factory=lambda: Medium())

    _MASKS = {"rgb": _abi.PTB_MAT_RGB, "emission": _abi.PTB_MAT_EMISSION, "anisotropic": _abi.PTB_MAT_ANISOTROPIC,
              "metallic": _abi.PTB_MAT_METALLIC, "roughness": _abi.PTB_MAT_ROUGHNESS, "subsurface": _abi.PTB_MAT_SUBSURFACE,
              "specular_tint": _abi.PTB_MAT_SPECULAR_TINT, "sheen": _abi.PTB_MAT_SHEEN, "sheen_tint": _abi.PTB_MAT_SHEEN_TINT,
              "clearcoat": _abi.PTB_MAT_CLEARCOAT, "clearcoat_gloss": _abi.PTB_MAT_CLEARCOAT_GLOSS,
              "spec_trans": _abi.PTB_MAT_SPEC_TRANS, "ior": _abi.PTB_MAT_IOR}

    @staticmethod
    def new() -> "Material":
        return Material()

    @staticmethod
    def assigning(**fields) -> "Material":
        """A material that assigns exactly the given fields on top of Material::new(), the way a
        closest_hit branch of the reference does (analytical.rs:56-58, 82-85, 115-116)."""
        mask = 0
        kw = {}
        for k, v in fields.items():
            if k in Material._MASKS:
                mask |= Material._MASKS[k]
                kw[k] = _f3(v) if k in ("rgb", "emission") else float(v)
            else:
                kw[k] = v
        if kw.get("albedo_kind", _abi.PTB_ALBEDO_CONSTANT) != _abi.PTB_ALBEDO_CONSTANT:
            mask |= _abi.PTB_MAT_RGB
        return Material(set_mask=mask, **kw)


@dataclass
class Light:
    """globals.rs:76-84"""
    light_type: int
    position: F3
    emission: F3
    radius: float
    area: float
    u: F3 = field(default_factory=lambda: F3(0.0, 0.0, 0.0))
    v: F3 = field(default_factory=lambda: F3(0.0, 0.0, 0.0))


class AnalyticalLight:
    """light.rs:5-28"""

    def __init__(self, light: Light):
        self.light = light

    @staticmethod
    def spherical(position, radius: float, emission) -> "AnalyticalLight":
        import math
        return AnalyticalLight(Light(_abi.PTB_LIGHT_SPHERICAL, _f3(position), _f3(emission), float(radius),
                                     4.0 * math.pi * radius * radius))

    # The reference has only `spherical` (light.rs:13-28) although Light carries u / v / area for the other two kinds
    # (globals.rs:76-84); they take effect with PTB_SCENE_EXTENDED_LIGHTS (include/ptb200.h).
    @staticmethod
    def rectangular(position, u, v, emission) -> "AnalyticalLight":
        """quad position + s*u + t*v, emitting from the side cross(u, v) points to"""
        uu, vv = tuple(_f3(u)), tuple(_f3(v))
        cx = (uu[1] * vv[2] - uu[2] * vv[1], uu[2] * vv[0] - uu[0] * vv[2], uu[0] * vv[1] - uu[1] * vv[0])
        area = (cx[0] ** 2 + cx[1] ** 2 + cx[2] ** 2) ** 0.5
        return AnalyticalLight(Light(_abi.PTB_LIGHT_RECTANGULAR, _f3(position), _f3(emission), 0.0, area, _f3(u), _f3(v)))

    @staticmethod
    def distant(direction, emission) -> "AnalyticalLight":
        """light from infinitely far away in `direction` (stored as the position, like upstream)"""
        return AnalyticalLight(Light(_abi.PTB_LIGHT_DISTANT, _f3(direction), _f3(emission), 0.0, 0.0))


class Camera3D:
    """camera/mod.rs:7-18"""

    def set(self, origin, center):
        raise NotImplementedError

    def set_fov(self, fov: float):
        raise NotImplementedError


class Pinhole(Camera3D):
    """camera/pinhole.rs:5-36 — origin (0,0,3), centre 0, fov 80 degrees (horizontal)."""

    def __init__(self):
        self.origin = F3(0.0, 0.0, 3.0)
        self.center = F3(0.0, 0.0, 0.0)
        self.fov = 80.0

    @staticmethod
    def new() -> "Pinhole":
        return Pinhole()

    def set(self, origin, center):
        self.origin, self.center = _f3(origin), _f3(center)

    def set_fov(self, fov: float):
        self.fov = float(fov)


@dataclass
class Sphere:
    center: F3
    radius: float
    material: int


@dataclass
class Plane:
    point: F3
    normal: F3
    material: int


@dataclass
class Background:
    kind: int = _abi.PTB_BG_GRADIENT_Y
    colour_a: F3 = field(default_factory=lambda: F3(1.0, 1.0, 1.0))
    colour_b: F3 = field(default_factory=lambda: F3(0.5, 0.7, 1.0))
    scale: float = 0.5
    gamma: float = 2.2


@dataclass
class SdfNode:
    """One instruction of a signed-distance program in postfix order (ptb_sdf_node_*): primitives push (distance, material),
    combinators pop two entries and push one."""
    op: int
    material: int = 0
    p: F3 = field(default_factory=lambda: F3(0.0, 0.0, 0.0))
    a: Sequence[float] = (0.0, 0.0, 0.0, 0.0)

    @staticmethod
    def sphere(center, radius: float, material: int) -> "SdfNode":
        return SdfNode(_abi.PTB_SDF_SPHERE, material, _f3(center), (radius, 0.0, 0.0, 0.0))

    @staticmethod
    def box(center, half_extents, material: int, rounding: float = 0.0) -> "SdfNode":
        h = tuple(_f3(half_extents))
        return SdfNode(_abi.PTB_SDF_BOX, material, _f3(center), (h[0], h[1], h[2], rounding))

    @staticmethod
    def torus(center, ring_radius: float, tube_radius: float, material: int) -> "SdfNode":
        return SdfNode(_abi.PTB_SDF_TORUS, material, _f3(center), (ring_radius, tube_radius, 0.0, 0.0))

    @staticmethod
    def plane(normal, offset: float, material: int) -> "SdfNode":
        n = tuple(_f3(normal))
        return SdfNode(_abi.PTB_SDF_PLANE, material, F3(0.0, 0.0, 0.0), (n[0], n[1], n[2], offset))

    @staticmethod
    def union() -> "SdfNode":
        return SdfNode(_abi.PTB_SDF_UNION)

    @staticmethod
    def smooth_union(k: float) -> "SdfNode":
        return SdfNode(_abi.PTB_SDF_SMOOTH_UNION, 0, F3(0.0, 0.0, 0.0), (k, 0.0, 0.0, 0.0))

    @staticmethod
    def subtract() -> "SdfNode":
        return SdfNode(_abi.PTB_SDF_SUBTRACT)

    @staticmethod
    def intersect() -> "SdfNode":
        return SdfNode(_abi.PTB_SDF_INTERSECT)


@dataclass
class SdfProgram:
    """A signed-distance body for `DeviceScene.sdf` (ptb_sdf_*): sphere-traced after the planes in closest_hit order."""
    nodes: List[SdfNode] = field(default_factory=list)
    hit_eps: float = 1e-4
    max_dist: float = 100.0
    normal_h: float = 1e-3
    max_steps: int = 192

    def to_c(self, precision: str = "f32"):
        T = _abi.TYPES[precision]
        real = _abi.REAL[precision]
        arr = (T["SdfNode"] * max(1, len(self.nodes)))()
        for i, n in enumerate(self.nodes):
            arr[i].op = n.op; arr[i].material = n.material
            arr[i].p = (real * 3)(*tuple(_f3(n.p))); arr[i].a = (real * 4)(*[float(x) for x in n.a])
        sd = T["Sdf"]()
        sd.n_nodes = len(self.nodes); sd.nodes = C.cast(arr, C.POINTER(T["SdfNode"]))
        sd.hit_eps, sd.max_dist, sd.normal_h, sd.max_steps = self.hit_eps, self.max_dist, self.normal_h, self.max_steps
        return sd, arr


@dataclass
class DeviceScene:
    """What `Scene.device_export()` returns: the scene as data (ptb_scene_f32 / _f64)."""
    spheres: List[Sphere] = field(default_factory=list)
    planes: List[Plane] = field(default_factory=list)
    materials: List[Material] = field(default_factory=list)
    lights: List[AnalyticalLight] = field(default_factory=list)
    camera: Pinhole = field(default_factory=Pinhole)
    background: Background = field(default_factory=Background)
    depth: int = 4          # Scene::recursion_depth, scene.rs:28-30
    flags: int = 0
    eps: float = 0.005      # tracer.rs:16
    sdf: Optional[SdfProgram] = None     # signed-distance body (ptb_set_sdf_*), SURVEY.md §8 f2

    def to_c(self, precision: str = "f32"):
        """Build the POD struct; returns (scene_struct, keepalive)."""
        T = _abi.TYPES[precision]
        real = _abi.REAL[precision]

        def v3(v):
            return (real * 3)(*tuple(_f3(v)))

        sph = (T["Sphere"] * max(1, len(self.spheres)))()
        for i, s in enumerate(self.spheres):
            sph[i].center = v3(s.center); sph[i].radius = s.radius; sph[i].material = s.material
        pl = (T["Plane"] * max(1, len(self.planes)))()
        for i, p in enumerate(self.planes):
            pl[i].point = v3(p.point); pl[i].normal = v3(p.normal); pl[i].material = p.material
        mats = (T["Material"] * max(1, len(self.materials)))()
        for i, m in enumerate(self.materials):
            d = mats[i]
            d.rgb = v3(m.rgb); d.emission = v3(m.emission)
            for k in ("anisotropic", "metallic", "roughness", "subsurface", "specular_tint", "sheen", "sheen_tint", "clearcoat",
                      "clearcoat_gloss", "spec_trans", "ior", "checker_a", "checker_b", "checker_scale", "checker_offset"):
                setattr(d, k, getattr(m, k))
            d.set_mask = m.set_mask; d.albedo_kind = m.albedo_kind
            d.medium_type = m.medium.medium_type; d.medium_density = m.medium.density
            d.medium_color = v3(m.medium.color); d.medium_anisotropy = m.medium.anisotropy
        li = (T["Light"] * max(1, len(self.lights)))()
        for i, l in enumerate(self.lights):
            li[i].position = v3(l.light.position); li[i].radius = l.light.radius
            li[i].emission = v3(l.light.emission); li[i].type = l.light.light_type
            li[i].u = v3(l.light.u); li[i].v = v3(l.light.v)
        sc = T["Scene"]()
        sc.n_spheres, sc.n_planes, sc.n_materials, sc.n_lights = len(self.spheres), len(self.planes), len(self.materials), len(self.lights)
        sc.spheres = C.cast(sph, C.POINTER(T["Sphere"])); sc.planes = C.cast(pl, C.POINTER(T["Plane"]))
        sc.materials = C.cast(mats, C.POINTER(T["Material"])); sc.lights = C.cast(li, C.POINTER(T["Light"]))
        sc.camera.origin = v3(self.camera.origin); sc.camera.center = v3(self.camera.center); sc.camera.fov = self.camera.fov
        sc.background.kind = self.background.kind
        sc.background.colour_a = v3(self.background.colour_a); sc.background.colour_b = v3(self.background.colour_b)
        sc.background.scale = self.background.scale; sc.background.gamma = self.background.gamma
        sc.depth, sc.flags, sc.eps = self.depth, self.flags, self.eps
        return sc, (sph, pl, mats, li)


class Scene:
    """scene.rs:5-90 — the plug-in trait.  The per-ray callbacks (`closest_hit`, `any_hit`,
    `background`) are host code the device cannot call; a scene instead describes itself once
    through `device_export()`."""

    @classmethod
    def new(cls):
        return cls()

    def camera(self) -> Camera3D:
        raise NotImplementedError

    def number_of_lights(self) -> int:
        raise NotImplementedError

    def light_at(self, index: int) -> AnalyticalLight:
        raise NotImplementedError

    def recursion_depth(self) -> int:     # scene.rs:28-30
        return 4

    def device_export(self) -> Optional[DeviceScene]:
        """NEW trait method (default None => Tracer.new() fails: no CPU fallback)."""
        return None

    def as_any(self):                     # scene.rs:88
        return self


class ColorBuffer:
    """buffer.rs:6-102 — `pixels` is the running MEAN, RGBA interleaved, row 0 = top.

    `pixels` is a public field in the reference (buffer.rs:9) that the app may edit between render calls, so the tracer treats
    the host copy as the source of truth and uploads it before every `render()`.  Touching `.pixels` here hands out the
    writable array and marks the buffer as *escaped* for good (the array may be edited later through that reference); a buffer
    that never escaped — the reference's own loop only calls `render` and `convert_to_u8`, renderer/src/main.rs:113-122 —
    lets `Tracer.render` skip the redundant upload (PTB_FRAME_HOST_UNCHANGED).  `read_pixels()` returns a read-only view
    without escaping."""

    def __init__(self, width: int, height: int, precision: str = None, storage: np.ndarray = None):
        self.precision = precision or F
        self.width, self.height = int(width), int(height)
        if storage is not None:      # caller-provided (e.g. page-locked) memory for `pixels`
            assert storage.dtype == _NP[self.precision] and storage.size == self.width * self.height * 4 and storage.flags.c_contiguous
            self._pixels = storage.reshape(-1)
            self._pixels[:] = 0
        else:
            self._pixels = np.zeros(self.width * self.height * 4, dtype=_NP[self.precision])
        self._own_storage = storage is None
        self._escaped = False
        self._unpin = None        # weakref.finalize that releases the page-lock taken on first render
        self.frames = 0
        self._tracer = None       # tracer whose device image mirrors (pixels, frames)

    @staticmethod
    def new(width: int, height: int, precision: str = None, storage: np.ndarray = None) -> "ColorBuffer":
        return ColorBuffer(width, height, precision, storage)

    @property
    def pixels(self) -> np.ndarray:
        self._escaped = True
        return self._pixels

    @pixels.setter
    def pixels(self, value: np.ndarray):
        value = np.ascontiguousarray(value, dtype=_NP[self.precision]).reshape(-1)
        assert value.size == self.width * self.height * 4
        if self._unpin is not None:
            self._unpin()
            self._unpin = None
        self._pixels, self._own_storage, self._escaped = value, True, True

    def read_pixels(self) -> np.ndarray:
        """read-only view of the running mean (does not mark the buffer as edited)"""
        v = self._pixels.view()
        v.flags.writeable = False
        return v

    def _pin(self, lib) -> None:
        """Page-lock `pixels` for DMA-speed copies; released when this buffer is collected (the wrapper owns the registration,
        never the library: ptb_pin_host in include/ptb200.h)."""
        if self._unpin is not None or not self._own_storage:
            return
        import weakref
        ptr, nbytes = self._pixels.ctypes.data, self._pixels.nbytes
        if lib.ptb_pin_host(C.c_void_p(ptr), nbytes) == _abi.PTB_OK:
            self._unpin = weakref.finalize(self, lib.ptb_unpin_host, C.c_void_p(ptr))
        else:
            self._unpin = lambda: None      # not an error: copies go through the driver's staging path

    def at(self, x: int, y: int):                                   # buffer.rs:29-32
        i = y * self.width * 4 + x * 4
        return [self._pixels[i], self._pixels[i + 1], self._pixels[i + 2], self._pixels[i + 3]]

    def _bound_tracer(self):
        if self._tracer is None or not self._tracer._alive():
            raise RuntimeError("ColorBuffer conversions run on the device: render into the buffer with a Tracer first "
                               "(there is no CPU path)")
        return self._tracer

    def convert_to_u8(self, frame) -> None:                         # buffer.rs:55-64
        """Gamma-encode `self.pixels` into `frame` (w*h*4 bytes) on the device.  Like the reference
        this converts the HOST pixels (a public field the app may have edited, any alpha): H2D, kernel,
        D2H.  `Tracer.convert_to_u8` is the zero-upload variant for the device-resident image."""
        out = np.frombuffer(frame, dtype=np.uint8) if not isinstance(frame, np.ndarray) else frame
        assert out.size >= self.width * self.height * 4
        t = self._bound_tracer()
        fn = getattr(t._lib, f"ptb_convert_pixels_to_u8_{self.precision}")
        t._check(fn(t._handle(), self.width * self.height, self._pixels.ctypes.data, out.ctypes.data))

    def to_u8_vec(self) -> np.ndarray:                              # buffer.rs:37-52
        out = np.zeros(self.width * self.height * 4, dtype=np.uint8)
        self.convert_to_u8(out)
        return out

    def convert_to_u8_at(self, frame, at: Sequence[int]) -> None:   # buffer.rs:67-102
        out = np.frombuffer(frame, dtype=np.uint8) if not isinstance(frame, np.ndarray) else frame
        t = self._bound_tracer()
        fn = getattr(t._lib, f"ptb_convert_pixels_to_u8_at_{self.precision}")
        t._check(fn(t._handle(), self._pixels.ctypes.data, self.width, self.height, out.ctypes.data, at[0], at[1], at[2], at[3]))


class Tracer:
    """tracer.rs:5-19, 22-123, 629-631."""

    def __init__(self, scene: Scene, device: int = 0, precision: str = None, integrator: int = _abi.PTB_INTEGRATOR_AUTO,
                 seed: int = 0, collect_counters: bool = False, rr_start: int = 0, wave_paths: int = 0, bvh_threshold: int = 0,
                 strict: bool = None):
        self.eps = 0.005
        self._scene = scene
        self.precision = precision or F
        self.strict = _abi.strict_default() if strict is None else bool(strict)
        self._lib = _abi.load(self.strict)       # strict: the IEEE / reference-operation-order build of the same kernels
        export = scene.device_export()
        if export is None:
            raise RuntimeError("scene does not implement device_export(); the B200 tracer has no CPU fallback")
        cfg = _abi.Config(device=device, integrator=integrator, seed=seed, rr_start=rr_start, wave_paths=wave_paths,
                          bvh_threshold=bvh_threshold, collect_counters=1 if collect_counters else 0)
        h = C.c_void_p()
        self._check(self._lib.ptb_create(C.byref(cfg), C.byref(h)))
        self._ptr = h
        self._size = (0, 0)
        self.sync_scene()

    @staticmethod
    def new(scene: Scene, **kw) -> "Tracer":
        return Tracer(scene, **kw)

    def scene(self) -> Scene:                                       # tracer.rs:629-631
        """Mutable access to the scene; call `sync_scene()` after editing it."""
        return self._scene

    def prepare_scene(self):
        """the scene as the POD structs of include/ptb200.h, built once (a Rust or C++ host holds these arrays anyway; building
        100 000 ctypes structs is the slow part of `sync_scene` in this Python mirror)"""
        export = self._scene.device_export()
        sc, keep = export.to_c(self.precision)
        sd = export.sdf.to_c(self.precision) if (export.sdf is not None and export.sdf.nodes) else None
        return (sc, keep, sd, export.eps)

    def sync_scene(self, prepared=None) -> None:
        """Re-export the scene to the device (the reference re-reads the scene every ray).  `prepared`: what prepare_scene()
        returned, to upload the same POD arrays again without rebuilding them."""
        sc, keep, sd, eps = prepared if prepared is not None else self.prepare_scene()
        fn = self._lib.ptb_set_scene_f32 if self.precision == "f32" else self._lib.ptb_set_scene_f64
        self._check(fn(self._handle(), C.byref(sc)))
        self.scene_bytes = C.sizeof(sc) + sum(C.sizeof(k) for k in keep)
        if sd is not None:
            fn = self._lib.ptb_set_sdf_f32 if self.precision == "f32" else self._lib.ptb_set_sdf_f64
            self._check(fn(self._handle(), C.byref(sd[0])))
            self.scene_bytes += C.sizeof(sd[0]) + C.sizeof(sd[1])
        self.eps = eps

    # -- internals ------------------------------------------------------------------------------
    def _check(self, code: int) -> None:
        _abi.check(code, self._lib)

    def _handle(self):
        if self._ptr is None:
            raise RuntimeError("tracer destroyed")
        return self._ptr

    def _alive(self) -> bool:
        return self._ptr is not None

    def _device_frames(self) -> int:
        f = C.c_uint64()
        self._check(self._lib.ptb_frames(self._handle(), C.byref(f)))
        return f.value

    def _ensure_size(self, buffer: ColorBuffer):
        if self._size != (buffer.width, buffer.height):
            self._check(self._lib.ptb_resize(self._handle(), buffer.width, buffer.height))
            self._size = (buffer.width, buffer.height)

    def _upload(self, buffer: ColorBuffer):
        self._ensure_size(buffer)
        fn = self._lib.ptb_upload_f32 if self.precision == "f32" else self._lib.ptb_upload_f64
        self._check(fn(self._handle(), buffer._pixels.ctypes.data, buffer.frames))
        buffer._tracer = self

    def convert_to_u8(self, out: np.ndarray):
        """buffer.rs:55-64 of the DEVICE-RESIDENT image (no upload): kernel + D2H of w*h*4 bytes."""
        self._check(self._lib.ptb_convert_to_u8(self._handle(), out.ctypes.data))

    def convert_to_u8_at(self, out: np.ndarray, at):
        """buffer.rs:67-102 of the device-resident image into the host frame `out` (at = x, y, frame_w, frame_h)."""
        self._check(self._lib.ptb_convert_to_u8_at(self._handle(), out.ctypes.data, at[0], at[1], at[2], at[3]))

    # -- the hot path ---------------------------------------------------------------------------
    def render(self, buffer: ColorBuffer) -> None:
        """tracer.rs:22-123: one more sample per pixel accumulated into buffer.pixels (running
        mean), buffer.frames += 1.  `buffer.frames = 0` restarts the accumulation, as in the
        reference."""
        assert buffer.precision == self.precision
        buffer._pin(self._lib)
        fn = self._lib.ptb_render_frame_ex_f32 if self.precision == "f32" else self._lib.ptb_render_frame_ex_f64
        # a buffer whose array never left the wrapper cannot have been edited: the library then skips the upload if the
        # buffer, the frame count and the device image are still what its previous call left behind
        flags = _abi.PTB_FRAME_HOST_UNCHANGED if (not buffer._escaped and buffer._tracer is self) else 0
        self._check(fn(self._handle(), buffer.width, buffer.height, buffer.frames, buffer._pixels.ctypes.data, flags))
        self._size = (buffer.width, buffer.height)
        buffer.frames += 1
        buffer._tracer = self

    def render_spp(self, buffer: ColorBuffer, spp: int, download=True) -> None:
        """Extension: `spp` samples per pixel in one device pass (the reference needs `spp` calls).
        Equivalent to calling render() spp times up to f32 summation order.  download: True (blocking), False, or "async"
        (the copy overlaps whatever is rendered next; `wait_download()` before reading the pixels)."""
        assert buffer.precision == self.precision
        self._ensure_size(buffer)
        if buffer.frames == 0:
            self._check(self._lib.ptb_clear(self._handle()))
        elif self._device_frames() != buffer.frames or buffer._tracer is not self:
            self._upload(buffer)
        self._check(self._lib.ptb_render(self._handle(), spp, buffer.frames))
        buffer.frames += spp
        buffer._tracer = self
        if download == "async":
            self.download_async(buffer)
        elif download:
            self.download(buffer)

    def download(self, buffer: ColorBuffer) -> None:
        fn = self._lib.ptb_download_f32 if self.precision == "f32" else self._lib.ptb_download_f64
        self._check(fn(self._handle(), buffer._pixels.ctypes.data))

    def denoise(self, iterations: int = 4, sigma_color: float = 0.35) -> np.ndarray:
        """a denoised copy of the device-resident mean image (edge-avoiding a-trous wavelet filter, ptb_denoise_*); returns
        W*H*4 reals, the accumulators stay as they are"""
        out = np.empty(self._size[0] * self._size[1] * 4, dtype=_NP[self.precision])
        fn = getattr(self._lib, f"ptb_denoise_{self.precision}")
        self._check(fn(self._handle(), int(iterations), float(sigma_color), out.ctypes.data))
        return out

    def download_async(self, buffer: ColorBuffer) -> None:
        """ptb_download_async_*: resolve on the render stream, D2H on a side stream; `wait_download()` before reading."""
        buffer._pin(self._lib)
        fn = self._lib.ptb_download_async_f32 if self.precision == "f32" else self._lib.ptb_download_async_f64
        self._check(fn(self._handle(), buffer._pixels.ctypes.data))

    def wait_download(self) -> None:
        self._check(self._lib.ptb_wait_download(self._handle()))

    def synchronize(self) -> None:
        self._check(self._lib.ptb_synchronize(self._handle()))

    # -- instrumentation ------------------------------------------------------------------------
    def counters(self) -> dict:
        c = _abi.Counters()
        self._check(self._lib.ptb_get_counters(self._handle(), C.byref(c)))
        return c.as_dict()

    def reset_counters(self) -> None:
        self._check(self._lib.ptb_reset_counters(self._handle()))

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check(self._lib.ptb_launch_count(self._handle(), C.byref(n)))
        return n.value

    def last_render_ms(self) -> float:
        ms = C.c_float()
        self._check(self._lib.ptb_last_render_ms(self._handle(), C.byref(ms)))
        return ms.value

    def integrator_used(self) -> str:
        """which kernel family the last render ran (AUTO resolved): fused | wavefront | wavefront_rm | stream | stream_split,
        with _bvh / _f64 suffixes"""
        i, b = C.c_uint32(), C.c_uint32()
        self._check(self._lib.ptb_last_integrator(self._handle(), C.byref(i), C.byref(b)))
        name = {_abi.PTB_INTEGRATOR_FUSED: "fused", _abi.PTB_INTEGRATOR_WAVEFRONT: "wavefront", _abi.PTB_INTEGRATOR_STREAM: "stream"}[i.value]
        if b.value & _abi.PTB_KERNEL_RM_TABLE:
            name += "_rm"
        if b.value & _abi.PTB_KERNEL_SPLIT:
            name += "_split"
        if b.value & _abi.PTB_KERNEL_BVH:
            name += "_bvh"
        if b.value & _abi.PTB_KERNEL_F64:
            name += "_f64"
        return name

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self._lib.ptb_set_stream(self._handle(), C.c_void_p(cuda_stream)))

    def bind_accumulator(self, device_ptr: int, width: int, height: int) -> None:
        self._check(self._lib.ptb_bind_accumulator(self._handle(), C.c_void_p(device_ptr), width, height))
        self._size = (width, height)

    def render_samples(self, spp: int, sample_base: int) -> None:
        """Raw ptb_render: add samples [sample_base, sample_base+spp) to the device accumulators."""
        self._check(self._lib.ptb_render(self._handle(), spp, sample_base))

    # -- multi-GPU gather over peer memory (ptb_peer_*, include/ptb200.h) --
    def peer_slots_create(self, n_slots: int) -> bytes:
        buf = C.create_string_buffer(_abi.PTB_PEER_HANDLE_BYTES)
        self._check(self._lib.ptb_peer_slots_create(self._handle(), n_slots, buf))
        return buf.raw

    def peer_slots_open(self, handle: bytes, n_slots: int) -> None:
        self._check(self._lib.ptb_peer_slots_open(self._handle(), C.create_string_buffer(handle, _abi.PTB_PEER_HANDLE_BYTES), n_slots))

    def peer_set_target(self, slot: int, parity: int = 0) -> None:
        self._check(self._lib.ptb_peer_set_target(self._handle(), slot & 0xffffffff, parity))

    def peer_sum(self, parity: int = 0) -> None:
        self._check(self._lib.ptb_peer_sum(self._handle(), parity))

    def peer_slots_close(self) -> None:
        self._check(self._lib.ptb_peer_slots_close(self._handle()))

    def clear(self) -> None:
        self._check(self._lib.ptb_clear(self._handle()))

    def close(self) -> None:
        if getattr(self, "_ptr", None) is not None:
            self._lib.ptb_destroy(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
