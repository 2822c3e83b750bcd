"""rust-pathtracer_b200 — B200-native drop-in for rust-pathtracer's per-pixel path-tracing loop.

Only the hot path (`Tracer::render` and what it calls) lives here: hand-written
sm_100a CUDA in `csrc/` behind the C ABI of `include/ptb200.h`, plus this thin host mirror of the
reference crate's prelude.  There is no CPU render path; importing works without a GPU, creating
a `Tracer` does not.
"""
from . import _abi
from .prelude import (F, I, F3, AnalyticalLight, Background, Camera3D, ColorBuffer, DeviceScene, Light, Material, Medium, MediumType, Pinhole, Plane,
                      Scene, SdfNode, SdfProgram, Sphere, Tracer)
from .scenes import AnalyticalScene, ExportedScene, divergence_stress_scene, lights_demo_scene, media_demo_scene, sdf_demo_scene, sphere_field_scene

__all__ = ["F", "I", "F3", "AnalyticalLight", "Background", "Camera3D", "ColorBuffer", "DeviceScene", "Light", "Material", "Medium", "MediumType", "Pinhole",
           "Plane", "Scene", "SdfNode", "SdfProgram", "Sphere", "Tracer", "AnalyticalScene", "ExportedScene", "divergence_stress_scene", "lights_demo_scene", "media_demo_scene", "sdf_demo_scene",
           "sphere_field_scene",
           "_abi"]
