"""Scenes for the device path.

`AnalyticalScene` is the data form of the reference's only Scene impl
(renderer/src/analytical.rs, constants in SURVEY.md Appendix E).  `sphere_field_scene` and
`divergence_stress_scene` are the synthetic configs 4 and 5 of BASELINE.json,
generated with a seeded splitmix64 so the oracle and the device are fed identical data.
"""
from __future__ import annotations

import math

from . import _abi
from .prelude import (AnalyticalLight, Background, DeviceScene, F3, Material, Pinhole, Plane, Scene, SdfNode, SdfProgram, Sphere)


class AnalyticalScene(Scene):
    """renderer/src/analytical.rs:4-159 as data."""

    def __init__(self):
        em = 3.0
        self.lights = [AnalyticalLight.spherical(F3(3.0, 2.0, 2.0), 1.0, F3(em, em, em))]   # analytical.rs:15-16
        self.pinhole = Pinhole.new()                                                          # analytical.rs:20

    def camera(self):
        return self.pinhole

    def number_of_lights(self):
        return len(self.lights)

    def light_at(self, index):
        return self.lights[index]

    def device_export(self) -> DeviceScene:
        mats = [
            # analytical.rs:56-58 — metal sphere assigns rgb, roughness, metallic
            Material.assigning(rgb=(1.0, 1.0, 1.0), roughness=0.05, metallic=1.0),
            # analytical.rs:82-85 — orange clearcoat sphere assigns rgb, clearcoat, clearcoat_gloss, roughness
            Material.assigning(rgb=(1.0, 0.186, 0.0), clearcoat=1.0, clearcoat_gloss=1.0, roughness=0.1),
            # analytical.rs:107-116 — plane assigns rgb (direction-ratio checker) and roughness
            Material.assigning(albedo_kind=_abi.PTB_ALBEDO_CHECKER_DIR_RATIO, checker_a=0.25, checker_b=0.1, checker_scale=0.5,
                               checker_offset=100.0, roughness=1.0),
        ]
        return DeviceScene(
            spheres=[Sphere(F3(-1.1, 0.0, 0.0), 1.0, 0), Sphere(F3(1.1, 0.0, 0.0), 1.0, 1)],       # analytical.rs:41,70
            planes=[Plane(F3(0.0, -1.0, 0.0), F3(0.0, 1.0, 0.0), 2)],                               # analytical.rs:193-198
            materials=mats,
            lights=list(self.lights),
            camera=self.pinhole,
            background=Background(_abi.PTB_BG_GRADIENT_Y, F3(1.0, 1.0, 1.0), F3(0.5, 0.7, 1.0), 0.5, 2.2),  # analytical.rs:28-32
            depth=self.recursion_depth(),
            flags=_abi.PTB_SCENE_ANYHIT_IGNORES_MAX_DIST,                                           # analytical.rs:130
            eps=0.005,
        )


class _SplitMix64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def u(self) -> float:
        return (self.next() >> 11) * (1.0 / 9007199254740992.0)

    def uniform(self, a, b) -> float:
        return a + (b - a) * self.u()


class ExportedScene(Scene):
    """A Scene that simply carries a prepared DeviceScene."""

    def __init__(self, export: DeviceScene):
        self._export = export

    def camera(self):
        return self._export.camera

    def number_of_lights(self):
        return len(self._export.lights)

    def light_at(self, index):
        return self._export.lights[index]

    def recursion_depth(self):
        return self._export.depth

    def device_export(self):
        return self._export


def _checker_plane_material():
    m = Material()      # full mask: order-independent (required with the BVH)
    m.albedo_kind = _abi.PTB_ALBEDO_CHECKER_DIR_RATIO
    m.roughness = 1.0
    return m


def sphere_field_scene(n_spheres: int = 100_000, n_lights_side: int = 8, seed: int = 0xB200) -> ExportedScene:
    """BASELINE.json config 4: procedural sphere field, mixed Disney materials
    (metal / glass / clearcoat by id % 3), n_lights_side^2 spherical lights, sphere BVH."""
    rng = _SplitMix64(seed)
    spheres, mats = [], [_checker_plane_material()]
    for i in range(n_spheres):
        r = 0.05 * math.exp(rng.u() * math.log(0.6 / 0.05))            # log-uniform [0.05, 0.6]
        x = rng.uniform(-60.0, 60.0)
        z = rng.uniform(-120.0, 0.0)
        y = -1.0 + r + rng.uniform(0.0, 6.0)
        kind = i % 3
        m = Material()
        if kind == 0:      # metal
            m.metallic = 1.0; m.roughness = rng.uniform(0.02, 0.4)
            m.rgb = F3(rng.uniform(0.5, 1.0), rng.uniform(0.5, 1.0), rng.uniform(0.5, 1.0))
        elif kind == 1:    # glass
            m.spec_trans = 1.0; m.ior = 1.45; m.roughness = rng.uniform(0.01, 0.1); m.rgb = F3(1.0, 1.0, 1.0)
        else:              # clearcoat
            m.clearcoat = 1.0; m.clearcoat_gloss = rng.uniform(0.5, 1.0); m.roughness = rng.uniform(0.1, 0.6)
            m.rgb = F3(rng.u(), rng.u(), rng.u())
        mats.append(m)
        spheres.append(Sphere(F3(x, y, z), r, len(mats) - 1))
    lights = []
    for gz in range(n_lights_side):
        for gx in range(n_lights_side):
            px = -52.5 + 105.0 * (gx / max(1, n_lights_side - 1)) if n_lights_side > 1 else 0.0
            pz = -112.5 + 105.0 * (gz / max(1, n_lights_side - 1)) if n_lights_side > 1 else -60.0
            lights.append(AnalyticalLight.spherical(F3(px, 10.0, pz), 0.5, F3(40.0, 40.0, 40.0)))
    cam = Pinhole.new()
    cam.set(F3(0.0, 4.0, 14.0), F3(0.0, 0.0, -40.0))
    cam.set_fov(60.0)
    return ExportedScene(DeviceScene(spheres=spheres, planes=[Plane(F3(0.0, -1.0, 0.0), F3(0.0, 1.0, 0.0), 0)], materials=mats,
                                     lights=lights, camera=cam, depth=4, flags=0, eps=0.005))


def divergence_stress_scene(side: int = 64, depth: int = 16, seed: int = 0xD1CE) -> ExportedScene:
    """BASELINE.json config 5: side*side sphere grid, high roughness, half with transmission."""
    rng = _SplitMix64(seed)
    spheres, mats = [], [_checker_plane_material()]
    for iz in range(side):
        for ix in range(side):
            m = Material()
            m.roughness = rng.uniform(0.6, 1.0)
            m.rgb = F3(rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0))
            if (ix + iz) % 2 == 0:
                m.spec_trans = rng.uniform(0.5, 1.0)
            mats.append(m)
            spheres.append(Sphere(F3((ix - side / 2 + 0.5) * 1.2, -0.5, -iz * 1.2 - 2.0), 0.5, len(mats) - 1))
    lights = [AnalyticalLight.spherical(F3(0.0, 12.0, -side * 0.6), 3.0, F3(20.0, 20.0, 20.0)),
              AnalyticalLight.spherical(F3(-side * 0.5, 8.0, -4.0), 2.0, F3(15.0, 12.0, 9.0))]
    cam = Pinhole.new()
    cam.set(F3(0.0, 5.0, 6.0), F3(0.0, -0.5, -side * 0.5))
    cam.set_fov(70.0)
    return ExportedScene(DeviceScene(spheres=spheres, planes=[Plane(F3(0.0, -1.0, 0.0), F3(0.0, 1.0, 0.0), 0)], materials=mats,
                                     lights=lights, camera=cam, depth=depth, flags=0, eps=0.005))


def sdf_demo_scene(depth: int = 4) -> ExportedScene:
    """The reference's open todo "Implement a SDF based example scene" (Readme.md:18) in device form: a rounded box smoothly
    joined to a torus with a sphere carved out of it (clearcoat paint), a glass ball and a metal ball — all one signed-distance
    program —, standing on the demo scene's checker plane under the demo scene's light and sky."""
    paint = Material(); paint.rgb = F3(0.9, 0.25, 0.1); paint.clearcoat = 1.0; paint.clearcoat_gloss = 0.8; paint.roughness = 0.3
    metal = Material(); metal.rgb = F3(0.95, 0.9, 0.8); metal.metallic = 1.0; metal.roughness = 0.08
    glass = Material(); glass.rgb = F3(1.0, 1.0, 1.0); glass.spec_trans = 1.0; glass.roughness = 0.03; glass.ior = 1.45
    mats = [_checker_plane_material(), paint, metal, glass]
    prog = SdfProgram(nodes=[
        SdfNode.box((-0.6, -0.35, 0.0), (0.55, 0.55, 0.55), 1, rounding=0.1),
        SdfNode.torus((-0.6, 0.35, 0.0), 0.55, 0.16, 1),
        SdfNode.smooth_union(0.25),
        SdfNode.sphere((-0.15, 0.15, 0.55), 0.45, 1),
        SdfNode.subtract(),
        SdfNode.sphere((1.2, -0.45, 0.3), 0.55, 3),
        SdfNode.union(),
        SdfNode.sphere((0.55, -0.7, 1.1), 0.3, 2),
        SdfNode.union(),
    ], hit_eps=1e-4, max_dist=60.0, normal_h=1e-3, max_steps=192)
    cam = Pinhole.new()
    return ExportedScene(DeviceScene(spheres=[], planes=[Plane(F3(0.0, -1.0, 0.0), F3(0.0, 1.0, 0.0), 0)], materials=mats,
                                     lights=[AnalyticalLight.spherical(F3(3.0, 2.0, 2.0), 1.0, F3(3.0, 3.0, 3.0))], camera=cam,
                                     depth=depth, flags=0, eps=0.005, sdf=prog))


def media_demo_scene(depth: int = 12) -> ExportedScene:
    """The reference's open item "Support of mediums / volumetric objects" (Readme.md:13) in device form, on the demo scene's checker
    plane under the demo scene's light and sky: three glass balls filled with an absorbing (Beer-Lambert tint), a scattering
    (Henyey-Greenstein fog) and an emissive medium (PTB_MEDIUM_* in include/ptb200.h), plus the demo scene's metal ball."""
    from .prelude import Medium, MediumType

    def glass(medium):
        m = Material(); m.rgb = F3(1.0, 1.0, 1.0); m.spec_trans = 1.0; m.roughness = 0.02; m.ior = 1.3; m.medium = medium
        return m
    metal = Material(); metal.rgb = F3(1.0, 1.0, 1.0); metal.metallic = 1.0; metal.roughness = 0.05
    mats = [_checker_plane_material(),
            glass(Medium(MediumType.ABSORB, 1.5, F3(0.2, 0.8, 0.3), 0.0)),
            glass(Medium(MediumType.SCATTER, 2.5, F3(0.9, 0.9, 0.95), 0.4)),
            glass(Medium(MediumType.EMISSIVE, 0.4, F3(1.0, 0.5, 0.1), 0.0)),
            metal]
    return ExportedScene(DeviceScene(
        spheres=[Sphere(F3(-2.1, 0.0, 0.0), 1.0, 1), Sphere(F3(0.0, 0.0, 0.0), 1.0, 2), Sphere(F3(2.1, 0.0, 0.0), 1.0, 3),
                 Sphere(F3(0.9, -0.55, 1.5), 0.45, 4)],
        planes=[Plane(F3(0.0, -1.0, 0.0), F3(0.0, 1.0, 0.0), 0)], materials=mats,
        lights=[AnalyticalLight.spherical(F3(3.0, 2.0, 2.0), 1.0, F3(3.0, 3.0, 3.0))], camera=Pinhole.new(),
        background=Background(_abi.PTB_BG_GRADIENT_Y, F3(1.0, 1.0, 1.0), F3(0.5, 0.7, 1.0), 0.5, 2.2),
        depth=depth, flags=0, eps=0.005))


def lights_demo_scene(depth: int = 4) -> ExportedScene:
    """The demo scene lit by the two light kinds the reference enumerates but never implements (globals.rs:69-73, tracer.rs:217):
    a downward-facing quad above the spheres and a faint bluish distant light, next to the reference's spherical light
    (PTB_SCENE_EXTENDED_LIGHTS in include/ptb200.h)."""
    e = AnalyticalScene.new().device_export()
    e.lights = [AnalyticalLight.rectangular(F3(-1.0, 3.0, -0.6), F3(2.0, 0.0, 0.0), F3(0.0, 0.0, 1.2), F3(9.0, 8.0, 7.0)),
                e.lights[0],
                AnalyticalLight.distant(F3(0.4, 1.0, 0.3), F3(0.5, 0.6, 0.9))]
    e.flags |= _abi.PTB_SCENE_EXTENDED_LIGHTS
    e.depth = depth
    return ExportedScene(e)
