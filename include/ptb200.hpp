// ptb200.hpp — C++17 host mirror of `rust_pathtracer::prelude::*` (rust-pathtracer/src/lib.rs:24-48)
// above the C ABI of ptb200.h.
//
// The reference is compiled code (Rust) and the build image has no Rust toolchain, so this header
// is the compiled-language host side of the drop-in: same type and method names, same argument
// meaning, same behaviour as the crate (`Tracer::new(scene)`, `Tracer::render(&mut buffer)`,
// `ColorBuffer::{new, at, to_u8_vec, convert_to_u8, convert_to_u8_at}`, `Scene`, `Pinhole`,
// `AnalyticalLight::spherical`, `Material::new`, the `F` switch), plus the one API addition the
// device path needs: `Scene::device_export()`.  Header-only; link with libptb200.so.
//
// Errors: the reference's calls are infallible; here a failing C-ABI call (no GPU, CUDA error)
// throws std::runtime_error with ptb_last_error().  There is no CPU fallback.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "ptb200.h"

namespace rust_pathtracer {

// lib.rs:5-6 — the scalar switch (define PTB_F64 for the f64 instantiation)
#ifdef PTB_F64
using F = double;
#define PTB_SFX(name) name##_f64
#else
using F = float;
#define PTB_SFX(name) name##_f32
#endif
using I = int32_t;

// fx.rs:19-205, 209-515 (value types; only what scene descriptions need)
struct F2 { F x = 0, y = 0; F2() {} F2(F x_, F y_) : x(x_), y(y_) {} static F2 new_x(F v) { return F2(v, v); } static F2 zeros() { return F2(); } };
struct F3 {
    F x = 0, y = 0, z = 0;
    F3() {}
    F3(F x_, F y_, F z_) : x(x_), y(y_), z(z_) {}
    static F3 new_x(F v) { return F3(v, v, v); }
    static F3 zeros() { return F3(); }
};

// material.rs:48-114 — defaults are Material::new()'s (rgb 1.5, roughness 0.5, ior 1.45)
// material.rs:5-34 (semantics: PTB_MEDIUM_* in ptb200.h)
struct Medium {
    uint32_t medium_type = PTB_MEDIUM_NONE;
    F density = 0;
    F3 color{};
    F anisotropy = 0;
    static Medium new_() { return Medium(); }
};

struct Material {
    F3 rgb{F(1.5), F(1.5), F(1.5)};
    F3 emission{};
    F anisotropic = 0, metallic = 0, roughness = F(0.5), subsurface = 0, specular_tint = 0;
    F sheen = 0, sheen_tint = 0, clearcoat = 0, clearcoat_gloss = 0, spec_trans = 0, ior = F(1.45);
    // which fields the owning primitive's closest_hit branch assigns (PTB_MAT_*, see ptb200.h)
    uint32_t set_mask = PTB_MAT_ALL;
    uint32_t albedo_kind = PTB_ALBEDO_CONSTANT;
    F checker_a = F(0.25), checker_b = F(0.1), checker_scale = F(0.5), checker_offset = F(100);
    Medium medium;                                                       // material.rs:75
    static Material new_() { return Material(); }
};

// globals.rs:76-84, light.rs:5-28
struct Light { uint32_t light_type = PTB_LIGHT_SPHERICAL; F3 position, emission; F3 u{}, v{}; F radius = 0, area = 0; };
struct AnalyticalLight {
    Light light;
    static AnalyticalLight spherical(F3 position, F radius, F3 emission) {
        AnalyticalLight a;
        a.light.light_type = PTB_LIGHT_SPHERICAL;
        a.light.position = position; a.light.emission = emission; a.light.radius = radius;
        a.light.area = F(4) * F(3.14159265358979323846) * radius * radius;   // light.rs:22
        return a;
    }
    // the two kinds the reference declares but does not implement; they take effect with PTB_SCENE_EXTENDED_LIGHTS (ptb200.h)
    static AnalyticalLight rectangular(F3 position, F3 u, F3 v, F3 emission) {
        AnalyticalLight a;
        a.light.light_type = PTB_LIGHT_RECTANGULAR;
        a.light.position = position; a.light.emission = emission; a.light.u = u; a.light.v = v;
        const F cx = u.y * v.z - u.z * v.y, cy = u.z * v.x - u.x * v.z, cz = u.x * v.y - u.y * v.x;
        a.light.area = std::sqrt(cx * cx + cy * cy + cz * cz);
        return a;
    }
    static AnalyticalLight distant(F3 direction, F3 emission) {
        AnalyticalLight a;
        a.light.light_type = PTB_LIGHT_DISTANT;
        a.light.position = direction; a.light.emission = emission;
        return a;
    }
};

// camera/mod.rs:7-18, camera/pinhole.rs:5-36
struct Camera3D {
    virtual ~Camera3D() {}
    virtual void set(F3 origin, F3 center) = 0;
    virtual void set_fov(F fov) = 0;
};
struct Pinhole : Camera3D {
    F3 origin{0, 0, 3}, center{0, 0, 0};
    F fov = 80;
    static Pinhole new_() { return Pinhole(); }
    void set(F3 o, F3 c) override { origin = o; center = c; }
    void set_fov(F f) override { fov = f; }
};

struct Sphere { F3 center; F radius; uint32_t material; };
struct Plane { F3 point, normal; uint32_t material; };
struct Background { uint32_t kind = PTB_BG_GRADIENT_Y; F3 colour_a{1, 1, 1}, colour_b{F(0.5), F(0.7), F(1.0)}; F scale = F(0.5), gamma = F(2.2); };

/// What Scene::device_export() returns: the scene as data (ptb_scene_f32 / _f64).
struct DeviceScene {
    std::vector<Sphere> spheres;
    std::vector<Plane> planes;
    std::vector<Material> materials;
    std::vector<AnalyticalLight> lights;
    Pinhole camera;
    Background background;
    uint32_t depth = 4;       // Scene::recursion_depth, scene.rs:28-30
    uint32_t flags = 0;
    F eps = F(0.005);         // tracer.rs:16
};

/// scene.rs:5-90.  The per-ray callbacks of the reference (closest_hit, any_hit, background) are host
/// code a GPU cannot call; a scene describes itself once through device_export() instead.
struct Scene {
    virtual ~Scene() {}
    virtual Camera3D& camera() = 0;
    virtual size_t number_of_lights() const = 0;
    virtual const AnalyticalLight& light_at(size_t index) const = 0;
    virtual uint16_t recursion_depth() const { return 4; }
    /// NEW trait method; std::nullopt (the default) makes Tracer's constructor throw: no CPU fallback.
    virtual std::optional<DeviceScene> device_export() const { return std::nullopt; }
};

class Tracer;

/// buffer.rs:6-102 — `pixels` is the running mean, RGBA interleaved, row 0 = top; public fields.
struct ColorBuffer {
    size_t width, height;
    std::vector<F> pixels;
    size_t frames = 0;
    ColorBuffer(size_t w, size_t h) : width(w), height(h), pixels(w * h * 4, F(0)) {}
    static ColorBuffer new_(size_t w, size_t h) { return ColorBuffer(w, h); }
    std::array<F, 4> at(size_t x, size_t y) const {                      // buffer.rs:29-32
        size_t i = y * width * 4 + x * 4;
        return {pixels[i], pixels[i + 1], pixels[i + 2], pixels[i + 3]};
    }
    inline void convert_to_u8(uint8_t* frame) const;                      // buffer.rs:55-64
    inline std::vector<uint8_t> to_u8_vec() const;                        // buffer.rs:37-52
    inline void convert_to_u8_at(uint8_t* frame, size_t x, size_t y, size_t frame_w, size_t frame_h) const;   // buffer.rs:67-102
    const Tracer* tracer_ = nullptr;   // tracer that last rendered into this buffer (device conversions need a handle)
};

inline void ptb_check(int code) {
    if (code != PTB_OK) throw std::runtime_error("ptb200 error " + std::to_string(code) + ": " + ptb_last_error());
}

/// tracer.rs:5-19, 22-123, 629-631
class Tracer {
public:
    explicit Tracer(std::unique_ptr<Scene> scene, const ptb_config* cfg = nullptr) : scene_(std::move(scene)) {
        if (!scene_ || !scene_->device_export())
            throw std::runtime_error("scene does not implement device_export(); the B200 tracer has no CPU fallback");
        ptb_check(ptb_create(cfg, &handle_));
        try { sync_scene(); } catch (...) { ptb_destroy(handle_); handle_ = nullptr; throw; }
    }
    static std::unique_ptr<Tracer> new_(std::unique_ptr<Scene> scene) { return std::make_unique<Tracer>(std::move(scene)); }
    ~Tracer() { if (handle_) ptb_destroy(handle_); }
    Tracer(const Tracer&) = delete;
    Tracer& operator=(const Tracer&) = delete;

    /// tracer.rs:22-123: one more sample per pixel accumulated into buffer.pixels (running mean);
    /// buffer.frames += 1.  `buffer.frames = 0` restarts the accumulation, as in the reference.
    void render(ColorBuffer& buffer) {
        ptb_check(PTB_SFX(ptb_render_frame)(handle_, (uint32_t)buffer.width, (uint32_t)buffer.height, buffer.frames, buffer.pixels.data()));
        buffer.frames += 1;
        buffer.tracer_ = this;
    }
    /// Extension: `spp` samples in one device pass.
    void render_spp(ColorBuffer& buffer, uint32_t spp) {
        ptb_check(ptb_resize_if_needed(buffer));
        if (buffer.frames == 0) ptb_check(ptb_clear(handle_));
        else ptb_check(PTB_SFX(ptb_upload)(handle_, buffer.pixels.data(), buffer.frames));
        ptb_check(ptb_render(handle_, spp, buffer.frames));
        ptb_check(PTB_SFX(ptb_download)(handle_, buffer.pixels.data()));
        buffer.frames += spp;
        buffer.tracer_ = this;
    }
    /// tracer.rs:629-631 — call sync_scene() after editing the scene through this reference.
    Scene& scene() { return *scene_; }
    void sync_scene() {
        DeviceScene e = *scene_->device_export();
        using S = PTB_SFX(ptb_scene); using SP = PTB_SFX(ptb_sphere); using PL = PTB_SFX(ptb_plane);
        using MA = PTB_SFX(ptb_material); using LI = PTB_SFX(ptb_light);
        std::vector<SP> sp(e.spheres.size()); std::vector<PL> pl(e.planes.size()); std::vector<MA> ma(e.materials.size()); std::vector<LI> li(e.lights.size());
        for (size_t i = 0; i < sp.size(); ++i) sp[i] = SP{{e.spheres[i].center.x, e.spheres[i].center.y, e.spheres[i].center.z}, e.spheres[i].radius, e.spheres[i].material};
        for (size_t i = 0; i < pl.size(); ++i) pl[i] = PL{{e.planes[i].point.x, e.planes[i].point.y, e.planes[i].point.z}, {e.planes[i].normal.x, e.planes[i].normal.y, e.planes[i].normal.z}, e.planes[i].material};
        for (size_t i = 0; i < ma.size(); ++i) {
            const Material& m = e.materials[i];
            ma[i] = MA{{m.rgb.x, m.rgb.y, m.rgb.z}, {m.emission.x, m.emission.y, m.emission.z}, m.anisotropic, m.metallic, m.roughness, m.subsurface,
                       m.specular_tint, m.sheen, m.sheen_tint, m.clearcoat, m.clearcoat_gloss, m.spec_trans, m.ior, m.set_mask, m.albedo_kind,
                       m.checker_a, m.checker_b, m.checker_scale, m.checker_offset,
                       m.medium.medium_type, m.medium.density, {m.medium.color.x, m.medium.color.y, m.medium.color.z}, m.medium.anisotropy};
        }
        for (size_t i = 0; i < li.size(); ++i) {
            const Light& l = e.lights[i].light;
            li[i] = LI{{l.position.x, l.position.y, l.position.z}, l.radius, {l.emission.x, l.emission.y, l.emission.z}, l.light_type, {l.u.x, l.u.y, l.u.z}, {l.v.x, l.v.y, l.v.z}};
        }
        S s{};
        s.n_spheres = (uint32_t)sp.size(); s.n_planes = (uint32_t)pl.size(); s.n_materials = (uint32_t)ma.size(); s.n_lights = (uint32_t)li.size();
        s.spheres = sp.data(); s.planes = pl.data(); s.materials = ma.data(); s.lights = li.data();
        s.camera = {{e.camera.origin.x, e.camera.origin.y, e.camera.origin.z}, {e.camera.center.x, e.camera.center.y, e.camera.center.z}, e.camera.fov};
        s.background = {e.background.kind, {e.background.colour_a.x, e.background.colour_a.y, e.background.colour_a.z},
                        {e.background.colour_b.x, e.background.colour_b.y, e.background.colour_b.z}, e.background.scale, e.background.gamma};
        s.depth = e.depth; s.flags = e.flags; s.eps = e.eps;
        ptb_check(PTB_SFX(ptb_set_scene)(handle_, &s));
    }
    ptb_tracer* handle() const { return handle_; }

private:
    int ptb_resize_if_needed(const ColorBuffer& b) {
        if (w_ == b.width && h_ == b.height) return PTB_OK;
        w_ = b.width; h_ = b.height;
        return ptb_resize(handle_, (uint32_t)b.width, (uint32_t)b.height);
    }
    std::unique_ptr<Scene> scene_;
    ptb_tracer* handle_ = nullptr;
    size_t w_ = 0, h_ = 0;
};

inline void ColorBuffer::convert_to_u8(uint8_t* frame) const {
    if (!tracer_) throw std::runtime_error("ColorBuffer conversions run on the device: render into the buffer with a Tracer first");
    ptb_check(PTB_SFX(ptb_convert_pixels_to_u8)(tracer_->handle(), width * height, pixels.data(), frame));
}
inline std::vector<uint8_t> ColorBuffer::to_u8_vec() const {
    std::vector<uint8_t> out(width * height * 4);
    convert_to_u8(out.data());
    return out;
}
inline void ColorBuffer::convert_to_u8_at(uint8_t* frame, size_t x, size_t y, size_t frame_w, size_t frame_h) const {
    if (!tracer_) throw std::runtime_error("ColorBuffer conversions run on the device: render into the buffer with a Tracer first");
    ptb_check(PTB_SFX(ptb_convert_pixels_to_u8_at)(tracer_->handle(), pixels.data(), (uint32_t)width, (uint32_t)height, frame, (uint32_t)x, (uint32_t)y,
                                                   (uint32_t)frame_w, (uint32_t)frame_h));
}

/// renderer/src/analytical.rs:4-159 — the reference's demo scene, exporting itself as data.
struct AnalyticalScene : Scene {
    std::vector<AnalyticalLight> lights;
    Pinhole pinhole;
    AnalyticalScene() { lights.push_back(AnalyticalLight::spherical(F3(3, 2, 2), 1, F3(3, 3, 3))); }   // analytical.rs:15-16
    Camera3D& camera() override { return pinhole; }
    size_t number_of_lights() const override { return lights.size(); }
    const AnalyticalLight& light_at(size_t i) const override { return lights[i]; }
    std::optional<DeviceScene> device_export() const override {
        DeviceScene e;
        e.spheres = {{F3(F(-1.1), 0, 0), 1, 0}, {F3(F(1.1), 0, 0), 1, 1}};            // analytical.rs:41,70
        e.planes = {{F3(0, -1, 0), F3(0, 1, 0), 2}};                                    // analytical.rs:193-198
        Material metal;  metal.set_mask = PTB_MAT_RGB | PTB_MAT_ROUGHNESS | PTB_MAT_METALLIC;           // analytical.rs:56-58
        metal.rgb = F3(1, 1, 1); metal.roughness = F(0.05); metal.metallic = 1;
        Material orange; orange.set_mask = PTB_MAT_RGB | PTB_MAT_CLEARCOAT | PTB_MAT_CLEARCOAT_GLOSS | PTB_MAT_ROUGHNESS;   // :82-85
        orange.rgb = F3(1, F(0.186), 0); orange.clearcoat = 1; orange.clearcoat_gloss = 1; orange.roughness = F(0.1);
        Material floor;  floor.set_mask = PTB_MAT_RGB | PTB_MAT_ROUGHNESS;                                // :107-116
        floor.albedo_kind = PTB_ALBEDO_CHECKER_DIR_RATIO; floor.roughness = 1;
        e.materials = {metal, orange, floor};
        e.lights = lights;
        e.camera = pinhole;
        e.depth = recursion_depth();
        e.flags = PTB_SCENE_ANYHIT_IGNORES_MAX_DIST;                                    // analytical.rs:130
        return e;
    }
};

}  // namespace rust_pathtracer
