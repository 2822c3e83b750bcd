//! The crate API of markusmoenig/rust-pathtracer (`prelude::*`, lib.rs:24-48) with the per-pixel loop — `Tracer::render` and
//! everything it calls — running on a B200 through `ptb200-sys` (include/ptb200.h).
//!
//! What stays: every value type (`F2`, `F3`, `Ray`, `Material`, `State`, `Pinhole`, `AnalyticalLight`, the `math` helpers) is
//! re-exported from the upstream crate unchanged.  What changes is declared here: the `Scene` trait (all of scene.rs:7-25, 88 plus
//! ONE new method, `device_export`), `ColorBuffer` (buffer.rs:6-102, conversions on the device) and `Tracer` (tracer.rs:13-19,
//! 22-123, 629-631).  An existing `impl Scene for MyScene { .. }` compiles unchanged and gains a GPU path by adding
//! `fn device_export(&self)`.
//!
//! UNVERIFIED by a compiler: the build image has no cargo / rustc.  The executable twin of this file is
//! rust_pathtracer_b200/prelude.py, which drives the same C ABI and is what tests/ exercise; tests/test_rust_bindings.py checks
//! the surface below against the reference's signatures.
use ptb200_sys as sys;
use std::any::Any;
use std::ffi::CStr;
use std::sync::{Mutex, OnceLock};

pub use upstream::prelude::{AnalyticalLight, Camera3D, Pinhole, Ray, B3, F2, F3, I};
pub use upstream::{globals::*, material::*, math::*};

/// lib.rs:6 — the scalar switch.  `--features f64` selects the `_f64` entry points of the library.
#[cfg(not(feature = "f64"))]
pub type F = f32;
#[cfg(feature = "f64")]
pub type F = f64;
const _: () = assert!(std::mem::size_of::<F>() == std::mem::size_of::<upstream::F>(), "feature f64 must match the upstream crate's `F` (lib.rs:6)");

// the two instantiations of the C ABI behind one set of names
#[cfg(not(feature = "f64"))]
mod abi {
    pub use ptb200_sys::{
        ptb_background_f32 as Background, ptb_camera_f32 as Camera, ptb_convert_pixels_to_u8_at_f32 as convert_pixels_at,
        ptb_convert_pixels_to_u8_f32 as convert_pixels, ptb_download_f32 as download, ptb_light_f32 as Light, ptb_material_f32 as Material,
        ptb_plane_f32 as Plane, ptb_render_frame_f32 as render_frame, ptb_scene_f32 as Scene, ptb_sdf_f32 as Sdf, ptb_sdf_node_f32 as SdfNode,
        ptb_set_scene_f32 as set_scene, ptb_set_sdf_f32 as set_sdf, ptb_sphere_f32 as Sphere, ptb_upload_f32 as upload,
    };
}
#[cfg(feature = "f64")]
mod abi {
    pub use ptb200_sys::{
        ptb_background_f64 as Background, ptb_camera_f64 as Camera, ptb_convert_pixels_to_u8_at_f64 as convert_pixels_at,
        ptb_convert_pixels_to_u8_f64 as convert_pixels, ptb_download_f64 as download, ptb_light_f64 as Light, ptb_material_f64 as Material,
        ptb_plane_f64 as Plane, ptb_render_frame_f64 as render_frame, ptb_scene_f64 as Scene, ptb_sdf_f64 as Sdf, ptb_sdf_node_f64 as SdfNode,
        ptb_set_scene_f64 as set_scene, ptb_set_sdf_f64 as set_sdf, ptb_sphere_f64 as Sphere, ptb_upload_f64 as upload,
    };
}

pub mod prelude {
    pub use crate::{AnalyticalLight, Camera3D, ColorBuffer, DeviceScene, Pinhole, Ray, Scene, Tracer, B3, F, F2, F3, I};
    pub use upstream::{globals::*, material::*, math::*};
}

fn check(code: i32) {
    if code != sys::PTB_OK {
        let msg = unsafe { CStr::from_ptr(sys::ptb_last_error()) };
        panic!("ptb200 error {}: {}", code, msg.to_string_lossy());      // the reference's calls are infallible: no Result in the API
    }
}

// ------------------------------------------------------------------------------------------------
/// What `Scene::device_export` returns: the scene as data (`ptb_scene_*`), because a GPU cannot call `dyn Scene` methods.
/// `Pinhole` keeps its fields private (camera/pinhole.rs:5-10), so the scene states its camera parameters itself.
pub struct DeviceScene {
    pub spheres: Vec<abi::Sphere>,
    pub planes: Vec<abi::Plane>,
    pub materials: Vec<abi::Material>,
    pub lights: Vec<abi::Light>,
    pub camera: abi::Camera,
    pub background: abi::Background,
    pub depth: u32,
    pub flags: u32,
    pub eps: F,
    /// optional signed-distance body (`ptb_set_sdf_*`): postfix program + (hit_eps, max_dist, normal_h, max_steps)
    pub sdf: Option<(Vec<abi::SdfNode>, F, F, F, u32)>,
}

impl DeviceScene {
    /// `Pinhole::new()` as data: origin (0, 0, 3), centre 0, horizontal fov 80 degrees (camera/pinhole.rs:14-25)
    pub fn pinhole_default() -> abi::Camera { abi::Camera { origin: [0.0, 0.0, 3.0], center: [0.0, 0.0, 0.0], fov: 80.0 } }
    /// a material that assigns every field, starting from `Material::new()`'s defaults (material.rs:82-114)
    pub fn material_default() -> abi::Material {
        abi::Material { rgb: [1.5, 1.5, 1.5], roughness: 0.5, ior: 1.45, set_mask: sys::PTB_MAT_ALL, checker_a: 0.25, checker_b: 0.1,
                        checker_scale: 0.5, checker_offset: 100.0, ..Default::default() }
    }
    /// all fields of a reference `Material` (material.rs:48-78), its `Medium` included (material.rs:15-22; PTB_MEDIUM_* in ptb200.h)
    pub fn material(m: &Material) -> abi::Material {
        abi::Material { rgb: [m.rgb.x, m.rgb.y, m.rgb.z], emission: [m.emission.x, m.emission.y, m.emission.z], anisotropic: m.anisotropic,
                        metallic: m.metallic, roughness: m.roughness, subsurface: m.subsurface, specular_tint: m.specular_tint, sheen: m.sheen,
                        sheen_tint: m.sheen_tint, clearcoat: m.clearcoat, clearcoat_gloss: m.clearcoat_gloss, spec_trans: m.spec_trans, ior: m.ior,
                        medium_type: match m.medium.medium_type { MediumType::None => sys::PTB_MEDIUM_NONE, MediumType::Absorb => sys::PTB_MEDIUM_ABSORB,
                                                                  MediumType::Scatter => sys::PTB_MEDIUM_SCATTER, MediumType::Emissive => sys::PTB_MEDIUM_EMISSIVE },
                        medium_density: m.medium.density, medium_color: [m.medium.color.x, m.medium.color.y, m.medium.color.z],
                        medium_anisotropy: m.medium.anisotropy, ..Self::material_default() }
    }
    /// a spherical light; the library derives `area = 4 pi r^2` in `F` (light.rs:22)
    pub fn light(l: &AnalyticalLight) -> abi::Light {
        let p = l.light.position; let e = l.light.emission;
        let (u, v) = (l.light.u, l.light.v);
        abi::Light { position: [p.x, p.y, p.z], radius: l.light.radius, emission: [e.x, e.y, e.z], u: [u.x, u.y, u.z], v: [v.x, v.y, v.z],
                     type_: match l.light.light_type { LightType::Rectangular => sys::PTB_LIGHT_RECTANGULAR, LightType::Spherical => sys::PTB_LIGHT_SPHERICAL,
                                                        LightType::Distant => sys::PTB_LIGHT_DISTANT } }
    }
}

/// scene.rs:5-90, signature for signature, plus `device_export`.  The per-ray callbacks stay so that existing impls compile,
/// and a host-side caller may still use them; the device tracer never calls them.
#[allow(unused)]
pub trait Scene: Sync + Send {
    fn new() -> Self where Self: Sized;
    /// Background color for the given ray
    fn background(&self, ray: &Ray) -> F3;
    /// Closest hit should return the state.hit_dist, state.normal and fill out the state.material as needed
    fn closest_hit(&self, ray: &Ray, state: &mut State, light: &mut LightSampleRec) -> bool;
    /// Used for shadow rays.
    fn any_hit(&self, ray: &Ray, max_dist: F) -> bool;
    /// Return the camera for the scene
    fn camera(&self) -> &Box<dyn Camera3D>;
    /// Return the number of lights in the scene
    fn number_of_lights(&self) -> usize;
    /// Return a reference for the light at the given index
    fn light_at(&self, index: usize) -> &AnalyticalLight;
    /// The recursion depth for the path tracer (scene.rs:28-30)
    fn recursion_depth(&self) -> u16 { 4 }
    /// scene.rs:32-34
    fn to_linear(&self, c: F3) -> F3 { F3::new(c.x.powf(2.2), c.y.powf(2.2), c.z.powf(2.2)) }
    /// scene.rs:36-86: the nearest spherical light in front of `state.hit_dist` (which is NOT reset between bounces — the stale
    /// value is part of the reference's behaviour, SURVEY.md A.1 — and the device path reproduces it).  Host-side helper for impls
    /// whose `closest_hit` calls it; written here from the reference's semantics: nearest root of the ray-sphere quadratic,
    /// `pdf = d^2 / (area * cos * 0.5)`.
    fn sample_lights(&self, ray: &Ray, state: &mut State, light_sample: &mut LightSampleRec, lights: &Vec<AnalyticalLight>) -> bool {
        let mut found = false;
        let mut nearest = state.hit_dist;
        for l in lights.iter().filter(|l| l.light.light_type == LightType::Spherical) {
            let to_centre = l.light.position - ray.origin;
            let along = to_centre.dot(&ray.direction);
            let off2 = to_centre.dot(&to_centre) - along * along;
            let r2 = l.light.radius * l.light.radius;
            if off2 > r2 { continue; }
            let half = (r2 - off2).sqrt();
            let (near, far) = if along - half > along + half { (along + half, along - half) } else { (along - half, along + half) };
            let d = if near < 0.0 { far } else { near };
            if d < 0.0 || !(d < nearest) { continue; }
            nearest = d;
            let n = normalize(&(ray.at(&d) - l.light.position));
            let cos_theta = dot(&-ray.direction, &n);
            light_sample.pdf = (nearest * nearest) / (l.light.area * cos_theta * 0.5);
            light_sample.emission = l.light.emission;
            state.is_emitter = true;
            state.hit_dist = d;
            found = true;
        }
        found
    }
    /// NEW — describe the scene as data.  `None` (the default) makes `Tracer::new` panic: there is no CPU fallback.
    fn device_export(&self) -> Option<DeviceScene> { None }
    fn as_any(&mut self) -> &mut dyn Any;
}

// ------------------------------------------------------------------------------------------------
// ColorBuffer's conversions run on the device but the struct has no tracer: one library handle per process serves them.
struct Helper(*mut sys::ptb_tracer);
unsafe impl Send for Helper {}
fn helper() -> &'static Mutex<Helper> {
    static H: OnceLock<Mutex<Helper>> = OnceLock::new();
    H.get_or_init(|| {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::ptb_create(&sys::ptb_config::default(), &mut h) });
        Mutex::new(Helper(h))
    })
}

/// buffer.rs:6-26 — same public fields, same methods.
#[derive(PartialEq, Debug, Clone)]
pub struct ColorBuffer {
    pub width: usize,
    pub height: usize,
    pub pixels: Vec<F>,
    pub frames: usize,
}

impl ColorBuffer {
    pub fn new(width: usize, height: usize) -> Self { Self { width, height, pixels: vec![0.0; width * height * 4], frames: 0 } }
    #[inline(always)]
    pub fn at(&self, x: usize, y: usize) -> [F; 4] {
        let i = y * self.width * 4 + x * 4;
        [self.pixels[i], self.pixels[i + 1], self.pixels[i + 2], self.pixels[i + 3]]
    }
    /// buffer.rs:37-52
    pub fn to_u8_vec(&self) -> Vec<u8> {
        let mut out = vec![0u8; self.width * self.height * 4];
        self.convert_to_u8(&mut out);
        out
    }
    /// buffer.rs:55-64: `x^0.4545 * 255` saturating, alpha `* 255`, on the device
    pub fn convert_to_u8(&self, frame: &mut [u8]) {
        assert!(frame.len() >= self.width * self.height * 4);
        let h = helper().lock().unwrap();
        check(unsafe { abi::convert_pixels(h.0, self.width * self.height, self.pixels.as_ptr(), frame.as_mut_ptr()) });
    }
    /// buffer.rs:67-102: copy into a larger frame at `(x, y, frame_width, frame_height)` — no gamma, strict `>` bounds,
    /// bottom-up rows, exactly like the reference
    pub fn convert_to_u8_at(&self, frame: &mut [u8], at: (usize, usize, usize, usize)) {
        assert!(frame.len() >= at.2 * at.3 * 4);
        let h = helper().lock().unwrap();
        check(unsafe { abi::convert_pixels_at(h.0, self.pixels.as_ptr(), self.width as u32, self.height as u32, frame.as_mut_ptr(),
                                              at.0 as u32, at.1 as u32, at.2 as u32, at.3 as u32) });
    }
}

// ------------------------------------------------------------------------------------------------
/// tracer.rs:5-19
pub struct Tracer {
    handle: *mut sys::ptb_tracer,
    scene: Box<dyn Scene>,
    resident: Option<(usize, usize, usize)>,     // (pixels address, len, frames) of the buffer the device image mirrors
}
unsafe impl Send for Tracer {}

impl Tracer {
    /// tracer.rs:13-19
    pub fn new(scene: Box<dyn Scene>) -> Self {
        let mut handle = std::ptr::null_mut();
        check(unsafe { sys::ptb_create(&sys::ptb_config::default(), &mut handle) });
        let mut t = Self { handle, scene, resident: None };
        t.sync_scene();
        t
    }

    /// Re-export the scene after editing it through `scene()` (the reference re-reads the scene on every ray).
    pub fn sync_scene(&mut self) {
        let e = self.scene.device_export().expect("this Scene has no device_export(): the B200 tracer has no CPU fallback");
        let s = abi::Scene {
            n_spheres: e.spheres.len() as u32, n_planes: e.planes.len() as u32, n_materials: e.materials.len() as u32, n_lights: e.lights.len() as u32,
            spheres: e.spheres.as_ptr(), planes: e.planes.as_ptr(), materials: e.materials.as_ptr(), lights: e.lights.as_ptr(),
            camera: e.camera, background: e.background, depth: e.depth, flags: e.flags, eps: e.eps,
        };
        check(unsafe { abi::set_scene(self.handle, &s) });
        if let Some((nodes, hit_eps, max_dist, normal_h, max_steps)) = &e.sdf {
            let sd = abi::Sdf { n_nodes: nodes.len() as u32, nodes: nodes.as_ptr(), hit_eps: *hit_eps, max_dist: *max_dist, normal_h: *normal_h, max_steps: *max_steps };
            check(unsafe { abi::set_sdf(self.handle, &sd) });
        }
        self.resident = None;
    }

    /// tracer.rs:22-123 — one more sample per pixel, running mean in `buffer.pixels`, `buffer.frames += 1`.  `pixels` and `frames`
    /// are public fields the app may edit between calls, so the host buffer is the source of truth and is uploaded first.
    pub fn render(&mut self, buffer: &mut ColorBuffer) {
        check(unsafe { abi::render_frame(self.handle, buffer.width as u32, buffer.height as u32, buffer.frames as u64, buffer.pixels.as_mut_ptr()) });
        buffer.frames += 1;
        self.resident = None;
    }

    /// Extension: `spp` samples in ONE device pass; the image stays resident on the device between calls on the same, untouched
    /// buffer (checked by address, length and frame count), so only the first call uploads.
    pub fn render_spp(&mut self, buffer: &mut ColorBuffer, spp: u32) {
        let key = (buffer.pixels.as_ptr() as usize, buffer.pixels.len(), buffer.frames);
        unsafe {
            if self.resident != Some(key) {
                check(sys::ptb_resize(self.handle, buffer.width as u32, buffer.height as u32));
                if buffer.frames > 0 { check(abi::upload(self.handle, buffer.pixels.as_ptr(), buffer.frames as u64)); }
            }
            check(sys::ptb_render(self.handle, spp, buffer.frames as u64));
            check(abi::download(self.handle, buffer.pixels.as_mut_ptr()));
        }
        buffer.frames += spp as usize;
        self.resident = Some((key.0, key.1, buffer.frames));
    }

    /// `buffer.rs:55-64` straight from the device-resident image (no upload of `pixels`)
    pub fn convert_to_u8(&self, frame: &mut [u8]) { check(unsafe { sys::ptb_convert_to_u8(self.handle, frame.as_mut_ptr()) }); }

    /// tracer.rs:629-631
    pub fn scene(&mut self) -> &mut Box<dyn Scene> { &mut self.scene }
}

impl Drop for Tracer {
    fn drop(&mut self) { unsafe { sys::ptb_destroy(self.handle) } }
}
