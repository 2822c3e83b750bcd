//! The crate API of markusmoenig/rust-pathtracer (`prelude::*`, lib.rs:24-48) with the per-pixel
//! loop running on a B200 through `ptb200-sys`.  Value types (F2/F3/Material/Ray/Pinhole/
//! AnalyticalLight/State/...) are unchanged from the reference crate and elided here; this file
//! shows the three pieces that change: the `Scene` trait's new export method, `Tracer`, and
//! `ColorBuffer`'s conversions.  UNVERIFIED: never compiled (no cargo/rustc in the build image);
//! the executable counterpart of this wrapper is rust_pathtracer_b200/prelude.py, which drives the
//! same C ABI and is what tests/ exercise.
use ptb200_sys as sys;
use std::ffi::CStr;

pub type F = f32; // lib.rs:6 — `f64` selects the *_f64 symbols instead

/// What `Scene::device_export` returns: the scene as data (include/ptb200.h, ptb_scene_f32).
pub struct DeviceScene {
    pub spheres: Vec<sys::ptb_sphere_f32>,
    pub planes: Vec<sys::ptb_plane_f32>,
    pub materials: Vec<sys::ptb_material_f32>,
    pub lights: Vec<sys::ptb_light_f32>,
    pub camera: sys::ptb_camera_f32,
    pub background: sys::ptb_background_f32,
    pub depth: u32,
    pub flags: u32,
    pub eps: F,
}

/// scene.rs:5-90 plus ONE new method.  The per-ray callbacks stay for source compatibility but are
/// never called by this tracer: a GPU cannot call back into host code.
pub trait Scene: Sync + Send {
    fn recursion_depth(&self) -> u16 { 4 }
    /// NEW: describe the scene as data.  `None` (the default) makes `Tracer::new` panic — there is
    /// no CPU fallback.
    fn device_export(&self) -> Option<DeviceScene> { None }
    fn as_any(&mut self) -> &mut dyn std::any::Any;
}

/// buffer.rs:6-26
pub struct ColorBuffer { pub width: usize, pub height: usize, pub pixels: Vec<F>, pub frames: usize }
impl ColorBuffer {
    pub fn new(width: usize, height: usize) -> Self { Self { width, height, pixels: vec![0.0; width * height * 4], frames: 0 } }
    pub fn at(&self, x: usize, y: usize) -> [F; 4] { let i = y * self.width * 4 + x * 4; [self.pixels[i], self.pixels[i + 1], self.pixels[i + 2], self.pixels[i + 3]] }
}

pub struct Tracer { handle: *mut sys::ptb_tracer, scene: Box<dyn Scene> }
unsafe impl Send for Tracer {}

fn check(code: i32) { if code != sys::PTB_OK { let m = unsafe { CStr::from_ptr(sys::ptb_last_error()) }; panic!("ptb200 error {}: {}", code, m.to_string_lossy()); } }

impl Tracer {
    /// tracer.rs:13-19
    pub fn new(scene: Box<dyn Scene>) -> Self {
        let mut handle = std::ptr::null_mut();
        check(unsafe { sys::ptb_create(&sys::ptb_config::default(), &mut handle) });
        let mut t = Self { handle, scene };
        t.sync_scene();
        t
    }
    /// Re-export the scene after editing it through `scene()` (the reference re-reads it per ray).
    pub fn sync_scene(&mut self) {
        let e = self.scene.device_export().expect("scene has no device_export(); the B200 tracer has no CPU fallback");
        let s = sys::ptb_scene_f32 {
            n_spheres: e.spheres.len() as u32, n_planes: e.planes.len() as u32, n_materials: e.materials.len() as u32, n_lights: e.lights.len() as u32,
            spheres: e.spheres.as_ptr(), planes: e.planes.as_ptr(), materials: e.materials.as_ptr(), lights: e.lights.as_ptr(),
            camera: e.camera, background: e.background, depth: e.depth, flags: e.flags, eps: e.eps,
        };
        check(unsafe { sys::ptb_set_scene_f32(self.handle, &s) });
    }
    /// tracer.rs:22-123 — one more sample per pixel, running mean in `buffer.pixels`, `frames += 1`.
    pub fn render(&mut self, buffer: &mut ColorBuffer) {
        check(unsafe { sys::ptb_render_frame_f32(self.handle, buffer.width as u32, buffer.height as u32, buffer.frames as u64, buffer.pixels.as_mut_ptr()) });
        buffer.frames += 1;
    }
    /// Extension: `spp` samples in one device pass (the image stays resident between calls).
    pub fn render_spp(&mut self, buffer: &mut ColorBuffer, spp: u32) {
        unsafe {
            check(sys::ptb_resize(self.handle, buffer.width as u32, buffer.height as u32));
            if buffer.frames > 0 { check(sys::ptb_upload_f32(self.handle, buffer.pixels.as_ptr(), buffer.frames as u64)); }
            check(sys::ptb_render(self.handle, spp, buffer.frames as u64));
            check(sys::ptb_download_f32(self.handle, buffer.pixels.as_mut_ptr()));
        }
        buffer.frames += spp as usize;
    }
    /// buffer.rs:55-64 on the device
    pub fn convert_to_u8(&self, buffer: &ColorBuffer, frame: &mut [u8]) {
        check(unsafe { sys::ptb_convert_pixels_to_u8_f32(self.handle, buffer.width * buffer.height, buffer.pixels.as_ptr(), frame.as_mut_ptr()) });
    }
    /// tracer.rs:629-631
    pub fn scene(&mut self) -> &mut Box<dyn Scene> { &mut self.scene }
}
impl Drop for Tracer { fn drop(&mut self) { unsafe { sys::ptb_destroy(self.handle) } } }
