//! Raw bindings to `include/ptb200.h` (f32 surface; the `_f64` twins are analogous).
//! UNVERIFIED: written against the header, never compiled (no Rust toolchain in the build image).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const PTB_OK: c_int = 0;
pub const PTB_MAT_ALL: u32 = 0x1fff;
pub const PTB_ALBEDO_CONSTANT: u32 = 0;
pub const PTB_ALBEDO_CHECKER_DIR_RATIO: u32 = 1;
pub const PTB_LIGHT_SPHERICAL: u32 = 1;
pub const PTB_BG_CONSTANT: u32 = 0;
pub const PTB_BG_GRADIENT_Y: u32 = 1;
pub const PTB_SCENE_ANYHIT_IGNORES_MAX_DIST: u32 = 1;
pub const PTB_INTEGRATOR_AUTO: u32 = 0;
pub const PTB_INTEGRATOR_FUSED: u32 = 1;
pub const PTB_INTEGRATOR_WAVEFRONT: u32 = 2;
pub const PTB_INTEGRATOR_STREAM: u32 = 3;
pub const PTB_PEER_HANDLE_BYTES: usize = 64;

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct ptb_material_f32 {
    pub rgb: [f32; 3], pub emission: [f32; 3],
    pub anisotropic: f32, pub metallic: f32, pub roughness: f32, pub subsurface: f32, pub specular_tint: f32,
    pub sheen: f32, pub sheen_tint: f32, pub clearcoat: f32, pub clearcoat_gloss: f32, pub spec_trans: f32, pub ior: f32,
    pub set_mask: u32, pub albedo_kind: u32,
    pub checker_a: f32, pub checker_b: f32, pub checker_scale: f32, pub checker_offset: f32,
}
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct ptb_sphere_f32 { pub center: [f32; 3], pub radius: f32, pub material: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct ptb_plane_f32 { pub point: [f32; 3], pub normal: [f32; 3], pub material: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct ptb_light_f32 { pub position: [f32; 3], pub radius: f32, pub emission: [f32; 3], pub type_: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct ptb_camera_f32 { pub origin: [f32; 3], pub center: [f32; 3], pub fov: f32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct ptb_background_f32 { pub kind: u32, pub colour_a: [f32; 3], pub colour_b: [f32; 3], pub scale: f32, pub gamma: f32 }
#[repr(C)]
pub struct ptb_scene_f32 {
    pub n_spheres: u32, pub n_planes: u32, pub n_materials: u32, pub n_lights: u32,
    pub spheres: *const ptb_sphere_f32, pub planes: *const ptb_plane_f32,
    pub materials: *const ptb_material_f32, pub lights: *const ptb_light_f32,
    pub camera: ptb_camera_f32, pub background: ptb_background_f32,
    pub depth: u32, pub flags: u32, pub eps: f32,
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct ptb_config {
    pub device: i32, pub integrator: u32, pub seed: u64, pub rr_start: u32, pub wave_paths: u32,
    pub bvh_threshold: u32, pub collect_counters: u32,
}
#[repr(C)] pub struct ptb_tracer { _private: [u8; 0] }

extern "C" {
    pub fn ptb_create(cfg: *const ptb_config, out: *mut *mut ptb_tracer) -> c_int;
    pub fn ptb_destroy(t: *mut ptb_tracer);
    pub fn ptb_last_error() -> *const c_char;
    pub fn ptb_device_count() -> c_int;
    pub fn ptb_set_stream(t: *mut ptb_tracer, cuda_stream: *mut c_void) -> c_int;
    pub fn ptb_set_scene_f32(t: *mut ptb_tracer, scene: *const ptb_scene_f32) -> c_int;
    pub fn ptb_resize(t: *mut ptb_tracer, width: u32, height: u32) -> c_int;
    pub fn ptb_clear(t: *mut ptb_tracer) -> c_int;
    pub fn ptb_upload_f32(t: *mut ptb_tracer, pixels_rgba: *const f32, frames: u64) -> c_int;
    pub fn ptb_download_f32(t: *mut ptb_tracer, pixels_rgba: *mut f32) -> c_int;
    pub fn ptb_render(t: *mut ptb_tracer, spp: u32, sample_base: u64) -> c_int;
    pub fn ptb_render_frame_f32(t: *mut ptb_tracer, width: u32, height: u32, frames_before: u64, pixels_rgba_inout: *mut f32) -> c_int;
    pub fn ptb_synchronize(t: *mut ptb_tracer) -> c_int;
    pub fn ptb_convert_to_u8(t: *mut ptb_tracer, rgba8: *mut u8) -> c_int;
    pub fn ptb_convert_pixels_to_u8_f32(t: *mut ptb_tracer, n_pixels: usize, rgba: *const f32, rgba8: *mut u8) -> c_int;
    pub fn ptb_convert_pixels_to_u8_at_f32(t: *mut ptb_tracer, rgba: *const f32, width: u32, height: u32, frame_rgba8: *mut u8,
                                           x: u32, y: u32, frame_w: u32, frame_h: u32) -> c_int;
    // multi-GPU gather over peer memory (include/ptb200.h "ptb_peer_*"); handles are PTB_PEER_HANDLE_BYTES = 64 bytes
    pub fn ptb_bind_accumulator(t: *mut ptb_tracer, device_ptr: *mut c_void, width: u32, height: u32) -> c_int;
    pub fn ptb_peer_slots_create(t: *mut ptb_tracer, n_slots: u32, handle_out: *mut u8) -> c_int;
    pub fn ptb_peer_slots_open(t: *mut ptb_tracer, handle: *const u8, n_slots: u32) -> c_int;
    pub fn ptb_peer_set_target(t: *mut ptb_tracer, slot: u32, parity: u32) -> c_int;
    pub fn ptb_peer_sum(t: *mut ptb_tracer, parity: u32) -> c_int;
    pub fn ptb_peer_slots_close(t: *mut ptb_tracer) -> c_int;
}
