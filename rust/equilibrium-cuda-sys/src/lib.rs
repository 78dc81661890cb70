//! Raw bindings to `include/equilibrium_cuda.h` (ABI version 1).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct eq_fluid {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct EqParams {
    pub size: u32,
    pub delta_t: f32,
    pub frames: i64,
    pub gs_iterations: i64,
    pub diffusion: f32,
    pub viscosity: f32,
    pub mode: i32,
    pub device: i32,
    pub rank: i32,
    pub world: i32,
    pub comm_id: [u8; 128],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct EqSource {
    pub frame: i64,
    pub x: u32,
    pub y: u32,
    pub d_vx: f32,
    pub d_vy: f32,
    pub d_density: f32,
}

/// Device-side add_noise (fluid.rs:575-599): Philox4x32-10 keyed by `seed`, counter `first_frame + f`.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct EqNoise {
    pub seed: u64,
    pub first_frame: u64,
    pub cos_t: f32,
    pub sin_t: f32,
    pub gain: f32,
    pub reserved: f32,
}

/// Colours of render_image (renderer_helpers.rs:122-143): r, g, b, a bytes.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct EqColors {
    pub world: [u8; 4],
    pub fluid: [u8; 4],
    pub obstacle: [u8; 4],
}

pub const EQ_SNAP_DENSITY: c_int = 0;
pub const EQ_SNAP_RGBA: c_int = 1;
pub const EQ_SNAPSHOT_SLOTS: c_int = 2;

pub const EQ_OK: c_int = 0;
pub const EQ_MODE_EXACT: i32 = 0;
pub const EQ_MODE_RED_BLACK: i32 = 1;
pub const EQ_F_DENSITY: c_int = 0;
pub const EQ_F_VX: c_int = 1;
pub const EQ_F_VY: c_int = 2;
pub const EQ_F_VX0: c_int = 3;
pub const EQ_F_VY0: c_int = 4;
pub const EQ_F_SCRATCH: c_int = 5;
pub const EQ_F_CELLS: c_int = 6;

extern "C" {
    pub fn eq_last_error() -> *const c_char;
    pub fn eq_abi_version() -> c_int;
    pub fn eq_device_count() -> c_int;
    pub fn eq_create(params: *const EqParams, out: *mut *mut eq_fluid) -> c_int;
    pub fn eq_destroy(h: *mut eq_fluid) -> c_int;
    pub fn eq_clone(h: *mut eq_fluid, out: *mut *mut eq_fluid) -> c_int;
    pub fn eq_init_default(h: *mut eq_fluid) -> c_int;
    pub fn eq_add_density(h: *mut eq_fluid, x: u32, y: u32, amount: f32) -> c_int;
    pub fn eq_add_velocity(h: *mut eq_fluid, x: u32, y: u32, ax: f32, ay: f32) -> c_int;
    pub fn eq_rect_valid(x0: i64, y0: i64, x1: i64, y1: i64, size: i64) -> c_int;
    pub fn eq_fill_rect(h: *mut eq_fluid, x0: i64, y0: i64, x1: i64, y1: i64) -> c_int;
    pub fn eq_reset_walls(h: *mut eq_fluid) -> c_int;
    pub fn eq_set_params(h: *mut eq_fluid, params: *const EqParams) -> c_int;
    pub fn eq_get_params(h: *mut eq_fluid, out: *mut EqParams) -> c_int;
    pub fn eq_step(h: *mut eq_fluid) -> c_int;
    pub fn eq_step_n(h: *mut eq_fluid, n: i64, sources: *const EqSource, n_sources: i64) -> c_int;
    pub fn eq_add_noise(h: *mut eq_fluid, noise: *const EqNoise) -> c_int;
    pub fn eq_step_n_noise(h: *mut eq_fluid, n: i64, noise: *const EqNoise) -> c_int;
    pub fn eq_op_add_source(h: *mut eq_fluid, x_field: c_int, s_field: c_int, scale: f32) -> c_int;
    pub fn eq_sync(h: *mut eq_fluid) -> c_int;
    pub fn eq_upload(h: *mut eq_fluid, field: c_int, host: *const c_void, bytes: usize) -> c_int;
    pub fn eq_download(h: *mut eq_fluid, field: c_int, host: *mut c_void, bytes: usize) -> c_int;
    // the frame hand-off (renderer_helpers.rs:61-65, 145-167), see INTEGRATION.md section 4
    pub fn eq_snapshot_begin(h: *mut eq_fluid, kind: c_int, slot: c_int, colors: *const EqColors,
                             host_dst: *mut c_void, bytes: usize) -> c_int;
    pub fn eq_snapshot_wait(h: *mut eq_fluid, slot: c_int) -> c_int;
    pub fn eq_render_rgba(h: *mut eq_fluid, colors: *const EqColors, host_rgba: *mut c_void, bytes: usize) -> c_int;
    pub fn eq_host_alloc(out: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn eq_host_free(p: *mut c_void) -> c_int;
    // row slabs over several GPUs: one handle per device, see CudaFluidGroup
    pub fn eq_upload_rows(h: *mut eq_fluid, field: c_int, row_begin: u32, n_rows: u32, host: *const c_void) -> c_int;
    pub fn eq_download_rows(h: *mut eq_fluid, field: c_int, row_begin: u32, n_rows: u32, host: *mut c_void) -> c_int;
    pub fn eq_owned_rows(h: *mut eq_fluid, row_begin: *mut u32, row_end: *mut u32) -> c_int;
    pub fn eq_ipc_blob_bytes() -> c_int;
    pub fn eq_ipc_export(h: *mut eq_fluid, blob: *mut c_void, capacity: usize) -> c_int;
    pub fn eq_ipc_attach(h: *mut eq_fluid, blobs: *const c_void, blob_bytes: usize, world: c_int) -> c_int;
    pub fn eq_divergence_l2(h: *mut eq_fluid, vx_field: c_int, vy_field: c_int, out: *mut f64) -> c_int;
}
