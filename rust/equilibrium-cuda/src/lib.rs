//! `CudaFluid`: the device-resident twin of `equilibrium::simulation::fluid::Fluid`
//! (src/simulation/fluid.rs:51-110).  The reference crate's `Fluid` delegates to it when
//! built with `--features cuda`; see INTEGRATION.md for the exact patch.
//!
//! NOTE: the build image of this repo has no Rust toolchain; this crate is kept thin and
//! reviewed by eye.  The same C ABI is exercised from Python (equilibrium_b200/fluid.py).
use equilibrium_cuda_sys as sys;
use std::ffi::CStr;
use std::os::raw::c_void;

pub struct CudaFluid {
    h: *mut sys::eq_fluid,
    size: u32,
}

// The reference moves its Fluid to the simulation thread (renderer.rs:125-128).
unsafe impl Send for CudaFluid {}

fn check(code: i32) {
    if code != sys::EQ_OK {
        let msg = unsafe { CStr::from_ptr(sys::eq_last_error()) }.to_string_lossy().into_owned();
        // the reference reports errors by panicking (obstacle.rs:67, renderer_helpers.rs:65)
        panic!("equilibrium_cuda error {}: {}", code, msg);
    }
}

impl CudaFluid {
    /// Fluid::new (fluid.rs:93-110)
    pub fn new(size: u32, delta_t: f32, frames: i64, diffusion: f32, viscousity: f32) -> Self {
        let p = sys::EqParams {
            size, delta_t, frames, gs_iterations: 0, diffusion, viscosity: viscousity,
            mode: sys::EQ_MODE_EXACT, device: 0, rank: 0, world: 1, comm_id: [0u8; 128],
        };
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::eq_create(&p, &mut h) });
        CudaFluid { h, size }
    }
    /// Fluid::step (fluid.rs:437-524)
    pub fn step(&mut self) { check(unsafe { sys::eq_step(self.h) }); }
    /// add_velocity (fluid.rs:127-131) -- what add_noise ends in (fluid.rs:593-598)
    pub fn add_velocity(&mut self, x: u32, y: u32, ax: f32, ay: f32) {
        check(unsafe { sys::eq_add_velocity(self.h, x, y, ax, ay) });
    }
    /// `n` x { add_noise(); step() } (renderer_helpers.rs:54-60) with the impulses drawn on the device: a seeded
    /// Philox stream replaces the reference's unseeded thread_rng (fluid.rs:584-585).  `angle_deg` is what
    /// add_noise passes to rotate_around_point (fluid.rs:578-583, a constant of delta_t).
    pub fn step_n_noise(&mut self, n: i64, seed: u64, first_frame: u64, angle_deg: f32) {
        let th = (angle_deg as f64).to_radians();
        let nz = sys::EqNoise { seed, first_frame, cos_t: th.cos() as f32, sin_t: th.sin() as f32, gain: 2.0, reserved: 0.0 };
        check(unsafe { sys::eq_step_n_noise(self.h, n, &nz) });
    }
    /// fill_obstacle (fluid.rs:610-619) for the two approximate points of a Rectangle
    pub fn fill_rect(&mut self, p0: (i64, i64), p1: (i64, i64)) {
        check(unsafe { sys::eq_fill_rect(self.h, p0.0, p0.1, p1.0, p1.1) });
    }
    /// refresh a host-side pub field (density, velocities_x, velocities_y)
    pub fn download_f32(&mut self, field: i32, out: &mut [f32]) {
        assert_eq!(out.len(), (self.size * self.size) as usize);
        check(unsafe { sys::eq_download(self.h, field, out.as_mut_ptr() as *mut c_void, out.len() * 4) });
    }
    /// cells_type as u8: 0 = NoWall, 1 = DefaultWall
    pub fn download_cells(&mut self, out: &mut [u8]) {
        assert_eq!(out.len(), (self.size * self.size) as usize);
        check(unsafe { sys::eq_download(self.h, sys::EQ_F_CELLS, out.as_mut_ptr() as *mut c_void, out.len()) });
    }
}

/// A frame buffer in pinned host memory (eq_host_alloc) for the snapshot path.
pub struct PinnedFrame {
    ptr: *mut c_void,
    bytes: usize,
}
unsafe impl Send for PinnedFrame {}
impl PinnedFrame {
    pub fn new(bytes: usize) -> Self {
        let mut ptr = std::ptr::null_mut();
        check(unsafe { sys::eq_host_alloc(&mut ptr, bytes) });
        PinnedFrame { ptr, bytes }
    }
    pub fn as_bytes(&self) -> &[u8] { unsafe { std::slice::from_raw_parts(self.ptr as *const u8, self.bytes) } }
    pub fn as_f32(&self) -> &[f32] { unsafe { std::slice::from_raw_parts(self.ptr as *const f32, self.bytes / 4) } }
}
impl Drop for PinnedFrame {
    fn drop(&mut self) { unsafe { sys::eq_host_free(self.ptr); } }
}

impl CudaFluid {
    /// What `fluid.clone()` + `tx.send` is for the render thread (renderer_helpers.rs:61-65): start an asynchronous
    /// snapshot of the current frame (f32 density, or RGBA pixels by render_image's rule, renderer_helpers.rs:145-167)
    /// into `dst`; later steps may run while it is in flight.  `dst` must not be read before `snapshot_wait(slot)`.
    pub fn snapshot_begin(&mut self, rgba: Option<&sys::EqColors>, slot: i32, dst: &mut PinnedFrame) {
        let (kind, colors) = match rgba {
            Some(c) => (sys::EQ_SNAP_RGBA, c as *const sys::EqColors),
            None => (sys::EQ_SNAP_DENSITY, std::ptr::null()),
        };
        check(unsafe { sys::eq_snapshot_begin(self.h, kind, slot, colors, dst.ptr, dst.bytes) });
    }
    pub fn snapshot_wait(&mut self, slot: i32) { check(unsafe { sys::eq_snapshot_wait(self.h, slot) }); }
}

impl Clone for CudaFluid {
    /// #[derive(Clone)] (fluid.rs:51), used once per frame by the caller (renderer_helpers.rs:61-65)
    fn clone(&self) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::eq_clone(self.h, &mut h) });
        CudaFluid { h, size: self.size }
    }
}

impl Drop for CudaFluid {
    fn drop(&mut self) { unsafe { sys::eq_destroy(self.h); } }
}
