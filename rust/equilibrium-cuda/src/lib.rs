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

/// `exact` reproduces the reference's lexicographic Gauss-Seidel bit for bit; `RedBlack` is the fast path with the
/// stated tolerance.
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum Mode { Exact, RedBlack }

impl CudaFluid {
    fn params(size: u32, delta_t: f32, frames: i64, gs_iterations: i64, diffusion: f32, viscousity: f32, mode: Mode,
              device: i32, rank: i32, world: i32) -> sys::EqParams {
        sys::EqParams {
            size, delta_t, frames, gs_iterations, diffusion, viscosity: viscousity,
            mode: if mode == Mode::Exact { sys::EQ_MODE_EXACT } else { sys::EQ_MODE_RED_BLACK },
            device, rank, world, comm_id: [0u8; 128],
        }
    }
    /// Fluid::new (fluid.rs:93-110): allocates the seven arrays on the device and runs init() once
    /// (velocities 1, the 21x21 density block, the frame walls: fluid.rs:527-570, 602-606).
    pub fn new(size: u32, delta_t: f32, frames: i64, diffusion: f32, viscousity: f32) -> Self {
        Self::with_mode(size, delta_t, frames, diffusion, viscousity, Mode::Exact, 0)
    }
    /// Fluid::new on a chosen device / in a chosen mode (`gs_iterations` = 0 keeps quirk Q1: iterations = frames)
    pub fn with_mode(size: u32, delta_t: f32, frames: i64, diffusion: f32, viscousity: f32, mode: Mode, device: i32) -> Self {
        let p = Self::params(size, delta_t, frames, 0, diffusion, viscousity, mode, device, 0, 1);
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::eq_create(&p, &mut h) });
        CudaFluid { h, size }
    }
    /// init() (fluid.rs:602-606).  `Fluid::default()` calls it a SECOND time on top of `new` (fluid.rs:83-89): see Default.
    pub fn init_default(&mut self) { check(unsafe { sys::eq_init_default(self.h) }); }
    /// The config structs are plain pub fields in the reference (fluid.rs:54-60); the drop-in pushes them before a step.
    /// Legal between steps only.
    pub fn set_params(&mut self, delta_t: f32, frames: i64, gs_iterations: i64, diffusion: f32, viscousity: f32, mode: Mode) {
        let mut p = Self::params(self.size, delta_t, frames, gs_iterations, diffusion, viscousity, mode, 0, 0, 1);
        let mut cur = p;
        check(unsafe { sys::eq_get_params(self.h, &mut cur) });
        p.device = cur.device; p.rank = cur.rank; p.world = cur.world; p.comm_id = cur.comm_id;
        check(unsafe { sys::eq_set_params(self.h, &p) });
    }
    /// add_density (fluid.rs:120-124): density AND scratch_space
    pub fn add_density(&mut self, x: u32, y: u32, amount: f32) { check(unsafe { sys::eq_add_density(self.h, x, y, amount) }); }
    /// "remove obstacle" = a fresh frame-only mask (renderer.rs:145-149 builds a new Fluid for that)
    pub fn reset_walls(&mut self) { check(unsafe { sys::eq_reset_walls(self.h) }); }
    /// `n` frames with scripted point sources, no host round trip per frame (the loop of renderer_helpers.rs:54-60 when
    /// the host computes add_noise's impulses itself)
    pub fn step_n(&mut self, n: i64, sources: &[sys::EqSource]) {
        let p = if sources.is_empty() { std::ptr::null() } else { sources.as_ptr() };
        check(unsafe { sys::eq_step_n(self.h, n, p, sources.len() as i64) });
    }
    /// wait for everything enqueued so far
    pub fn sync(&mut self) { check(unsafe { sys::eq_sync(self.h) }); }
    /// Rectangle::are_all_points_valid (obstacle.rs:74-87) without a device
    pub fn rect_valid(p0: (i64, i64), p1: (i64, i64), size: i64) -> bool {
        unsafe { sys::eq_rect_valid(p0.0, p0.1, p1.0, p1.1, size) == 1 }
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

impl Default for CudaFluid {
    /// `Fluid::default()` (fluid.rs:83-89): `Fluid::new(defaults)` -- which already ran init() (fluid.rs:108) -- followed by
    /// a SECOND init(): velocities 2, density 1.8 in the block.  configs.rs:14-22, 50-60 for the defaults.
    fn default() -> Self {
        let mut f = CudaFluid::new(128, 0.02, 16, 0.0, 0.001);
        f.init_default();
        f
    }
}

/// Row slabs over the GPUs of one box behind ONE object ("multi-GPU handled inside one handle"): the C ABI
/// has one handle per device (each entry point runs on its device's stream), the group owns them, wires their peer
/// mappings once (eq_ipc_export / eq_ipc_attach) and fans every call out from one thread per device -- a slab solver
/// waits for its neighbours inside the kernels, so the per-device calls of one step must be in flight together.
pub struct CudaFluidGroup {
    parts: Vec<CudaFluid>,
}

impl CudaFluidGroup {
    pub fn new(size: u32, delta_t: f32, frames: i64, diffusion: f32, viscousity: f32, mode: Mode, devices: &[i32]) -> Self {
        let world = devices.len() as i32;
        let mut parts = Vec::new();
        for (rank, dev) in devices.iter().enumerate() {
            let p = CudaFluid::params(size, delta_t, frames, 0, diffusion, viscousity, mode, *dev, rank as i32, world);
            let mut h = std::ptr::null_mut();
            check(unsafe { sys::eq_create(&p, &mut h) });
            parts.push(CudaFluid { h, size });
        }
        let bb = unsafe { sys::eq_ipc_blob_bytes() } as usize;
        let mut blobs = vec![0u8; bb * parts.len()];
        for (r, f) in parts.iter().enumerate() {
            check(unsafe { sys::eq_ipc_export(f.h, blobs[r * bb..].as_mut_ptr() as *mut c_void, bb) });
        }
        for f in parts.iter() {
            check(unsafe { sys::eq_ipc_attach(f.h, blobs.as_ptr() as *const c_void, bb, world) });
        }
        CudaFluidGroup { parts }
    }
    fn fan_out<F: Fn(&mut CudaFluid) + Sync>(&mut self, f: F) {
        std::thread::scope(|s| { for p in self.parts.iter_mut() { let f = &f; s.spawn(move || f(p)); } });
    }
    pub fn fill_rect(&mut self, p0: (i64, i64), p1: (i64, i64)) { self.fan_out(|p| p.fill_rect(p0, p1)); }
    pub fn add_velocity(&mut self, x: u32, y: u32, ax: f32, ay: f32) { self.fan_out(|p| p.add_velocity(x, y, ax, ay)); }
    pub fn step(&mut self) { self.fan_out(|p| p.step()); }
    pub fn step_n(&mut self, n: i64, sources: &[sys::EqSource]) { self.fan_out(|p| p.step_n(n, sources)); }
    /// gather a pub field: every rank downloads the rows it owns
    pub fn download_f32(&mut self, field: i32, out: &mut [f32]) {
        let n = self.parts[0].size as usize;
        assert_eq!(out.len(), n * n);
        for p in self.parts.iter_mut() {
            let (mut r0, mut r1) = (0u32, 0u32);
            check(unsafe { sys::eq_owned_rows(p.h, &mut r0, &mut r1) });
            let dst = out[(r0 as usize) * n..].as_mut_ptr() as *mut c_void;
            check(unsafe { sys::eq_download_rows(p.h, field, r0, r1 - r0, dst) });
        }
    }
}
