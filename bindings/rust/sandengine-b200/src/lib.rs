//! Rust side of the drop-in boundary: raw FFI for `include/sandengine_b200.h` plus a `Simulation` type with the
//! public surface of `sandengine_core::simulation::Simulation` (sandengine-core/src/simulation.rs:97-253 of the
//! reference): `new`, `run`, `params`, `modifications`, `MODSHAPE_*`.
//! Not compiled in the repository's build image (no Rust toolchain there) -- see INTEGRATION.md.
#![allow(non_camel_case_types, non_snake_case)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};
use std::ptr;

pub mod ffi {
    use super::*;

    #[repr(C)]
    pub struct se_rules { _private: [u8; 0] }
    #[repr(C)]
    pub struct se_sim { _private: [u8; 0] }

    /// == simulation.rs:45-56 `SimModification` (32 bytes, std140 stride of the shader's UBO)
    #[repr(C)]
    #[derive(Clone, Copy, Default, Debug)]
    pub struct se_modification {
        pub position: [i32; 2],
        pub mod_shape: i32,
        pub mod_size: i32,
        pub mod_matID: i32,
        pub _pad4: [i32; 3],
    }

    #[repr(C)]
    #[derive(Clone, Copy, Default, Debug)]
    pub struct se_create_params {
        pub width: u32,
        pub height: u32,
        pub flags: u32,
        pub device: i32,
        pub row_begin: u32,
        pub row_end: u32,
        pub halo_rows: u32,
        pub temporal_block: u32,
        pub device_share: u32,
    }
    pub const SE_FLAG_LIGHTING: u32 = 1;

    extern "C" {
        pub fn se_rules_compile_yaml(yaml: *const c_char, len: usize, out: *mut *mut se_rules) -> c_int;
        pub fn se_rules_destroy(r: *mut se_rules) -> c_int;
        pub fn se_rules_counts(r: *const se_rules, n_rules: *mut i32, n_types: *mut i32, n_materials: *mut i32) -> c_int;
        pub fn se_rules_material(r: *const se_rules, id: i32, name: *mut *const c_char, type_name: *mut *const c_char,
                                 density: *mut f32, color4: *mut f32, emission4: *mut f32, selectable: *mut i32) -> c_int;
        pub fn se_rules_material_id(r: *const se_rules, name: *const c_char, id: *mut i32) -> c_int;
        pub fn se_sim_create(rules: *const se_rules, params: *const se_create_params, out: *mut *mut se_sim) -> c_int;
        pub fn se_sim_destroy(s: *mut se_sim) -> c_int;
        pub fn se_sim_step(s: *mut se_sim, n_steps: u32) -> c_int;
        pub fn se_sim_push_modifications(s: *mut se_sim, mods: *const se_modification, n: u32) -> c_int;
        pub fn se_sim_set_frame(s: *mut se_sim, frame: i32) -> c_int;
        pub fn se_sim_get_frame(s: *const se_sim, frame: *mut i32) -> c_int;
        pub fn se_sim_upload_cells(s: *mut se_sim, host_cells: *const u32) -> c_int;
        pub fn se_sim_download_cells(s: *mut se_sim, host_cells: *mut u32) -> c_int;
        pub fn se_sim_upload_light(s: *mut se_sim, host_rgba: *const f32) -> c_int;
        pub fn se_sim_download_light(s: *mut se_sim, host_rgba: *mut f32) -> c_int;
        pub fn se_sim_download_color(s: *mut se_sim, host_rgba_f32: *mut f32, host_rgba8: *mut u32) -> c_int;
        pub fn se_sim_device_cells(s: *mut se_sim, dptr: *mut *mut c_void, pitch_bytes: *mut usize) -> c_int;
        pub fn se_sim_census(s: *mut se_sim, counts256: *mut u64) -> c_int;
        pub fn se_sim_checksum(s: *mut se_sim, sum: *mut u64) -> c_int;
        pub fn se_sim_synchronize(s: *mut se_sim) -> c_int;
        pub fn se_last_error() -> *const c_char;
    }
}

pub const MODSHAPE_CIRCLE: i32 = 0; // simulation.rs:41
pub const MODSHAPE_SQUARE: i32 = 1; // simulation.rs:42
pub type SimModification = ffi::se_modification;

/// simulation.rs:70-92 (`brushMaterial: SandMaterial` becomes the material id)
#[derive(Clone, Debug)]
pub struct Params {
    pub moveRight: bool,
    pub mousePos: (f32, f32),
    pub mousePressed: bool,
    pub brushSize: u32,
    pub brushMaterialId: i32,
    pub time: f32,
    pub frame: i32,
}
impl Params {
    pub fn new() -> Self {
        Self { moveRight: true, mousePos: (0.0, 0.0), mousePressed: false, brushSize: 5, brushMaterialId: 0, time: 0.0, frame: 0 }
    }
}

#[derive(Debug)]
pub struct Error(pub i32, pub String);

fn check(code: c_int) -> Result<(), Error> {
    if code == 0 {
        Ok(())
    } else {
        let msg = unsafe { CStr::from_ptr(ffi::se_last_error()) }.to_string_lossy().into_owned();
        Err(Error(code, msg))
    }
}

pub struct Simulation {
    rules: *mut ffi::se_rules,
    sim: *mut ffi::se_sim,
    pub size: (u32, u32),
    pub params: Params,
    pub modifications: Vec<SimModification>,
}

impl Simulation {
    /// `Simulation::new` (simulation.rs:128): `yaml` is the text of data/materials.yaml (src/lib.rs:10).
    pub fn new(yaml: &str, size: (u32, u32), lighting: bool) -> Result<Self, Error> {
        let mut rules = ptr::null_mut();
        check(unsafe { ffi::se_rules_compile_yaml(yaml.as_ptr() as *const c_char, yaml.len(), &mut rules) })?;
        let prm = ffi::se_create_params {
            width: size.0, height: size.1, flags: if lighting { ffi::SE_FLAG_LIGHTING } else { 0 }, ..Default::default()
        };
        let mut sim = ptr::null_mut();
        if let Err(e) = check(unsafe { ffi::se_sim_create(rules, &prm, &mut sim) }) {
            unsafe { ffi::se_rules_destroy(rules) };
            return Err(e);
        }
        Ok(Self { rules, sim, size, params: Params::new(), modifications: Vec::new() })
    }

    /// `Simulation::run` (simulation.rs:195-253): one step; `modifications` is consumed (:252).
    pub fn run(&mut self) -> Result<(), Error> {
        check(unsafe { ffi::se_sim_set_frame(self.sim, self.params.frame) })?;
        if !self.modifications.is_empty() {
            check(unsafe { ffi::se_sim_push_modifications(self.sim, self.modifications.as_ptr(), self.modifications.len() as u32) })?;
            self.modifications.clear();
        }
        check(unsafe { ffi::se_sim_step(self.sim, 1) })?;
        check(unsafe { ffi::se_sim_get_frame(self.sim, &mut self.params.frame) })
    }

    /// sandengine-core/src/lib.rs:59-67: the brush stamp pushed while the mouse button is held.
    pub fn push_brush(&mut self) {
        self.modifications.push(SimModification {
            position: [(self.params.mousePos.0 * self.size.0 as f32) as i32, (self.params.mousePos.1 * self.size.1 as f32) as i32],
            mod_shape: MODSHAPE_CIRCLE,
            mod_size: self.params.brushSize as i32,
            mod_matID: self.params.brushMaterialId,
            ..Default::default()
        });
    }

    pub fn upload_cells(&mut self, cells: &[u32]) -> Result<(), Error> {
        assert_eq!(cells.len(), (self.size.0 * self.size.1) as usize);
        check(unsafe { ffi::se_sim_upload_cells(self.sim, cells.as_ptr()) })
    }

    /// Headless read-out of the material ids (the .r channel of the reference's `input_data` texture).
    pub fn cells(&mut self) -> Result<Vec<u32>, Error> {
        let mut v = vec![0u32; (self.size.0 * self.size.1) as usize];
        check(unsafe { ffi::se_sim_download_cells(self.sim, v.as_mut_ptr()) })?;
        Ok(v)
    }

    /// `output_color` as packed RGBA8 (operations.glsl:100-108, shaded on demand).
    pub fn color_rgba8(&mut self) -> Result<Vec<u32>, Error> {
        let mut v = vec![0u32; (self.size.0 * self.size.1) as usize];
        check(unsafe { ffi::se_sim_download_color(self.sim, ptr::null_mut(), v.as_mut_ptr()) })?;
        Ok(v)
    }
}

impl Drop for Simulation {
    fn drop(&mut self) {
        unsafe {
            ffi::se_sim_destroy(self.sim);
            ffi::se_rules_destroy(self.rules);
        }
    }
}
