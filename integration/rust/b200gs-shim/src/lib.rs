//! b200gs-shim — the hot-path surface of `wgpu-3dgs-viewer` 0.2.0 that LioQing/wgpu-3dgs-viewer-app drives
//! (`src/tab/scene.rs:856-869, 2302-2314` and the setters of `scene.rs:785-835`), re-exposed on top of libb200gs.so.
//!
//! NOT COMPILED in the repository's own image (no rustc / cargo there or on its GPU box): written against
//! `include/b200gs.h` through the generated `b200gs-sys`; the same layer IS compiled and tested in C++
//! (`wgpu-3dgs-viewer-app_b200/host/gs.hpp`, `tests/gs_hpp_check.cpp`) and in Python (`__init__.py`).
//!
//! Shape of the replacement: the reference records preprocess / sort / render into a wgpu command encoder; here every
//! call enqueues CUDA work on the viewer's stream, so the encoder / bind-group / queue arguments of the original
//! signatures are dropped (the app-side edit is mechanical: delete those arguments).  Names, argument meaning and error
//! behaviour follow the crate; every method cites the call site in the app it serves.
use b200gs_sys as sys;
use glam::{Mat4, Quat, UVec2, Vec2, Vec3, Vec4};
use std::collections::HashMap;
use std::ffi::{CStr, CString};
use std::marker::PhantomData;

#[derive(Debug)]
pub enum Error {
    Invalid(String),
    Cuda(String),
    OutOfMemory(String),
    Io(String),
    Format(String),
    /// the frame was delivered, but its (bin, splat) entry list was truncated (`b200gs_set_tile_entry_capacity`)
    Overflow(String),
}
impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "{self:?}") }
}
impl std::error::Error for Error {}

fn check(rc: i32) -> Result<(), Error> {
    if rc == sys::B200GS_OK as i32 { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(sys::b200gs_last_error()) }.to_string_lossy().into_owned();
    Err(match rc as u32 {
        sys::B200GS_ERR_INVALID => Error::Invalid(msg),
        sys::B200GS_ERR_OOM => Error::OutOfMemory(msg),
        sys::B200GS_ERR_IO => Error::Io(msg),
        sys::B200GS_ERR_FORMAT => Error::Format(msg),
        sys::B200GS_ERR_OVERFLOW => Error::Overflow(msg),
        _ => Error::Cuda(msg),
    })
}

/// `gs::Gaussian` (src/app.rs:512, 1066)
pub type Gaussian = sys::b200gs_gaussian;
/// `gs::GaussianEditPod` (src/app.rs:1556-1563)
pub type GaussianEditPod = sys::b200gs_edit_pod;
/// `gs::QueryPod` family (scene.rs:1622, 1633)
pub type QueryPod = sys::b200gs_query_pod;
/// `gs::QueryHitResultPod` (scene.rs:650-657)
pub type QueryHitResultPod = sys::b200gs_hit;

/// `gs::GaussianDisplayMode` (src/tab/transform.rs:129-131)
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(u32)]
pub enum GaussianDisplayMode { Splat = 0, Ellipse = 1, Point = 2 }

/// `gs::GaussianShDegree` (src/tab/transform.rs)
#[derive(Clone, Copy, Debug)]
pub struct GaussianShDegree(pub u32);

/// The eight `GaussianPodWith…Configs` of src/app.rs:250-257 select the packed layout at compile time.
pub trait GaussianPod { const SH: u32; const COV3D: u32; }
macro_rules! pod { ($n:ident, $sh:expr, $cov:expr) => { pub struct $n; impl GaussianPod for $n { const SH: u32 = $sh; const COV3D: u32 = $cov; } }; }
pod!(GaussianPodWithShSingleCov3dSingleConfigs, 0, 0);
pod!(GaussianPodWithShSingleCov3dHalfConfigs, 0, 1);
pod!(GaussianPodWithShHalfCov3dSingleConfigs, 1, 0);
pod!(GaussianPodWithShHalfCov3dHalfConfigs, 1, 1);
pod!(GaussianPodWithShNorm8Cov3dSingleConfigs, 2, 0);
pod!(GaussianPodWithShNorm8Cov3dHalfConfigs, 2, 1);
pod!(GaussianPodWithShNoneCov3dSingleConfigs, 3, 0);
pod!(GaussianPodWithShNoneCov3dHalfConfigs, 3, 1);

/// `gs::CameraTrait` (src/app.rs:1236-1247): glam's look_at_rh / perspective_rh, column-major.
pub trait CameraTrait {
    fn view(&self) -> Mat4;
    fn projection(&self, aspect_ratio: f32) -> Mat4;
}

/// One model of a `MultiModelViewer`: `gs::MultiModelViewerGaussianBuffers` + its bind groups (scene.rs:2111-2139).
pub struct Model { h: *mut sys::b200gs_model, len: u64 }
impl Model {
    /// `gaussians_buffer.len()` — scene.rs:608, 862
    pub fn len(&self) -> usize { self.len as usize }
    /// `gaussians_buffer.update_range(queue, start, &[Gaussian])` — scene.rs:2076-2084 (called every frame while a PLY
    /// streams in: the library copies into a pinned ring and returns, nothing synchronises)
    pub fn update_range(&mut self, start: usize, gaussians: &[Gaussian]) -> Result<(), Error> {
        check(unsafe { sys::b200gs_model_update_range(self.h, start as u64, gaussians.as_ptr(), gaussians.len() as u64) })
    }
    /// `gaussians_edit_buffer.download()` — app.rs:789
    pub fn download_edits(&self) -> Result<Vec<GaussianEditPod>, Error> {
        let mut v = vec![GaussianEditPod { flag: 0, color: [0.0, 1.0, 1.0], contrast: 0.0, exposure: 0.0, gamma: 1.0, alpha: 1.0 }; self.len as usize];
        check(unsafe { sys::b200gs_model_download_edits(self.h, v.as_mut_ptr(), v.len() as u64) })?;
        Ok(v)
    }
    /// `mask_buffer.download()` — app.rs:806 (bit i & 31 of word i >> 5; 1 = shown)
    pub fn download_mask(&self) -> Result<Vec<u32>, Error> {
        let mut v = vec![0u32; (self.len as usize + 31) / 32];
        check(unsafe { sys::b200gs_model_download_mask(self.h, v.as_mut_ptr(), v.len() as u64) })?;
        Ok(v)
    }
}

/// `gs::MultiModelViewer<G, K>` — created at scene.rs:1969-1980.
pub struct MultiModelViewer<G: GaussianPod, K: std::hash::Hash + Eq + Clone = String> {
    h: *mut sys::b200gs_viewer,
    pub models: HashMap<K, Model>,
    size: UVec2,
    _g: PhantomData<G>,
}

impl<G: GaussianPod, K: std::hash::Hash + Eq + Clone + ToString> MultiModelViewer<G, K> {
    /// `MultiModelViewer::new_with(device, format, depth_stencil, size)` — scene.rs:1969-1980.  `cuda_device` replaces
    /// the wgpu device; fails with `Error::Cuda` when no sm_100 GPU is usable (there is no CPU fallback).
    pub fn new_with(cuda_device: i32, size: UVec2) -> Result<Self, Error> {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::b200gs_viewer_create(cuda_device, G::SH, G::COV3D, size.x, size.y, &mut h) })?;
        Ok(Self { h, models: HashMap::new(), size, _g: PhantomData })
    }
    /// `MultiModelViewerGaussianBuffers::new_empty(device, count)` + `viewer.models.insert(key, ..)` — scene.rs:2111-2139
    pub fn insert_model(&mut self, key: K, count: usize) -> Result<&mut Model, Error> {
        let name = CString::new(key.to_string()).map_err(|e| Error::Invalid(e.to_string()))?;
        let mut m = std::ptr::null_mut();
        check(unsafe { sys::b200gs_model_create(self.h, name.as_ptr(), count as u64, &mut m) })?;
        Ok(self.models.entry(key).or_insert(Model { h: m, len: count as u64 }))
    }
    /// `viewer.remove_model(key)` — scene.rs:2176
    pub fn remove_model(&mut self, key: &K) -> Result<(), Error> {
        match self.models.remove(key) {
            Some(m) => check(unsafe { sys::b200gs_model_destroy(self.h, m.h) }),
            None => Ok(()),
        }
    }
    /// `viewer.update_camera(queue, &camera, size)` — scene.rs:795
    pub fn update_camera(&mut self, camera: &impl CameraTrait, size: UVec2) -> Result<(), Error> {
        let (v, p) = (camera.view().to_cols_array(), camera.projection(size.x as f32 / size.y as f32).to_cols_array());
        let s = [size.x as f32, size.y as f32];
        self.size = size;
        check(unsafe { sys::b200gs_set_camera(self.h, v.as_ptr(), p.as_ptr(), s.as_ptr()) })
    }
    /// `viewer.update_query_texture_size(device, size)` — scene.rs:740 (also the render-target size)
    pub fn update_query_texture_size(&mut self, size: UVec2) -> Result<(), Error> {
        self.size = size;
        check(unsafe { sys::b200gs_resize(self.h, size.x, size.y) })
    }
    /// `viewer.update_model_transform(queue, key, pos, quat, scale)` — scene.rs:796-802
    pub fn update_model_transform(&mut self, key: &K, pos: Vec3, quat: Quat, scale: Vec3) -> Result<(), Error> {
        let m = self.models.get(key).ok_or_else(|| Error::Invalid("unknown model key".into()))?;
        let (p, q, s) = (pos.to_array(), [quat.x, quat.y, quat.z, quat.w], scale.to_array());
        check(unsafe { sys::b200gs_model_set_transform(m.h, p.as_ptr(), q.as_ptr(), s.as_ptr()) })
    }
    /// `viewer.update_gaussian_transform(queue, size, display_mode, sh_deg, no_sh0)` — scene.rs:803-809
    pub fn update_gaussian_transform(&mut self, size: f32, mode: GaussianDisplayMode, sh_deg: GaussianShDegree, no_sh0: bool) -> Result<(), Error> {
        check(unsafe { sys::b200gs_set_gaussian_transform(self.h, size, mode as u32, sh_deg.0, no_sh0 as u32) })
    }
    /// `update_selection_edit_with_pod` — scene.rs:815-833
    pub fn update_selection_edit_with_pod(&mut self, pod: &GaussianEditPod) -> Result<(), Error> {
        check(unsafe { sys::b200gs_set_selection_edit(self.h, pod) })
    }
    /// `update_selection_highlight` — scene.rs:848
    pub fn update_selection_highlight(&mut self, rgba: Vec4) -> Result<(), Error> {
        let c = rgba.to_array();
        check(unsafe { sys::b200gs_set_selection_highlight(self.h, c.as_ptr()) })
    }
    /// `viewer.update_query(queue, &pod)` — scene.rs:785 (rect / brush in immediate mode, or `kind = TEXTURE`)
    pub fn update_query(&mut self, pod: &QueryPod) -> Result<(), Error> { check(unsafe { sys::b200gs_set_query(self.h, pod) }) }
    /// `query_toolset.render(queue, encoder, &viewer.world_buffers.query_texture)` — scene.rs:767-791: one stroke segment
    /// of the non-immediate tools, accumulated in the viewer's query texture
    pub fn paint_query_texture(&mut self, stroke: &QueryPod) -> Result<(), Error> { check(unsafe { sys::b200gs_query_texture_paint(self.h, stroke) }) }
    pub fn clear_query_texture(&mut self) -> Result<(), Error> { check(unsafe { sys::b200gs_query_texture_clear(self.h) }) }

    /// `viewer.preprocessor.preprocess(&mut encoder, bind_group, n)` — scene.rs:856-863 (`use_unedited`: the bind group
    /// of scene.rs:858-861)
    pub fn preprocess(&mut self, key: &K, use_unedited: bool) -> Result<(), Error> {
        let m = self.models.get(key).ok_or_else(|| Error::Invalid("unknown model key".into()))?;
        check(unsafe { sys::b200gs_model_preprocess(m.h, use_unedited as i32) })
    }
    /// `viewer.radix_sorter.sort(&mut encoder, bind_group, indirect_args)` — scene.rs:865-869
    pub fn sort(&mut self, key: &K) -> Result<(), Error> {
        let m = self.models.get(key).ok_or_else(|| Error::Invalid("unknown model key".into()))?;
        check(unsafe { sys::b200gs_model_sort(m.h) })
    }
    /// `viewer.postprocessor.postprocess(..)` — scene.rs:604-610
    pub fn postprocess(&mut self, key: &K) -> Result<(), Error> {
        let m = self.models.get(key).ok_or_else(|| Error::Invalid("unknown model key".into()))?;
        check(unsafe { sys::b200gs_model_postprocess(m.h) })
    }
    /// `viewer.renderer.render_with_pass(pass, bind_group, indirect_args)` for each key of `model_render_keys`
    /// (farthest first, scene.rs:533-558, 2302-2314): ONE call for the whole layered frame into a device RGBA8 target.
    pub fn render(&mut self, keys_far_to_near: &[K], rgba8_device: *mut std::ffi::c_void, pitch: usize) -> Result<(), Error> {
        let hs: Vec<*mut sys::b200gs_model> = keys_far_to_near.iter().filter_map(|k| self.models.get(k)).map(|m| m.h).collect();
        check(unsafe { sys::b200gs_render(self.h, hs.as_ptr(), hs.len() as u32, rgba8_device, pitch) })
    }
    /// Headless form: preprocess + sort + render of every model and the read-back of the image into `rgba8` (W*H*4 bytes).
    pub fn render_frame_host(&mut self, keys_far_to_near: &[K], camera: &impl CameraTrait, rgba8: &mut [u8]) -> Result<(), Error> {
        assert!(rgba8.len() >= (self.size.x * self.size.y * 4) as usize);
        let hs: Vec<*mut sys::b200gs_model> = keys_far_to_near.iter().filter_map(|k| self.models.get(k)).map(|m| m.h).collect();
        let (v, p) = (camera.view().to_cols_array(), camera.projection(self.size.x as f32 / self.size.y as f32).to_cols_array());
        check(unsafe { sys::b200gs_render_frame_host(self.h, hs.as_ptr(), hs.len() as u32, v.as_ptr(), p.as_ptr(), rgba8.as_mut_ptr() as *mut _) })
    }
    /// `QueryHitPod::new(pos)` + `gs::query::download` — scene.rs:617-676: the splats under one pixel, front to back
    pub fn query_hits(&mut self, keys_far_to_near: &[K], pixel: Vec2, cap: usize) -> Result<Vec<QueryHitResultPod>, Error> {
        let hs: Vec<*mut sys::b200gs_model> = keys_far_to_near.iter().filter_map(|k| self.models.get(k)).map(|m| m.h).collect();
        let mut out = vec![QueryHitResultPod { model: 0, index: 0, alpha: 0.0, depth: 0.0 }; cap];
        let mut n = 0u64;
        check(unsafe { sys::b200gs_query_hits(self.h, hs.as_ptr(), hs.len() as u32, pixel.x as u32, pixel.y as u32, out.as_mut_ptr(), cap as u64, &mut n) })?;
        out.truncate((n as usize).min(cap));
        Ok(out)
    }
    /// `queue.submit(..); device.poll(Maintain::Wait)` — scene.rs:613-614, 872-873
    pub fn sync(&mut self) -> Result<(), Error> { check(unsafe { sys::b200gs_sync(self.h) }) }
}

impl<G: GaussianPod, K: std::hash::Hash + Eq + Clone> Drop for MultiModelViewer<G, K> {
    fn drop(&mut self) { unsafe { sys::b200gs_viewer_destroy(self.h); } }   // (destroys its models, too)
}
