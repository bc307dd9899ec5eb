// Mirrors INTEGRATION.md.  Not compiled in the build image (no cargo / rustc there); the same C ABI
// is exercised by depthhead_b200/capi.py in the test-suite.
extern crate image;

//! Drop-in for depthhead's `HoughPrediction` prediction path on an NVIDIA B200.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct DhForest { _p: [u8; 0] }
#[repr(C)] pub struct DhCtx { _p: [u8; 0] }

/// include/depthhead_cuda.h: dh_result  <->  prediction.rs:259-267 PredictionResult
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct DhResult {
    pub mid_point: [f32; 3],
    _pad: u32,
    pub rotation: [f64; 3],
    pub bounding_box: [u32; 4],
}

extern "C" {
    fn dh_last_error() -> *const c_char;
    fn dh_forest_from_json(json: *const c_char, len: usize, out: *mut *mut DhForest) -> c_int;
    fn dh_forest_free(f: *mut DhForest);
    fn dh_forest_get_stepwidth(f: *const DhForest) -> u32;
    fn dh_forest_set_stepwidth(f: *mut DhForest, v: u32) -> c_int;
    fn dh_forest_get_meanshift_iterations(f: *const DhForest) -> u32;
    fn dh_forest_set_meanshift_iterations(f: *mut DhForest, v: u32) -> c_int;
    fn dh_forest_get_sigma(f: *const DhForest) -> f32;
    fn dh_forest_set_sigma(f: *mut DhForest, v: f32) -> c_int;
    fn dh_ctx_create(device: c_int, out: *mut *mut DhCtx) -> c_int;
    fn dh_ctx_free(c: *mut DhCtx);
    fn dh_predict(c: *mut DhCtx, f: *const DhForest, depth: *const u16, w: u32, h: u32, k: *const f32,
                  midp_guess: *const f32, rot_guess: *const f64, out: *mut DhResult) -> c_int;
    fn dh_predict_batch(c: *mut DhCtx, f: *const DhForest, depth: *const u16, n: u32, w: u32, h: u32,
                        k: *const f32, depth_loc: c_int, out: *mut DhResult) -> c_int;
    fn dh_predict_mask(c: *mut DhCtx, f: *const DhForest, depth: *const u16, w: u32, h: u32,
                       mask: *mut u8) -> c_int;
    // Biwi wire formats (src/db_reader/biwi.rs:27-103)
    fn dh_biwi_depth_dims(file: *const u8, len: usize, w: *mut u32, h: *mut u32) -> c_int;
    fn dh_biwi_decode_depth(c: *mut DhCtx, blob: *const u8, offsets: *const u64, n: u32, w: u32, h: u32,
                            out: *mut u16, out_loc: c_int) -> c_int;
    fn dh_predict_batch_biwi(c: *mut DhCtx, f: *const DhForest, blob: *const u8, offsets: *const u64, n: u32,
                             w: u32, h: u32, k: *const f32, out: *mut DhResult) -> c_int;
    fn dh_biwi_parse_cal(text: *const c_char, len: usize, k: *mut f32) -> c_int;
    fn dh_biwi_parse_pose(file: *const u8, len: usize, k: *const f32, pos3d: *mut f32, pos2d: *mut f32,
                          rot: *mut f32) -> c_int;
    // training (src/hough/prediction.rs:106-234, src/hough/houghforest.rs:196-311)
    fn dh_train_learn(c: *mut DhCtx, p: *const DhTrainParams, n_frames: u32, w: u32, h: u32, depth: *const u16,
                      mask: *const u8, k: *const f32, pos3d: *const f32, rot: *const f32, out: *mut *mut DhForest) -> c_int;
    fn dh_train_forest(c: *mut DhCtx, p: *const DhTrainParams, patches: *const u16, n: u64, is_object: *const u8,
                       offsets: *const f32, rotations: *const f64, out: *mut *mut DhForest) -> c_int;
    fn dh_forest_to_json(f: *const DhForest, buf: *mut c_char, cap: usize, needed: *mut usize) -> c_int;
}

/// HoughLearning::new arguments (prediction.rs:106-116), learn's sigma, and a seed (the reference
/// draws from thread_rng; here the same seed always gives the same forest)
#[repr(C)]
#[derive(Clone, Copy)]
pub struct DhTrainParams {
    pub stepwidth: u32, pub subimage_width: u32, pub subimage_height: u32,
    pub max_depth: u32, pub n_trees: u32, pub subset_per_tree: u32,
    pub subrect_feature_scale: f64, pub features_per_node: u32, pub min_subset_size: u32,
    pub steepness: f64, pub gaussian_sigma: f32, _pad: u32, pub seed: u64,
}

#[derive(Debug)]
pub struct Error { pub code: i32, pub message: String }
fn check(rc: c_int) -> Result<(), Error> {
    if rc == 0 { return Ok(()); }
    let message = unsafe { CStr::from_ptr(dh_last_error()) }.to_string_lossy().into_owned();
    Err(Error { code: rc, message })
}

/// types.rs:405-446.  The inverse is computed inside the library (adjugate / det, f32).
#[derive(Clone, Copy)]
pub struct IntrinsicMatrix(pub [[f32; 3]; 3]);
impl IntrinsicMatrix {
    pub fn new(m: [[f32; 3]; 3]) -> Self { IntrinsicMatrix(m) }
    pub fn default_kinect_intrinsic() -> Self {           // types.rs:418-420
        IntrinsicMatrix([[560.0, 0.0, 320.0], [0.0, 560.0, 240.0], [0.0, 0.0, 1.0]])
    }
}

/// prediction.rs:259-267
pub struct PredictionResult { pub mid_point: [f32; 3], pub rotation: [f64; 3], pub bounding_box: [u32; 4] }

/// prediction.rs:239-256.  Holds the flattened model and one GPU context; like the reference it is
/// Send but not Sync (one predictor per thread).
pub struct HoughPrediction { forest: *mut DhForest, ctx: *mut DhCtx }
unsafe impl Send for HoughPrediction {}

impl HoughPrediction {
    /// replaces `serde_json::from_str::<HoughPrediction>(&text)` (Readme.md:82-86)
    pub fn from_json_on(text: &str, device: i32) -> Result<Self, Error> {
        let (mut f, mut c) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(unsafe { dh_forest_from_json(text.as_ptr() as *const c_char, text.len(), &mut f) })?;
        if let Err(e) = check(unsafe { dh_ctx_create(device, &mut c) }) {
            unsafe { dh_forest_free(f) };
            return Err(e);
        }
        Ok(HoughPrediction { forest: f, ctx: c })
    }
    pub fn stepwidth(&self) -> u32 { unsafe { dh_forest_get_stepwidth(self.forest) } }
    pub fn set_stepwidth(&mut self, v: u32) -> Result<(), Error> { check(unsafe { dh_forest_set_stepwidth(self.forest, v) }) }
    pub fn meanshift_iterations(&self) -> u32 { unsafe { dh_forest_get_meanshift_iterations(self.forest) } }
    pub fn set_meanshift_iterations(&mut self, v: u32) -> Result<(), Error> {
        check(unsafe { dh_forest_set_meanshift_iterations(self.forest, v) })
    }
    pub fn sigma(&self) -> f32 { unsafe { dh_forest_get_sigma(self.forest) } }                 // prediction.rs:329
    pub fn update_sigma(&mut self, v: f32) { unsafe { dh_forest_set_sigma(self.forest, v); } } // prediction.rs:320

    /// prediction.rs:397-409 — `img` is `DepthImage = ImageBuffer<Luma<u16>, Vec<u16>>` (types.rs:10):
    /// contiguous row-major u16, so `img.as_ptr()` is passed as is.
    pub fn predict_parameter_parallel(&self, img: &image::ImageBuffer<image::Luma<u16>, Vec<u16>>,
                                      intrinsic: &IntrinsicMatrix, midp_guess: Option<[f32; 3]>,
                                      rot_guess: Option<[f64; 3]>) -> PredictionResult {
        let mut out = DhResult::default();
        let mg = midp_guess.as_ref().map_or(std::ptr::null(), |g| g.as_ptr());
        let rg = rot_guess.as_ref().map_or(std::ptr::null(), |g| g.as_ptr());
        let rc = unsafe {
            dh_predict(self.ctx, self.forest, img.as_ptr(), img.width(), img.height(),
                       intrinsic.0.as_ptr() as *const f32, mg, rg, &mut out)
        };
        // the reference signature is infallible and panics on degenerate input (prediction.rs:565,716)
        check(rc).expect("depthhead-cuda: prediction failed");
        PredictionResult { mid_point: out.mid_point, rotation: out.rotation, bounding_box: out.bounding_box }
    }
    /// identical results to the parallel variant (prediction.rs:376-388)
    pub fn predict_parameter(&self, img: &image::ImageBuffer<image::Luma<u16>, Vec<u16>>, k: &IntrinsicMatrix,
                             m: Option<[f32; 3]>, r: Option<[f64; 3]>) -> PredictionResult {
        self.predict_parameter_parallel(img, k, m, r)
    }
    /// batched form: `frames` = n images back to back; results in frame order
    pub fn predict_batch(&self, frames: &[u16], n: u32, w: u32, h: u32, k: &IntrinsicMatrix) -> Result<Vec<DhResult>, Error> {
        assert_eq!(frames.len(), (n as usize) * (w as usize) * (h as usize));
        let mut out = vec![DhResult::default(); n as usize];
        check(unsafe { dh_predict_batch(self.ctx, self.forest, frames.as_ptr(), n, w, h,
                                        k.0.as_ptr() as *const f32, 0 /* DH_DEPTH_HOST */, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// prediction.rs:850-905
    pub fn predict_mask(&self, img: &image::ImageBuffer<image::Luma<u16>, Vec<u16>>)
        -> image::ImageBuffer<image::Luma<u8>, Vec<u8>> {
        let mut buf = vec![0u8; (img.width() * img.height()) as usize];
        check(unsafe { dh_predict_mask(self.ctx, self.forest, img.as_ptr(), img.width(), img.height(), buf.as_mut_ptr()) })
            .expect("depthhead-cuda: predict_mask failed");
        image::ImageBuffer::from_raw(img.width(), img.height(), buf).unwrap()
    }
}
impl Drop for HoughPrediction {
    fn drop(&mut self) { unsafe { dh_ctx_free(self.ctx); dh_forest_free(self.forest); } }
}

impl HoughPrediction {
    /// HoughLearning::learn (prediction.rs:145-234) on the GPU: `depth`/`mask` are n frames back to
    /// back, `k` one row-major 3x3 matrix per frame, `pos3d`/`rot` the annotated head pose per frame.
    pub fn learn_on(params: &DhTrainParams, n: u32, w: u32, h: u32, depth: &[u16], mask: &[u8], k: &[f32],
                    pos3d: &[f32], rot: &[f32], device: i32) -> Result<Self, Error> {
        let (mut f, mut c) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(unsafe { dh_ctx_create(device, &mut c) })?;
        if let Err(e) = check(unsafe { dh_train_learn(c, params, n, w, h, depth.as_ptr(), mask.as_ptr(), k.as_ptr(),
                                                      pos3d.as_ptr(), rot.as_ptr(), &mut f) }) {
            unsafe { dh_ctx_free(c) };
            return Err(e);
        }
        Ok(HoughPrediction { forest: f, ctx: c })
    }
    /// what `tojson(&tree, filename)` writes (examples/hough_tree_trainer.rs:182)
    pub fn to_json(&self) -> Result<String, Error> {
        let mut need = 0usize;
        check(unsafe { dh_forest_to_json(self.forest, std::ptr::null_mut(), 0, &mut need) })?;
        let mut buf = vec![0u8; need];
        check(unsafe { dh_forest_to_json(self.forest, buf.as_mut_ptr() as *mut c_char, need, &mut need) })?;
        Ok(String::from_utf8_lossy(&buf).into_owned())
    }
}

/// src/db_reader/biwi.rs: the three file formats on the way to the prediction path.
pub mod biwi {
    use super::*;
    /// Compressed depth files of one sequence, packed for the GPU: every file starts at a multiple
    /// of 16 bytes of `blob`; `offsets` has one more entry than there are files.
    pub struct PackedFiles { pub blob: Vec<u8>, pub offsets: Vec<u64>, pub width: u32, pub height: u32 }
    impl PackedFiles {
        pub fn new(files: &[Vec<u8>]) -> Result<Self, Error> {
            let (mut w, mut h) = (0u32, 0u32);
            if let Some(first) = files.first() {
                check(unsafe { dh_biwi_depth_dims(first.as_ptr(), first.len(), &mut w, &mut h) })?;
            }
            let (mut blob, mut offsets) = (Vec::new(), Vec::with_capacity(files.len() + 1));
            for f in files {
                offsets.push(blob.len() as u64);
                blob.extend_from_slice(f);
                while blob.len() % 16 != 0 { blob.push(0); }
            }
            offsets.push(blob.len() as u64);
            blob.extend_from_slice(&[0u8; 16]);
            Ok(PackedFiles { blob, offsets, width: w, height: h })
        }
    }
    /// read_cal (biwi.rs:27-60)
    pub fn read_cal(text: &str) -> Result<IntrinsicMatrix, Error> {
        let mut k = [[0f32; 3]; 3];
        check(unsafe { dh_biwi_parse_cal(text.as_ptr() as *const c_char, text.len(), k.as_mut_ptr() as *mut f32) })?;
        Ok(IntrinsicMatrix(k))
    }
    /// read_gt (biwi.rs:63-77): (pos3d, pos2d, rot)
    pub fn read_gt(file: &[u8], k: &IntrinsicMatrix) -> Result<([f32; 3], [f32; 2], [f32; 3]), Error> {
        let (mut p3, mut p2, mut rot) = ([0f32; 3], [0f32; 2], [0f32; 3]);
        check(unsafe { dh_biwi_parse_pose(file.as_ptr(), file.len(), k.0.as_ptr() as *const f32, p3.as_mut_ptr(),
                                          p2.as_mut_ptr(), rot.as_mut_ptr()) })?;
        Ok((p3, p2, rot))
    }
    impl HoughPrediction {
        /// read_depth (biwi.rs:81-103) for every file, expanded on the GPU: n images back to back
        pub fn read_depth(&self, files: &PackedFiles) -> Result<Vec<u16>, Error> {
            let n = (files.offsets.len() - 1) as u32;
            let mut out = vec![0u16; n as usize * files.width as usize * files.height as usize];
            check(unsafe { dh_biwi_decode_depth(self.ctx, files.blob.as_ptr(), files.offsets.as_ptr(), n, files.width,
                                                files.height, out.as_mut_ptr(), 0 /* DH_DEPTH_HOST */) })?;
            Ok(out)
        }
        /// what examples/db_evaluate.rs:296 does per file — read_depth + predict_parameter(img, K, None,
        /// None) — for a whole sequence, the compressed bytes crossing PCIe
        pub fn predict_files(&self, files: &PackedFiles, k: &IntrinsicMatrix) -> Result<Vec<DhResult>, Error> {
            let n = (files.offsets.len() - 1) as u32;
            let mut out = vec![DhResult::default(); n as usize];
            check(unsafe { dh_predict_batch_biwi(self.ctx, self.forest, files.blob.as_ptr(), files.offsets.as_ptr(), n,
                                                 files.width, files.height, k.0.as_ptr() as *const f32, out.as_mut_ptr()) })?;
            Ok(out)
        }
    }
}
