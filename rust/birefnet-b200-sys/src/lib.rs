//! Raw binding of `include/birefnet_b200.h`, one declaration per exported function, in header order.
//! `tests/test_rust_sources.py` keeps this file and the header in step (names, parameter counts, constants).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

/// Opaque model handle (`typedef struct brn_model brn_model`).
#[repr(C)]
pub struct brn_model {
    _private: [u8; 0],
}
/// Opaque multi-GPU handle (`typedef struct brn_sharded brn_sharded`).
#[repr(C)]
pub struct brn_sharded {
    _private: [u8; 0],
}

/// `brn_status`
pub type brn_status = c_int;
pub const BRN_OK: brn_status = 0;
pub const BRN_ERR_INVALID: brn_status = 1;
pub const BRN_ERR_CUDA: brn_status = 2;
pub const BRN_ERR_MISSING_TENSOR: brn_status = 3;
pub const BRN_ERR_UNKNOWN_TENSOR: brn_status = 4;
pub const BRN_ERR_SHAPE: brn_status = 5;
pub const BRN_ERR_STATE: brn_status = 6;
pub const BRN_ERR_UNSUPPORTED: brn_status = 7;

/// `brn_dtype`
pub const BRN_F32: c_int = 0;
pub const BRN_BF16: c_int = 1;
pub const BRN_F16: c_int = 2;
/// `brn_precision`
pub const BRN_PREC_FP32: i32 = 0;
pub const BRN_PREC_BF16: i32 = 1;
pub const BRN_PREC_FP16: i32 = 2;
/// `brn_deform_mode`
pub const BRN_DEFORM_CPU_FALLBACK: i32 = 0;
pub const BRN_DEFORM_DEFORMABLE: i32 = 1;

/// Mirror of `brn_config` (SwinConfig, src/swin.rs:13-23, plus the run-time knobs).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct brn_config {
    pub embed_dim: i32,
    pub depths: [i32; 4],
    pub num_heads: [i32; 4],
    pub window_size: i32,
    pub mlp_ratio: i32,
    pub patch_size: i32,
    pub precision: i32,
    pub deform_mode: i32,
    pub micro_batch: i32,
}

extern "C" {
    pub fn brn_config_swin_l(cfg: *mut brn_config);
    pub fn brn_config_swin_b(cfg: *mut brn_config);
    pub fn brn_config_swin_t(cfg: *mut brn_config);
    pub fn brn_config_swin_s(cfg: *mut brn_config);

    // ---- model lifetime: BiRefNet::new(config, vb) (src/birefnet.rs:389-409)
    pub fn brn_model_create(cfg: *const brn_config, device: c_int, out: *mut *mut brn_model) -> brn_status;
    pub fn brn_model_set_tensor(m: *mut brn_model, key: *const c_char, data: *const c_void, dtype: c_int,
                                shape: *const i64, rank: c_int) -> brn_status;
    pub fn brn_model_load_safetensors(m: *mut brn_model, path: *const c_char, n_loaded: *mut i32) -> brn_status;
    pub fn brn_model_num_tensors(m: *const brn_model) -> i32;
    pub fn brn_model_tensor_info(m: *const brn_model, index: i32, key: *mut *const c_char, shape: *mut i64,
                                 rank: *mut i32) -> brn_status;
    pub fn brn_model_finalize(m: *mut brn_model) -> brn_status;
    pub fn brn_model_set_precision(m: *mut brn_model, precision: c_int) -> brn_status;
    pub fn brn_model_set_deform_mode(m: *mut brn_model, deform_mode: c_int) -> brn_status;
    pub fn brn_model_set_cuda_graph(m: *mut brn_model, on: c_int) -> brn_status;
    pub fn brn_model_destroy(m: *mut brn_model);

    // ---- the hot path
    pub fn brn_forward_logits(m: *mut brn_model, x: *const f32, b: i32, h: i32, w: i32, x_is_device: c_int,
                              out: *mut f32, out_is_device: c_int, stream: *mut c_void) -> brn_status;
    pub fn brn_forward(m: *mut brn_model, x: *const f32, b: i32, h: i32, w: i32, x_is_device: c_int,
                       out: *mut f32, out_is_device: c_int, stream: *mut c_void) -> brn_status;
    pub fn brn_backbone_forward(m: *mut brn_model, x: *const f32, b: i32, h: i32, w: i32, x_is_device: c_int,
                                outs: *const *mut f32, out_is_device: c_int, stream: *mut c_void) -> brn_status;
    pub fn brn_features_forward(m: *mut brn_model, x: *const f32, b: i32, h: i32, w: i32, x_is_device: c_int,
                                outs: *const *mut f32, out_is_device: c_int, stream: *mut c_void) -> brn_status;
    pub fn brn_decoder_forward(m: *mut brn_model, x: *const f32, x1: *const f32, x2: *const f32, x3: *const f32,
                               x4: *const f32, b: i32, h: i32, w: i32, is_device: c_int, out: *mut f32,
                               stream: *mut c_void) -> brn_status;

    // ---- image sharding inside one process
    pub fn brn_sharded_create(cfg: *const brn_config, devices: *const i32, n_devices: i32,
                              out: *mut *mut brn_sharded) -> brn_status;
    pub fn brn_sharded_set_tensor(s: *mut brn_sharded, key: *const c_char, data: *const c_void, dtype: c_int,
                                  shape: *const i64, rank: c_int) -> brn_status;
    pub fn brn_sharded_load_safetensors(s: *mut brn_sharded, path: *const c_char, n_loaded: *mut i32) -> brn_status;
    pub fn brn_sharded_finalize(s: *mut brn_sharded) -> brn_status;
    pub fn brn_sharded_num_devices(s: *const brn_sharded) -> i32;
    pub fn brn_sharded_forward_logits(s: *mut brn_sharded, x: *const f32, b: i32, h: i32, w: i32,
                                      out: *mut f32) -> brn_status;
    pub fn brn_sharded_forward(s: *mut brn_sharded, x: *const f32, b: i32, h: i32, w: i32, out: *mut f32) -> brn_status;
    pub fn brn_sharded_destroy(s: *mut brn_sharded);
    pub fn brn_host_alloc(bytes: usize) -> *mut c_void;
    pub fn brn_host_free(p: *mut c_void);

    // ---- the steps either side of the path (examples/infer_image.rs:44-105)
    pub fn brn_preprocess_rgb8(device: c_int, rgb: *const u8, b: i32, h: i32, w: i32, out_h: i32, out_w: i32,
                               out: *mut f32) -> brn_status;
    pub fn brn_postprocess_mask(device: c_int, logits: *const f32, b: i32, h: i32, w: i32, orig_h: i32, orig_w: i32,
                                out: *mut u8) -> brn_status;
    pub fn brn_infer_rgb8(m: *mut brn_model, rgb: *const u8, b: i32, h: i32, w: i32, out_h: i32, out_w: i32,
                          masks: *mut u8) -> brn_status;

    // ---- operator level
    pub fn brn_window_attention(device: c_int, precision: c_int, qkv: *const f32, bias: *const f32, n_windows: i32,
                                heads: i32, window_size: i32, hp: i32, wp: i32, shift: i32, out: *mut f32) -> brn_status;
    pub fn brn_deform_conv2d(device: c_int, precision: c_int, x: *const f32, offset: *const f32, mask: *const f32,
                             weight: *const f32, bias: *const f32, b: i32, c: i32, h: i32, w: i32, o: i32, k: i32,
                             stride: i32, padding: i32, out: *mut f32) -> brn_status;
    pub fn brn_deformable_conv2d(device: c_int, precision: c_int, deform_mode: c_int, x: *const f32, b: i32, c: i32,
                                 h: i32, w: i32, offset_w: *const f32, offset_b: *const f32, modulator_w: *const f32,
                                 modulator_b: *const f32, regular_w: *const f32, regular_b: *const f32, o: i32, k: i32,
                                 stride: i32, padding: i32, out: *mut f32) -> brn_status;
    pub fn brn_linear(device: c_int, precision: c_int, a: *const f32, w: *const f32, bias: *const f32,
                      residual: *const f32, m: i32, n: i32, k: i32, act: i32, out: *mut f32) -> brn_status;
    pub fn brn_ln_linear(device: c_int, precision: c_int, x: *const f32, gamma: *const f32, beta: *const f32,
                         w: *const f32, bias: *const f32, m: i32, n: i32, k: i32, act: i32, out: *mut f32) -> brn_status;
    pub fn brn_swin_mlp(device: c_int, precision: c_int, x: *const f32, gamma: *const f32, beta: *const f32,
                        w1: *const f32, b1: *const f32, w2: *const f32, b2: *const f32, m: i32, c: i32, hidden: i32,
                        fused: i32, out: *mut f32, out_mean_rstd: *mut f32) -> brn_status;
    pub fn brn_conv2d(device: c_int, precision: c_int, x: *const f32, weight: *const f32, bias: *const f32, b: i32,
                      c: i32, h: i32, w: i32, o: i32, k: i32, act: i32, out: *mut f32) -> brn_status;
    pub fn brn_bench_op(device: c_int, precision: c_int, kind: c_int, b: i32, h: i32, w: i32, c: i32, n: i32, k: i32,
                        act: i32, with_res: i32, out_f32: i32, iters: i32, ms_out: *mut f32) -> brn_status;

    // ---- introspection
    pub fn brn_launch_count(m: *const brn_model) -> i64;
    pub fn brn_launch_count_reset(m: *mut brn_model);
    pub fn brn_profile_enable(m: *mut brn_model, on: c_int);
    pub fn brn_kernel_class_times(m: *const brn_model, ms: *mut f32, flops: *mut f64, bytes: *mut f64,
                                  counts: *mut i32, cap: i32) -> i32;
    pub fn brn_profile_get(m: *const brn_model, names: *mut *const *const c_char, ms: *mut *const f32,
                           flops: *mut *const f64) -> i32;
    pub fn brn_last_error() -> *const c_char;
    pub fn brn_version() -> *const c_char;
}
