//! The reference crate's public surface for the forward path (`src/lib.rs:6-14` of imperatormk/candle-birefnet),
//! backed by `libbirefnet_b200.so` instead of candle ops:
//!
//! | reference                                           | here                                   |
//! |-----------------------------------------------------|----------------------------------------|
//! | `BiRefNetConfig::swin_l()`          birefnet.rs:64  | `BiRefNetConfig::swin_l()`             |
//! | `BiRefNet::new(config, vb)`         birefnet.rs:389 | `BiRefNet::new(config, vb)`            |
//! | `forward_logits(&Tensor)`           birefnet.rs:412 | `BiRefNet::forward_logits`             |
//! | `forward` / `impl Module`           birefnet.rs:466 | `BiRefNet::forward`, `impl Module`     |
//! | `backbone.forward(&x)`              swin.rs:768     | `BiRefNet::backbone_forward`           |
//! | `DeformableConv2d::{new, forward}`  deform_conv.rs  | `DeformableConv2d::{new, forward}`     |
//!
//! NOT compiled in this repository (no Rust toolchain in the build image); see `rust/README.md`.
use birefnet_b200_sys as sys;
use candle_core::{DType, Device, Module, Result, Tensor};
use candle_nn::VarBuilder;
use std::ffi::{CStr, CString};

fn check(st: sys::brn_status) -> Result<()> {
    if st == sys::BRN_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(sys::brn_last_error()) }.to_string_lossy().into_owned();
    candle_core::bail!("birefnet_b200 (status {st}): {msg}")
}

/// Arithmetic of the path (`brn_precision`): `Fp32` = SIMT fp32 path (max |dlogit| <= 1e-3 vs candle CPU),
/// `Fp16` / `Bf16` = tcgen05 tensor-core path.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Precision {
    Fp32,
    Bf16,
    Fp16,
}
/// What `DeformConvASPP::forward` computes (src/aspp.rs:168-187): `CpuFallback` = `regular_conv(x)` (candle on
/// `Device::Cpu`), `Deformable` = the Metal path's modulated deformable convolution.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum DeformMode {
    CpuFallback,
    Deformable,
}

/// `BiRefNetConfig` (src/birefnet.rs:13-30); the fields the reference never reads in forward are kept for source
/// compatibility.
#[derive(Clone, Debug)]
pub struct BiRefNetConfig {
    pub size: (usize, usize),
    pub backbone: String,
    pub mul_scl_ipt: bool,
    pub ms_supervision: bool,
    pub dec_ipt: bool,
    pub precision: Precision,
    pub deform_mode: DeformMode,
    pub device: i32,
}
impl BiRefNetConfig {
    pub fn swin_l() -> Self {
        Self {
            size: (1024, 1024),
            backbone: "swin_v1_l".into(),
            mul_scl_ipt: true,
            ms_supervision: true,
            dec_ipt: true,
            precision: Precision::Fp16,
            deform_mode: DeformMode::Deformable,
            device: 0,
        }
    }
    fn raw(&self) -> sys::brn_config {
        let mut c = sys::brn_config::default();
        unsafe { sys::brn_config_swin_l(&mut c) };
        c.precision = match self.precision {
            Precision::Fp32 => sys::BRN_PREC_FP32,
            Precision::Bf16 => sys::BRN_PREC_BF16,
            Precision::Fp16 => sys::BRN_PREC_FP16,
        };
        c.deform_mode = match self.deform_mode {
            DeformMode::CpuFallback => sys::BRN_DEFORM_CPU_FALLBACK,
            DeformMode::Deformable => sys::BRN_DEFORM_DEFORMABLE,
        };
        c
    }
}

fn host_f32(t: &Tensor) -> Result<Vec<f32>> {
    t.to_dtype(DType::F32)?.to_device(&Device::Cpu)?.contiguous()?.flatten_all()?.to_vec1::<f32>()
}

pub struct BiRefNet {
    pub config: BiRefNetConfig,
    h: *mut sys::brn_model,
}
// the handle serialises planning / launching internally and keeps two calls in flight (include/birefnet_b200.h)
unsafe impl Send for BiRefNet {}
unsafe impl Sync for BiRefNet {}

impl BiRefNet {
    /// Same contract as the reference's constructor (src/birefnet.rs:389-409): every tensor is fetched through
    /// `vb.get(shape, key)`, so a missing key or a wrong shape fails where candle's own constructor would.
    pub fn new(config: BiRefNetConfig, vb: VarBuilder) -> Result<Self> {
        let c = config.raw();
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::brn_model_create(&c, config.device, &mut h) })?;
        let me = Self { config, h };
        let n = unsafe { sys::brn_model_num_tensors(h) };
        for i in 0..n {
            let (mut key, mut shape, mut rank) = (std::ptr::null(), [0i64; 4], 0i32);
            check(unsafe { sys::brn_model_tensor_info(h, i, &mut key, shape.as_mut_ptr(), &mut rank) })?;
            let k = unsafe { CStr::from_ptr(key) }.to_str().unwrap().to_owned();
            let dims: Vec<usize> = shape[..rank as usize].iter().map(|&d| d as usize).collect();
            let v = host_f32(&vb.get(dims.as_slice(), &k)?)?;
            let ck = CString::new(k).unwrap();
            check(unsafe {
                sys::brn_model_set_tensor(h, ck.as_ptr(), v.as_ptr() as _, sys::BRN_F32, shape.as_ptr(), rank)
            })?;
        }
        check(unsafe { sys::brn_model_finalize(h) })?;
        Ok(me)
    }

    /// `candle_core::safetensors::load` + `VarBuilder::from_tensors` + `BiRefNet::new` (examples/infer_image.rs:35-40)
    /// without materialising candle tensors: the library reads the file itself.
    pub fn from_safetensors(config: BiRefNetConfig, path: &str) -> Result<Self> {
        let c = config.raw();
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::brn_model_create(&c, config.device, &mut h) })?;
        let me = Self { config, h };
        let p = CString::new(path).unwrap();
        let mut n = 0i32;
        check(unsafe { sys::brn_model_load_safetensors(h, p.as_ptr(), &mut n) })?;
        check(unsafe { sys::brn_model_finalize(h) })?;
        Ok(me)
    }

    fn run(&self, x: &Tensor, sigmoid: bool) -> Result<Tensor> {
        let (b, c, h, w) = x.dims4()?;
        if c != 3 {
            candle_core::bail!("expected [B,3,H,W], got {:?}", x.shape())
        }
        let xin = host_f32(x)?;
        let mut out = vec![0f32; b * h * w];
        let f = if sigmoid { sys::brn_forward } else { sys::brn_forward_logits };
        check(unsafe {
            f(self.h, xin.as_ptr(), b as i32, h as i32, w as i32, 0, out.as_mut_ptr(), 0, std::ptr::null_mut())
        })?;
        Tensor::from_vec(out, (b, 1, h, w), x.device())
    }

    /// `BiRefNet::forward_logits` (src/birefnet.rs:412-461): `[B,3,H,W]` -> `[B,1,H,W]`.
    pub fn forward_logits(&self, x: &Tensor) -> Result<Tensor> {
        self.run(x, false)
    }
    /// `BiRefNet::forward` (src/birefnet.rs:466-469).
    pub fn forward(&self, x: &Tensor) -> Result<Tensor> {
        self.run(x, true)
    }
    /// `SwinTransformer::forward` (src/swin.rs:768-797): four NCHW feature maps.
    pub fn backbone_forward(&self, x: &Tensor) -> Result<Vec<Tensor>> {
        let (b, _, h, w) = x.dims4()?;
        let xin = host_f32(x)?;
        let chans = [192usize, 384, 768, 1536];
        let mut bufs: Vec<Vec<f32>> =
            (0..4).map(|i| vec![0f32; b * chans[i] * (h >> (2 + i)) * (w >> (2 + i))]).collect();
        let ptrs: Vec<*mut f32> = bufs.iter_mut().map(|v| v.as_mut_ptr()).collect();
        check(unsafe {
            sys::brn_backbone_forward(self.h, xin.as_ptr(), b as i32, h as i32, w as i32, 0, ptrs.as_ptr(), 0,
                                      std::ptr::null_mut())
        })?;
        bufs.into_iter()
            .enumerate()
            .map(|(i, v)| Tensor::from_vec(v, (b, chans[i], h >> (2 + i), w >> (2 + i)), x.device()))
            .collect()
    }
    /// examples/infer_image.rs:44-105 on the device: RGB8 `[h,w,3]` in, u8 mask `[h,w]` out.
    pub fn infer_rgb8(&self, rgb: &[u8], h: usize, w: usize) -> Result<Vec<u8>> {
        let mut mask = vec![0u8; h * w];
        let (ih, iw) = self.config.size;
        check(unsafe {
            sys::brn_infer_rgb8(self.h, rgb.as_ptr(), 1, h as i32, w as i32, ih as i32, iw as i32, mask.as_mut_ptr())
        })?;
        Ok(mask)
    }
}
impl Drop for BiRefNet {
    fn drop(&mut self) {
        unsafe { sys::brn_model_destroy(self.h) }
    }
}
impl Module for BiRefNet {
    fn forward(&self, x: &Tensor) -> Result<Tensor> {
        BiRefNet::forward(self, x)
    }
}

/// One handle + one host thread per GPU, contiguous image split, no collective.
pub struct ShardedBiRefNet {
    h: *mut sys::brn_sharded,
}
unsafe impl Send for ShardedBiRefNet {}
unsafe impl Sync for ShardedBiRefNet {}
impl ShardedBiRefNet {
    pub fn from_safetensors(config: &BiRefNetConfig, path: &str, devices: &[i32]) -> Result<Self> {
        let c = config.raw();
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::brn_sharded_create(&c, devices.as_ptr(), devices.len() as i32, &mut h) })?;
        let me = Self { h };
        let p = CString::new(path).unwrap();
        let mut n = 0i32;
        check(unsafe { sys::brn_sharded_load_safetensors(h, p.as_ptr(), &mut n) })?;
        check(unsafe { sys::brn_sharded_finalize(h) })?;
        Ok(me)
    }
    pub fn forward_logits(&self, x: &Tensor) -> Result<Tensor> {
        let (b, _, h, w) = x.dims4()?;
        let xin = host_f32(x)?;
        let mut out = vec![0f32; b * h * w];
        check(unsafe {
            sys::brn_sharded_forward_logits(self.h, xin.as_ptr(), b as i32, h as i32, w as i32, out.as_mut_ptr())
        })?;
        Tensor::from_vec(out, (b, 1, h, w), x.device())
    }
}
impl Drop for ShardedBiRefNet {
    fn drop(&mut self) {
        unsafe { sys::brn_sharded_destroy(self.h) }
    }
}

/// `DeformableConv2d` (src/deform_conv.rs:16-99), re-exported by the reference's crate root (src/lib.rs:13).
pub struct DeformableConv2d {
    offset_w: Vec<f32>,
    offset_b: Vec<f32>,
    modulator_w: Vec<f32>,
    modulator_b: Vec<f32>,
    regular_w: Vec<f32>,
    regular_b: Vec<f32>,
    in_channels: usize,
    out_channels: usize,
    kernel_size: usize,
    stride: usize,
    padding: usize,
    pub precision: Precision,
    pub deform_mode: DeformMode,
    pub device: i32,
}
impl DeformableConv2d {
    pub fn new(in_channels: usize, out_channels: usize, kernel_size: usize, stride: usize, padding: usize,
               vb: VarBuilder) -> Result<Self> {
        let k = kernel_size;
        let get = |p: &str, o: usize| -> Result<(Vec<f32>, Vec<f32>)> {
            Ok((host_f32(&vb.pp(p).get((o, in_channels, k, k), "weight")?)?, host_f32(&vb.pp(p).get(o, "bias")?)?))
        };
        let (offset_w, offset_b) = get("offset_conv", 2 * k * k)?;
        let (modulator_w, modulator_b) = get("modulator_conv", k * k)?;
        let (regular_w, regular_b) = get("regular_conv", out_channels)?;
        Ok(Self { offset_w, offset_b, modulator_w, modulator_b, regular_w, regular_b, in_channels, out_channels,
                  kernel_size, stride, padding, precision: Precision::Fp16, deform_mode: DeformMode::Deformable,
                  device: 0 })
    }
    pub fn forward(&self, x: &Tensor) -> Result<Tensor> {
        let (b, c, h, w) = x.dims4()?;
        if c != self.in_channels {
            candle_core::bail!("expected {} input channels, got {c}", self.in_channels)
        }
        let (k, s, p) = (self.kernel_size, self.stride, self.padding);
        let (ho, wo) = ((h + 2 * p - k) / s + 1, (w + 2 * p - k) / s + 1);
        let xin = host_f32(x)?;
        let mut out = vec![0f32; b * self.out_channels * ho * wo];
        let prec = match self.precision { Precision::Fp32 => 0, Precision::Bf16 => 1, Precision::Fp16 => 2 };
        let mode = match self.deform_mode { DeformMode::CpuFallback => 0, DeformMode::Deformable => 1 };
        check(unsafe {
            sys::brn_deformable_conv2d(self.device, prec, mode, xin.as_ptr(), b as i32, c as i32, h as i32, w as i32,
                                       self.offset_w.as_ptr(), self.offset_b.as_ptr(), self.modulator_w.as_ptr(),
                                       self.modulator_b.as_ptr(), self.regular_w.as_ptr(), self.regular_b.as_ptr(),
                                       self.out_channels as i32, k as i32, s as i32, p as i32, out.as_mut_ptr())
        })?;
        Tensor::from_vec(out, (b, self.out_channels, ho, wo), x.device())
    }
}
impl Module for DeformableConv2d {
    fn forward(&self, x: &Tensor) -> Result<Tensor> {
        DeformableConv2d::forward(self, x)
    }
}
