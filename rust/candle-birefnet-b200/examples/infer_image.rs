//! The reference's CLI (examples/infer_image.rs) on the B200 library: the image never leaves 8 bits on the host --
//! resize + normalise, forward_logits, sigmoid, u8 and the resize back all run on the device (`brn_infer_rgb8`).
//! Usage: infer_image <model.safetensors> <input image> <output.png>     (no HF download: there is no network here)
use candle_birefnet_b200::{BiRefNet, BiRefNetConfig};
use image::{GrayImage, ImageReader};

fn main() -> anyhow::Result<()> {
    let a: Vec<String> = std::env::args().collect();
    if a.len() != 4 {
        anyhow::bail!("usage: {} <model.safetensors> <input> <output.png>", a[0]);
    }
    let model = BiRefNet::from_safetensors(BiRefNetConfig::swin_l(), &a[1])?;
    let img = ImageReader::open(&a[2])?.decode()?.to_rgb8();
    let (w, h) = img.dimensions();
    let t0 = std::time::Instant::now();
    let mask = model.infer_rgb8(img.as_raw(), h as usize, w as usize)?;
    println!("inference (pre + forward + post, {}x{}): {:?}", w, h, t0.elapsed());
    GrayImage::from_raw(w, h, mask).expect("mask size").save(&a[3])?;
    Ok(())
}
