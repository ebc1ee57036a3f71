"""Texture PSNR as the reference evaluates it (utils/metric_utils/psnr_ssmi.py:23-42: 8-bit images,
peak 255, mean squared error over all channels), on the atlas image that demo.py:283-301 writes.
Host-side numpy; evaluation only."""
import numpy as np


def calculate_psnr(img1, img2, border=0):
    """img1, img2: [H,W,C] uint8 in 0..255 -> PSNR in dB (inf when identical); `border` pixels
    are cropped on every side first."""
    a, b = np.asarray(img1), np.asarray(img2)
    if a.shape != b.shape:
        raise ValueError(f"image shapes differ: {a.shape} vs {b.shape}")
    if border:
        a, b = a[border:-border, border:-border], b[border:-border, border:-border]
    err = a.astype(np.float64) - b.astype(np.float64)
    mse = float(np.mean(err * err))
    return float("inf") if mse == 0.0 else float(10.0 * np.log10(255.0 * 255.0 / mse))


def atlas_to_uint8(atlas):
    """demo.py:283-301: float atlas [R,R,3] in [0,1] -> the 8-bit image (x255, clipped, truncated,
    rows flipped) the reference saves as model_normalized.png."""
    q = np.clip(np.asarray(atlas, dtype=np.float32) * np.float32(255.0), 0.0, 255.0).astype(np.uint8)
    return np.ascontiguousarray(q[::-1])
