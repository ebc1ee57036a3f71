"""Texture PSNR as the reference evaluates it (utils/metric_utils/psnr_ssmi.py:23-42), on the
8-bit atlas image that demo.py:283-301 writes.  Host-side numpy; evaluation only."""
import math

import numpy as np


def calculate_psnr(img1, img2, border=0):
    """img1, img2: [H,W,C] uint8 in 0..255 -> PSNR in dB (inf when identical)."""
    if not img1.shape == img2.shape:
        raise ValueError('Input images must have the same dimensions.')
    h, w = img1.shape[:2]
    a = img1[border:h - border, border:w - border].astype(np.float64)
    b = img2[border:h - border, border:w - border].astype(np.float64)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return float('inf')
    return 20 * math.log10(255.0 / math.sqrt(mse))


def atlas_to_uint8(atlas):
    """demo.py:283-301: float atlas [R,R,3] in [0,1] -> uint8 image with flipped rows."""
    img = np.asarray(atlas, dtype=np.float32) * (255 / 1)
    return np.ascontiguousarray(img.clip(0, 255).astype(np.uint8)[::-1])
