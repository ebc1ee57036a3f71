"""Host-side image IO of the path's optional PNG dumps (outside the hot path).

Mirrors utils/utils_2d.py:351-381 (save_CHW_RGB_img / save_CHW_RGBA_img: `(img*255).astype(uint8)`
truncation) as used at ours_utils.py:873-880 and 924-928."""
import os

import numpy as np


def _save(chw, path):
    from PIL import Image
    arr = (np.asarray(chw) * 255).astype(np.uint8).transpose(1, 2, 0)
    mode = "RGBA" if arr.shape[2] == 4 else "RGB"
    Image.fromarray(np.ascontiguousarray(arr), mode).save(path)


def save_sparse_pngs(sparse, m0, m2, save_path):
    os.makedirs(save_path, exist_ok=True)
    s, a, b = sparse.cpu().numpy(), m0.cpu().numpy(), m2.cpu().numpy()
    for i in range(s.shape[0]):
        rgba = np.concatenate([s[i], (a[i, :1] * b[i, :1])], 0)
        _save(rgba, os.path.join(save_path, f"{i}_sparse.png"))
        _save(a[i], os.path.join(save_path, f"{i}_mask0.png"))
        _save(b[i], os.path.join(save_path, f"{i}_mask2.png"))


def save_inpainted_pngs(inpainted, m0, save_path, rgba=True):
    os.makedirs(save_path, exist_ok=True)
    x, a = inpainted.cpu().numpy(), m0.cpu().numpy()
    for i in range(x.shape[0]):
        img = np.concatenate([x[i], a[i, :1]], 0) if rgba else x[i]
        _save(img, os.path.join(save_path, f"{i}_inpainted.png"))


def load_inpainted_pngs(save_path, view_num, res):
    """demo.py:138-147: reuse `{i}_inpainted.png` when ALL views exist; returns [V,3,res,res]
    float32 numpy in [0,1] or None."""
    from PIL import Image
    out = np.zeros((view_num, 3, res, res), dtype=np.float32)
    for i in range(view_num):
        p = os.path.join(save_path, f"{i}_inpainted.png")
        if not os.path.exists(p):
            return None
        img = np.asarray(Image.open(p).convert("RGB"), dtype=np.float32) / 255.0
        out[i] = img.transpose(2, 0, 1)
    return out


def read_ply_xyzrgb(path):
    """utils/other_utils.py:155-162 without plyfile: the reference's demo clouds are binary
    little-endian PLY with `x y z` float32 and `red green blue` uchar per vertex (15 B/vertex).
    Returns (xyz float32 [N,3], rgb uint8 [N,3])."""
    with open(path, "rb") as f:
        header = b""
        while not header.endswith(b"end_header\n"):
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: no end_header")
            header += line
        text = header.decode("ascii", "replace")
        if "binary_little_endian" not in text:
            raise NotImplementedError("only binary_little_endian PLY clouds are supported")
        n = int([l for l in text.splitlines() if l.startswith("element vertex")][0].split()[-1])
        props = [l.split()[1:] for l in text.splitlines() if l.startswith("property")]
        dt = []
        for typ, name in props:
            dt.append((name, {"float": "<f4", "float32": "<f4", "uchar": "u1", "uint8": "u1",
                              "double": "<f8", "int": "<i4"}[typ]))
        data = np.frombuffer(f.read(n * np.dtype(dt).itemsize), dtype=np.dtype(dt), count=n)
    xyz = np.stack([data["x"], data["y"], data["z"]], -1).astype(np.float32)
    rgb = np.stack([data["red"], data["green"], data["blue"]], -1).astype(np.uint8)
    return xyz, rgb


def normalize_cloud(xyz):
    """demo.py:377-380: centre on the bbox centre, divide by the largest bbox extent."""
    vmin, vmax = xyz.min(0), xyz.max(0)
    xyz = xyz - (vmax + vmin) / np.float32(2.0)
    return (xyz / (vmax - vmin).max()).astype(np.float32)
