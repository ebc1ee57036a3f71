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


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2",
              "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4",
              "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8",
              "float64": "f8"}


def read_ply_xyzrgb(path):
    """utils/other_utils.py:155-162 without plyfile.  The reference's demo clouds are binary
    little-endian PLY with `x y z` float32 and `red green blue` uchar per vertex (15 B/vertex);
    like plyfile this reader also accepts big-endian and ASCII bodies, CRLF headers, extra
    vertex properties, and further elements (e.g. `face` with a `property list`) around the
    vertex element.  Returns (xyz float32 [N,3], rgb uint8 [N,3])."""
    with open(path, "rb") as f:
        lines = []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: no end_header")
            txt = line.decode("ascii", "replace").strip()  # strips \r\n as well as \n
            if txt == "end_header":
                break
            lines.append(txt)
        if not lines or lines[0] != "ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = None
        elements = []  # [name, count, [(kind, ...)]] in file order
        for l in lines[1:]:
            tok = l.split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append([tok[1], int(tok[2]), []])
            elif tok[0] == "property":
                if not elements:
                    raise ValueError(f"{path}: property before any element")
                if tok[1] == "list":
                    if tok[2] not in _PLY_TYPES or tok[3] not in _PLY_TYPES:
                        raise ValueError(f"{path}: unsupported PLY list types {tok[2:4]}")
                    elements[-1][2].append(("list", tok[2], tok[3], tok[4]))
                else:
                    if tok[1] not in _PLY_TYPES:
                        raise ValueError(f"{path}: unsupported PLY property type {tok[1]!r}")
                    elements[-1][2].append(("scalar", tok[1], tok[2]))
        if fmt not in ("binary_little_endian", "binary_big_endian", "ascii"):
            raise ValueError(f"{path}: unsupported PLY format {fmt!r}")
        end = ">" if fmt == "binary_big_endian" else "<"
        data = None
        for name, count, props in elements:
            has_list = any(p[0] == "list" for p in props)
            if name == "vertex":
                if has_list:
                    raise ValueError(f"{path}: list property on the vertex element")
                dt = np.dtype([(p[2], end + _PLY_TYPES[p[1]]) for p in props])
                if fmt == "ascii":
                    rows = [f.readline().split() for _ in range(count)]
                    data = np.zeros(count, dtype=dt)
                    for j, p in enumerate(props):
                        data[p[2]] = np.asarray([r[j] for r in rows], dtype=np.float64)
                else:
                    raw = f.read(count * dt.itemsize)
                    if len(raw) != count * dt.itemsize:
                        raise ValueError(f"{path}: truncated vertex data")
                    data = np.frombuffer(raw, dtype=dt, count=count)
                break
            # an element stored BEFORE the vertices has to be skipped
            if fmt == "ascii":
                for _ in range(count):
                    f.readline()
            elif not has_list:
                f.seek(count * sum(np.dtype(_PLY_TYPES[p[1]]).itemsize for p in props), 1)
            else:
                for _ in range(count):
                    for p in props:
                        if p[0] == "scalar":
                            f.seek(np.dtype(_PLY_TYPES[p[1]]).itemsize, 1)
                        else:
                            cdt = np.dtype(end + _PLY_TYPES[p[1]])
                            k = int(np.frombuffer(f.read(cdt.itemsize), dtype=cdt)[0])
                            f.seek(k * np.dtype(_PLY_TYPES[p[2]]).itemsize, 1)
        if data is None:
            raise ValueError(f"{path}: no vertex element")
    for k in ("x", "y", "z", "red", "green", "blue"):
        if k not in data.dtype.names:
            raise ValueError(f"{path}: vertex property {k!r} missing")
    xyz = np.stack([data["x"], data["y"], data["z"]], -1).astype(np.float32)
    rgb = np.stack([data["red"], data["green"], data["blue"]], -1).astype(np.uint8)
    return xyz, rgb


def normalize_cloud(xyz):
    """demo.py:377-380: centre on the bbox centre, divide by the largest bbox extent."""
    vmin, vmax = xyz.min(0), xyz.max(0)
    xyz = xyz - (vmax + vmin) / np.float32(2.0)
    return (xyz / (vmax - vmin).max()).astype(np.float32)


# ------------------------------------------------------------------------------------------
# "next" row N3: the output tree of demo.py (output/<name>/{models,others,geo})
# ------------------------------------------------------------------------------------------
def savemeshtes2(pointnp_px3, tcoords_px2, facenp_fx3, facetex_fx3, fname):
    """models/get3d/get3d_utils/utils_3d.py:27-64: Wavefront OBJ (`v`, `vt`, `f v/vt`) plus the
    fixed `model_normalized.mtl` next to it; same text, written in one go."""
    fol, na = os.path.split(fname)
    na, _ = os.path.splitext(na)
    with open(os.path.join(fol, "model_normalized.mtl"), "w") as fid:
        fid.write("newmtl material_0\nKd 1 1 1\nKa 0 0 0\nKs 0.4 0.4 0.4\nNs 10\nillum 2\n"
                  "map_Kd %s.png\n" % na)
    f1 = np.asarray(facenp_fx3) + 1
    f2 = np.asarray(facetex_fx3) + 1
    parts = ["mtllib %s.mtl\n" % na]
    parts += ["v %f %f %f\n" % (p[0], p[1], p[2]) for p in np.asarray(pointnp_px3)]
    parts += ["vt %f %f\n" % (p[0], p[1]) for p in np.asarray(tcoords_px2)]
    parts.append("usemtl material_0\n")
    parts += ["f %d/%d %d/%d %d/%d\n" % (a[0], b[0], a[1], b[1], a[2], b[2])
              for a, b in zip(f1, f2)]
    with open(fname, "w") as fid:
        fid.write("".join(parts))


def save_textured_mesh(vertices, uvs, faces, mesh_tex_idx, atlas_img, mask, output_root_path):
    """demo.py:264-307: models/model_normalized.{obj,mtl,png} and others/atlas_wo_background.png.
    atlas_img [R,R,3] f32 cuda, mask [1,R,R,1] bool.  The 8-bit quantisation + flip runs on the
    device (pdr_atlas_to_u8); only the uint8 images cross PCIe."""
    import torch
    from PIL import Image

    from . import _lib
    os.makedirs(os.path.join(output_root_path, "models"), exist_ok=True)
    os.makedirs(os.path.join(output_root_path, "others"), exist_ok=True)
    savemeshtes2(vertices.detach().cpu().numpy(), uvs.detach().cpu().numpy(),
                 faces.detach().cpu().numpy(), mesh_tex_idx.detach().cpu().numpy(),
                 os.path.join(output_root_path, "models", "model_normalized.obj"))
    R = atlas_img.shape[0]
    dev = atlas_img.device
    rgb = torch.empty(R, R, 3, dtype=torch.uint8, device=dev)
    rgba = torch.empty(R, R, 4, dtype=torch.uint8, device=dev)
    m = mask[0, :, :, 0].to(torch.uint8).contiguous()
    _lib.call("pdr_atlas_to_u8", atlas_img.float().contiguous(), m, R, rgb, rgba)
    Image.fromarray(rgb.cpu().numpy(), "RGB").save(
        os.path.join(output_root_path, "models", "model_normalized.png"))
    Image.fromarray(rgba.cpu().numpy(), "RGBA").save(
        os.path.join(output_root_path, "others", "atlas_wo_background.png"))


def loadobjtex(meshfile):
    """models/get3d/get3d_utils/utils_3d.py:94-140: `v`, `vt`, `f v/vt[/vn]` (triangles; quads are
    split like the reference).  Returns (vertices f32 [P,3], uvs f32 [T,2] or None,
    faces int64 [F,3], face_uv_idx int64 [F,3] or None)."""
    v, vt, f, ft = [], [], [], []
    with open(meshfile, "r") as fp:
        for line in fp:
            d = line.split()
            if not d:
                continue
            if d[0] == "v" and len(d) >= 4:
                v.append([float(x) for x in d[1:4]])
            elif d[0] == "vt" and len(d) >= 3:
                vt.append([float(x) for x in d[1:3]])
            elif d[0] == "f" and len(d) in (4, 5):
                c = [x.split("/") for x in d[1:]]
                tris = [(0, 1, 2)] if len(c) == 3 else [(0, 1, 2), (0, 2, 3)]
                for t in tris:
                    f.append([int(c[i][0]) for i in t])
                    if all(len(c[i]) > 1 and c[i][1] != "" for i in t):
                        ft.append([int(c[i][1]) for i in t])
    vertices = np.array(v, dtype=np.float32)
    faces = np.array(f, dtype=np.int64) - 1
    has_uv = len(vt) > 0 and len(ft) == len(f)
    uvs = np.array(vt, dtype=np.float32) if has_uv else None
    face_uv = (np.array(ft, dtype=np.int64) - 1) if has_uv else None
    return vertices, uvs, faces, face_uv


def save_colored_pc_ply(xyz, rgb, path):
    """utils/other_utils.py save_colored_pc_ply: binary little-endian PLY, x y z float32 +
    red green blue uchar (rgb given in [0,1])."""
    xyz = np.asarray(xyz, dtype=np.float32)
    rgb8 = (np.asarray(rgb) * 255).astype(np.uint8)
    rec = np.empty(len(xyz), dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"),
                                    ("green", "u1"), ("blue", "u1")])
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rec["red"], rec["green"], rec["blue"] = rgb8[:, 0], rgb8[:, 1], rgb8[:, 2]
    with open(path, "wb") as fh:
        fh.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\n"
                  "property float y\nproperty float z\nproperty uchar red\nproperty uchar green\n"
                  "property uchar blue\nend_header\n" % len(xyz)).encode("ascii"))
        fh.write(rec.tobytes())
