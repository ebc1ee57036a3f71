"""Test harness: deterministic proxy mesh + atlas for a real point cloud.

The reference ships clouds but no meshes/atlases (POCO weights and xatlas are unavailable), so
parity on `dataset/demo_data/clock.ply` uses a voxel-shell mesh built from the cloud itself:
occupied voxels of a G^3 grid, one quad (2 triangles) per face between an occupied and an empty
voxel, and an atlas with one c x c cell per quad.  It only has to be deterministic and identical
for the reference run and for ours."""
import numpy as np


def voxel_shell(xyz, G=16):
    lo, size = -0.5, 1.0 / G
    idx = np.clip(((xyz - lo) / size).astype(np.int64), 0, G - 1)
    occ = np.zeros((G + 2, G + 2, G + 2), dtype=bool)
    occ[idx[:, 0] + 1, idx[:, 1] + 1, idx[:, 2] + 1] = True
    verts, faces, normals = [], [], []
    vid = {}

    def v(i, j, k):
        key = (i, j, k)
        if key not in vid:
            vid[key] = len(verts)
            verts.append((lo + i * size, lo + j * size, lo + k * size))
        return vid[key]

    dirs = [((1, 0, 0), [(1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1)]),
            ((-1, 0, 0), [(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0)]),
            ((0, 1, 0), [(0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)]),
            ((0, -1, 0), [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1)]),
            ((0, 0, 1), [(0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]),
            ((0, 0, -1), [(0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 0, 0)])]
    for i, j, k in np.argwhere(occ):
        for (dx, dy, dz), corners in dirs:
            if occ[i + dx, j + dy, k + dz]:
                continue
            q = [v(i - 1 + a, j - 1 + b, k - 1 + c) for a, b, c in corners]
            faces.append((q[0], q[1], q[2]))
            faces.append((q[0], q[2], q[3]))
            normals.append((dx, dy, dz))
            normals.append((dx, dy, dz))
    return (np.asarray(verts, dtype=np.float32), np.asarray(faces, dtype=np.int64),
            np.asarray(normals, dtype=np.float32))


def quad_atlas(vertices, faces, res=256):
    """one cell per quad (= consecutive triangle pair 2q, 2q+1 sharing the diagonal v0-v2)."""
    Q = faces.shape[0] // 2
    cells = int(np.ceil(np.sqrt(Q)))
    c = res // cells
    assert c >= 3, "atlas too small for this many quads"
    gb_pos = np.zeros((1, res, res, 3), dtype=np.float32)
    mask = np.zeros((1, res, res, 1), dtype=bool)
    face_id = -np.ones((1, res, res), dtype=np.int64)
    uvs = np.zeros((faces.shape[0] * 3, 2), dtype=np.float32)
    inner = c - 1  # one texel gutter
    t = (np.arange(inner) + 0.5) / inner
    tt, ss = np.meshgrid(t, t, indexing="ij")  # rows -> t, cols -> s
    V = vertices.astype(np.float64)
    for q in range(Q):
        cy, cx = divmod(q, cells)
        f0, f1 = faces[2 * q], faces[2 * q + 1]
        p0, p1, p2, p3 = V[f0[0]], V[f0[1]], V[f0[2]], V[f1[2]]
        pos = ((1 - ss)[..., None] * (1 - tt)[..., None] * p0 + ss[..., None] * (1 - tt)[..., None] * p1
               + ss[..., None] * tt[..., None] * p2 + (1 - ss)[..., None] * tt[..., None] * p3)
        y0, x0 = cy * c, cx * c
        gb_pos[0, y0:y0 + inner, x0:x0 + inner] = pos.astype(np.float32)
        mask[0, y0:y0 + inner, x0:x0 + inner, 0] = True
        face_id[0, y0:y0 + inner, x0:x0 + inner] = np.where(tt > ss, 2 * q + 1, 2 * q)
        corner = lambda s_, t_: ((x0 + s_ * inner) / res, (y0 + t_ * inner) / res)
        uvs[6 * q:6 * q + 6] = [corner(0, 0), corner(1, 0), corner(1, 1),
                                corner(0, 0), corner(1, 1), corner(0, 1)]
    mesh_tex_idx = np.arange(faces.shape[0] * 3, dtype=np.int64).reshape(-1, 3)
    return dict(uvs=uvs, mesh_tex_idx=mesh_tex_idx, gb_pos=gb_pos, mask=mask,
                per_atlas_pixel_face_id=face_id)


def clock_scene(ply_path, G=16, atlas_res=512):
    from pointdreamer_b200.io_utils import normalize_cloud, read_ply_xyzrgb
    xyz, rgb8 = read_ply_xyzrgb(ply_path)
    xyz = normalize_cloud(xyz)
    rgb = rgb8.astype(np.float32) / np.float32(255.0)
    vertices, faces, f_normals = voxel_shell(xyz, G)
    return dict(xyz=xyz, rgb=rgb, vertices=vertices, faces=faces, f_normals=f_normals,
                xatlas_dict=quad_atlas(vertices, faces, atlas_res))
