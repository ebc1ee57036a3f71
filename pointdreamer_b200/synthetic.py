"""Deterministic synthetic inputs for the path.

There is no network for datasets and the reference ships no meshes/atlases, so benchmarks and
parity tests use an analytic closed surface (a bumpy torus):

  * coloured cloud: N points sampled on the surface, procedural RGB quantised to uint8/255
    (the PLY wire format is uchar, demo.py:376), normalised exactly like demo.py:377-380;
  * untextured mesh: the same surface tessellated on an nu x nv parameter grid
    (2*nu*nv = 10368 triangles, ~10k like models/POCO/generate_1.py:49);
  * UV atlas: gx x gy charts with gutters; `gb_pos`, `mask`, `per_atlas_pixel_face_id`,
    `uvs`, `mesh_tex_idx` in the layout of models/get3d/extract_texture_map.py:42-64.

Pure numpy, host side: this only produces INPUTS, it is not part of the hot path.
"""
import numpy as np


def _surface(u, v, R0=0.62, r0=0.26, bump=0.06):
    """bumpy torus; u, v in [0,1) (periodic).  Returns xyz float64."""
    a = 2 * np.pi * u
    b = 2 * np.pi * v
    r = r0 * (1.0 + bump * np.cos(3 * a) * np.cos(2 * b))
    x = (R0 + r * np.cos(b)) * np.cos(a)
    y = r * np.sin(b) * 1.3
    z = (R0 + r * np.cos(b)) * np.sin(a)
    return np.stack([x, y, z], -1)


def _color(u, v):
    """smooth + high-frequency procedural RGB in [0,1]."""
    a = 2 * np.pi * u
    b = 2 * np.pi * v
    r = 0.5 + 0.35 * np.sin(a) + 0.15 * np.sin(17 * a + 5 * b)
    g = 0.5 + 0.35 * np.cos(b) + 0.15 * np.sin(23 * b - 3 * a)
    bl = 0.5 + 0.3 * np.sin(a + b) + 0.2 * ((np.floor(u * 16) + np.floor(v * 12)) % 2 - 0.5)
    return np.clip(np.stack([r, g, bl], -1), 0, 1)


def make_cloud(n_points=30000, seed=0, noise_std=0.0):
    """Returns xyz [N,3] fp32 (normalised as demo.py:377-380), rgb [N,3] fp32 in {k/255},
    and the (centre, scale) used so the mesh can be normalised identically."""
    rng = np.random.default_rng(seed)
    # rejection sampling for (approximately) area-uniform samples on the torus
    u = np.empty(0)
    v = np.empty(0)
    while u.shape[0] < n_points:
        uu = rng.random(2 * n_points)
        vv = rng.random(2 * n_points)
        keep = rng.random(2 * n_points) < (0.62 + 0.26 * np.cos(2 * np.pi * vv)) / 0.88
        u = np.concatenate([u, uu[keep]])
        v = np.concatenate([v, vv[keep]])
    u, v = u[:n_points], v[:n_points]
    xyz = _surface(u, v)
    if noise_std > 0:
        xyz = xyz + rng.normal(0, noise_std, xyz.shape)
    rgb8 = np.floor(_color(u, v) * 255.0 + 0.5).astype(np.uint8)
    xyz = xyz.astype(np.float32)
    vmin = xyz.min(0)
    vmax = xyz.max(0)
    centre = (vmax + vmin) / np.float32(2.0)
    scale = (vmax - vmin).max()
    xyz = (xyz - centre) / scale
    return xyz.astype(np.float32), (rgb8.astype(np.float32) / np.float32(255.0)), (centre, scale)


def make_mesh(nu=72, nv=72, norm=None):
    """Tessellate the surface: vertices [nu*nv,3] fp32, faces [2*nu*nv,3] int64 (periodic grid),
    f_normals [F,3] fp32 (normalised cross product, the `kal.ops.mesh.face_normals` call at
    demo.py:422)."""
    uu, vv = np.meshgrid(np.arange(nu) / nu, np.arange(nv) / nv, indexing="ij")
    verts = _surface(uu.reshape(-1), vv.reshape(-1)).astype(np.float32)
    if norm is not None:
        centre, scale = norm
        verts = ((verts - centre) / scale).astype(np.float32)
    idx = lambda i, j: (i % nu) * nv + (j % nv)
    ii, jj = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    ii = ii.reshape(-1)
    jj = jj.reshape(-1)
    v00, v10, v01, v11 = idx(ii, jj), idx(ii + 1, jj), idx(ii, jj + 1), idx(ii + 1, jj + 1)
    # quad q -> faces 2q (v00,v10,v11) and 2q+1 (v00,v11,v01)
    faces = np.empty((2 * nu * nv, 3), dtype=np.int64)
    faces[0::2] = np.stack([v00, v10, v11], 1)
    faces[1::2] = np.stack([v00, v11, v01], 1)
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    n = np.cross((b - a).astype(np.float64), (c - a).astype(np.float64))
    n = n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20)
    return verts, faces, n.astype(np.float32)


def make_atlas(vertices, faces, nu=72, nv=72, res=1024, gx=4, gy=4, fill=0.92):
    """Chart atlas for the parameter-grid mesh of `make_mesh`.

    The (u,v) domain is cut into gx x gy charts; chart (cx,cy) occupies the centred `fill`
    fraction of atlas cell (cx,cy) so neighbouring charts are separated by gutters.
    Returns the xatlas dict of demo.py:441-448:
      uvs [F*3,2] fp32 (per face-corner), mesh_tex_idx [F,3] int64,
      gb_pos [1,res,res,3] fp32, mask [1,res,res,1] bool, per_atlas_pixel_face_id [1,res,res] int64
    """
    F = faces.shape[0]
    ys, xs = np.meshgrid(np.arange(res), np.arange(res), indexing="ij")
    ax = (xs + 0.5) / res  # atlas coords of texel centres
    ay = (ys + 0.5) / res
    cx = np.minimum((ax * gx).astype(np.int64), gx - 1)
    cy = np.minimum((ay * gy).astype(np.int64), gy - 1)
    lx = (ax * gx - cx - (1 - fill) / 2) / fill  # local chart coords in [0,1]
    ly = (ay * gy - cy - (1 - fill) / 2) / fill
    inside = (lx >= 0) & (lx < 1) & (ly >= 0) & (ly < 1)
    u = (cx + lx) / gx
    v = (cy + ly) / gy
    fu = np.clip(u * nu, 0, nu - 1e-9)
    fv = np.clip(v * nv, 0, nv - 1e-9)
    qi = np.floor(fu).astype(np.int64)
    qj = np.floor(fv).astype(np.int64)
    s = fu - qi
    t = fv - qj
    quad = qi * nv + qj
    upper = t > s  # triangle 2q+1 = (v00, v11, v01) holds the t > s half
    face = 2 * quad + upper.astype(np.int64)
    # barycentric weights
    w0 = np.where(upper, 1 - t, 1 - s)
    w1 = np.where(upper, s, s - t)
    w2 = np.where(upper, t - s, t)
    fa = faces[face.reshape(-1)]
    V = vertices.astype(np.float64)
    pos = (w0.reshape(-1, 1) * V[fa[:, 0]] + w1.reshape(-1, 1) * V[fa[:, 1]]
           + w2.reshape(-1, 1) * V[fa[:, 2]])
    gb_pos = np.where(inside.reshape(-1, 1), pos, 0.0).astype(np.float32).reshape(1, res, res, 3)
    face_id = np.where(inside, face, -1).astype(np.int64).reshape(1, res, res)
    mask = inside.reshape(1, res, res, 1)

    # per-face-corner uvs (each face gets its own 3 uv entries, like xatlas output)
    ii, jj = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    ii = ii.reshape(-1).astype(np.float64)
    jj = jj.reshape(-1).astype(np.float64)

    def to_atlas(gu, gv, qi_, qj_):
        # chart of the quad (by its lower corner), then atlas position of grid point (gu,gv)
        uu, vv = gu / nu, gv / nv
        ccx = np.minimum((qi_ / nu * gx).astype(np.int64), gx - 1)
        ccy = np.minimum((qj_ / nv * gy).astype(np.int64), gy - 1)
        lx_ = uu * gx - ccx
        ly_ = vv * gy - ccy
        return np.stack([(ccx + (1 - fill) / 2 + lx_ * fill) / gx,
                         (ccy + (1 - fill) / 2 + ly_ * fill) / gy], -1)

    c00 = to_atlas(ii, jj, ii, jj)
    c10 = to_atlas(ii + 1, jj, ii, jj)
    c01 = to_atlas(ii, jj + 1, ii, jj)
    c11 = to_atlas(ii + 1, jj + 1, ii, jj)
    uvs = np.empty((F, 3, 2), dtype=np.float64)
    uvs[0::2] = np.stack([c00, c10, c11], 1)
    uvs[1::2] = np.stack([c00, c11, c01], 1)
    uvs = uvs.reshape(F * 3, 2).astype(np.float32)
    mesh_tex_idx = np.arange(F * 3, dtype=np.int64).reshape(F, 3)
    return dict(uvs=uvs, mesh_tex_idx=mesh_tex_idx, gb_pos=gb_pos, mask=mask,
                per_atlas_pixel_face_id=face_id)


def make_scene(n_points=30000, seed=0, nu=72, nv=72, atlas_res=1024, noise_std=0.0,
               charts=(4, 4)):
    """Everything `colorize_one_mesh` needs besides cameras/config, as numpy arrays."""
    xyz, rgb, norm = make_cloud(n_points, seed, noise_std)
    vertices, faces, f_normals = make_mesh(nu, nv, norm)
    atlas = make_atlas(vertices, faces, nu, nv, atlas_res, charts[0], charts[1])
    return dict(xyz=xyz, rgb=rgb, vertices=vertices, faces=faces, f_normals=f_normals,
                xatlas_dict=atlas)


# ------------------------------------------------------------------------------------------
# Proxy mesh + atlas for a REAL point cloud (BASELINE configs[0], [2]).
# The reference ships clouds but no meshes/atlases (POCO weights and xatlas are unavailable), so
# runs on dataset/demo_data/*.ply use a voxel-shell mesh built from the cloud itself: occupied
# voxels of a G^3 grid, one quad (2 triangles) per face between an occupied and an empty voxel,
# and an atlas with one c x c cell per quad.  It only has to be deterministic and identical for
# the reference run and for ours.
# ------------------------------------------------------------------------------------------
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


def proxy_scene_from_ply(ply_path, G=16, atlas_res=512):
    """Scene dict (cloud, proxy mesh, atlas) for a PLY cloud, normalised as demo.py:376-380."""
    from .io_utils import normalize_cloud, read_ply_xyzrgb
    xyz, rgb8 = read_ply_xyzrgb(ply_path)
    xyz = normalize_cloud(xyz)
    rgb = rgb8.astype(np.float32) / np.float32(255.0)
    vertices, faces, f_normals = voxel_shell(xyz, G)
    return dict(xyz=xyz, rgb=rgb, vertices=vertices, faces=faces, f_normals=f_normals,
                xatlas_dict=quad_atlas(vertices, faces, atlas_res))
