"""Mesh subdivision with UVs — same name/arguments/returns as the reference's
utils/mesh_utils.py:7-114 `subdivide_with_uv` (a modified trimesh.remesh.subdivide), on device
tensors.  This is index bookkeeping on a few thousand faces (setup for the neighbour completion,
"next" row N2); it uses torch's sort/unique and no arithmetic besides the edge midpoints.

Numbering contract (it decides which vertex wins a shared texel later on): one midpoint per unique
undirected edge of the subdivided faces, appended after the existing vertices in the order of
trimesh.grouping.unique_rows, i.e. sorted by (larger endpoint, smaller endpoint).
"""
import torch


def _split(tri, attr):
    """tri [n,3] int64, attr [m,d] -> (4n children in trimesh's winding, midpoints)."""
    n = tri.shape[0]
    e = tri[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)
    lo = torch.minimum(e[:, 0], e[:, 1])
    hi = torch.maximum(e[:, 0], e[:, 1])
    key, inv = torch.unique((hi << 32) | lo, sorted=True, return_inverse=True)
    mid = (attr[key & 0xFFFFFFFF] + attr[key >> 32]) / 2
    m = inv.reshape(n, 3) + attr.shape[0]
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    m0, m1, m2 = m[:, 0], m[:, 1], m[:, 2]
    new = torch.stack([a, m0, m2, m0, b, m1, m2, m1, c, m0, m1, m2], 1).reshape(-1, 3)
    return new, mid


def subdivide_with_uv(vertices, faces, face_uv_idx, uvs, face_index=None):
    """Returns (new_vertices, new_faces, new_uvs, new_face_uv_idx).  Only the faces in
    `face_index` are split; the untouched faces come first in the new face list."""
    F = faces.shape[0]
    if face_index is None:
        fm = torch.ones(F, dtype=torch.bool, device=faces.device)
    else:
        fm = torch.zeros(F, dtype=torch.bool, device=faces.device)
        fm[face_index] = True
    f, mid = _split(faces[fm], vertices)
    f_uv, mid_uv = _split(face_uv_idx[fm], uvs)
    return (torch.cat((vertices, mid)), torch.cat((faces[~fm], f)), torch.cat((uvs, mid_uv)),
            torch.cat((face_uv_idx[~fm], f_uv)))
