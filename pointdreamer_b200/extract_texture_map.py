"""Atlas inputs ("next" row N4): `xatlas_uvmap_w_face_id` with the reference's signature
(models/get3d/extract_texture_map.py:42-64).  The UV-space rasterisation and the world-position
interpolation run in libpdr.so; `xatlas.parametrize` itself stays third-party: it is imported if
installed, otherwise the caller supplies its result through `parametrization=`.
"""
import numpy as np
import torch

from . import ours_utils as _ou


def xatlas_uvmap_w_face_id(ctx, mesh_v, mesh_pos_idx, resolution, parametrization=None):
    """Returns (uvs [Nu,2] f32, mesh_tex_idx [F,3] int64, gb_pos [1,R,R,3] f32,
    mask [1,R,R,1] bool, per_pixel_face_idx [1,R,R] int64).  `ctx` is ignored."""
    dev = mesh_v.device
    if parametrization is None:
        try:
            import xatlas
        except ImportError as e:
            raise RuntimeError("xatlas is not installed: pass parametrization=(vmapping, indices, "
                               "uvs), the result of xatlas.parametrize") from e
        parametrization = xatlas.parametrize(mesh_v.detach().cpu().numpy(),
                                             mesh_pos_idx.detach().cpu().numpy())
    _, indices, uvs = parametrization
    indices_int64 = np.asarray(indices).astype(np.uint64, casting='same_kind').view(np.int64)
    uvs = torch.as_tensor(np.asarray(uvs), dtype=torch.float32, device=dev)
    mesh_tex_idx = torch.as_tensor(indices_int64, dtype=torch.int64, device=dev)
    # extract_texture_map.py:50-54: clip-space position (uv*2-1, 0, 1)
    uv_clip = uvs[None, ...] * 2.0 - 1.0
    pos = torch.cat((uv_clip, torch.zeros_like(uv_clip[..., 0:1]),
                     torch.ones_like(uv_clip[..., 0:1])), dim=-1).contiguous()
    mask, face_idx, _, _ = _ou.rasterize(pos, mesh_tex_idx, resolution, resolution)
    gb_pos = _ou.interpolate(mesh_v, pos, mesh_tex_idx, face_idx, mesh_pos_idx)
    return uvs, mesh_tex_idx, gb_pos, mask[..., None], face_idx
