"""Oracle (TEST INFRASTRUCTURE): completion of never-seen texels from mesh neighbours
("next" row N2).

Restates
  * pointdreamer/unproject.py:93-196   paint_invisible_areas_by_neighbors (use_atlas=True)
  * pointdreamer/unproject.py:17-38    compute_vertex_only_uv_mask
  * utils/mesh_utils.py:7-114          subdivide_with_uv
in numpy.  Third-party pieces that are not vendored in the reference (PARITY UNPINNED, canonical
rules stated here):
  * trimesh.geometry.faces_to_edges / trimesh.grouping.unique_rows (trimesh is unpinned in
    requirements.txt): edges = faces[:, [0,1,1,2,2,0]]; unique rows are returned in the order of
    their packed 64-bit hash = (col1 << 32) ^ col0 after an offset, i.e. sorted by (col1, col0),
    each represented by its first occurrence;
  * kaolin 0.15.0 ops.mesh.uniform_laplacian: binary vertex adjacency divided by the vertex
    degree in fp32, diagonal -1 (0 after the reference adds the identity);
  * the dense fp32 matmul of the colouring loop: canonical summation order = ascending
    neighbour index, each term (1/deg_i) * (colour_j * count_j);
  * index_put with duplicate indices: the highest source index wins (the deterministic rule,
    SURVEY §8a P5);  scipy griddata 'nearest' ties: oracle/fill.py.
Pinned against the reference's own function run under oracle/ref_loader.py
(tests/golden/make_golden_neighbors.py -> neighbors_small.npz).
"""
import numpy as np

from . import fill as ofill

F32 = np.float32


def faces_to_edges(faces):
    return np.asarray(faces)[:, [0, 1, 1, 2, 2, 0]].reshape((-1, 2))


def unique_rows(data):
    """trimesh.grouping.unique_rows for two integer columns -> (unique, inverse)."""
    d = np.asarray(data, dtype=np.int64)
    if len(d) == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    threshold = 2 ** 31
    bit = (d.T + (threshold + 1)).astype(np.uint64)
    h = bit[0] ^ (bit[1] << np.uint64(32))
    _, unique, inverse = np.unique(h, return_index=True, return_inverse=True)
    return unique, inverse.reshape(-1)


def _split(tri, attr):
    """one midpoint per unique edge of `tri`; 4 children per triangle (corner, corner, corner,
    centre) in trimesh's winding."""
    e = np.sort(faces_to_edges(tri), axis=1)
    uq, inv = unique_rows(e)
    mid = attr[e[uq]].mean(axis=1) if len(uq) else np.zeros((0, attr.shape[1]), dtype=attr.dtype)
    m = inv.reshape((-1, 3)) + len(attr)
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    m0, m1, m2 = m[:, 0], m[:, 1], m[:, 2]
    new = np.stack([a, m0, m2, m0, b, m1, m2, m1, c, m0, m1, m2], 1).reshape((-1, 3))
    return new, mid


def subdivide_with_uv(vertices, faces, face_uv_idx, uvs, face_index=None):
    """utils/mesh_utils.py:7-114 -> (new_vertices, new_faces, new_uvs, new_face_uv_idx)."""
    if face_index is None:
        fm = np.ones(len(faces), dtype=bool)
    else:
        fm = np.zeros(len(faces), dtype=bool)
        fm[face_index] = True
    f, mid = _split(faces[fm], vertices)
    f_uv, mid_uv = _split(face_uv_idx[fm], uvs)
    return (np.vstack((vertices, mid)), np.vstack((faces[~fm], f)), np.vstack((uvs, mid_uv)),
            np.vstack((face_uv_idx[~fm], f_uv)))


def vertex_pixels(faces, face_uv_idx, uvs, n_vertices, atlas_res):
    """unproject.py:118-131: one uv per vertex (for a vertex on a seam the pair that sorts last,
    i.e. its largest uv index, wins the duplicate index_put), then (row, col) atlas pixel."""
    pairs = np.unique(np.stack((faces.reshape(-1), face_uv_idx.reshape(-1)), 1), axis=0)
    vert_uvs = np.zeros((n_vertices, 2), dtype=F32)
    vert_uvs[pairs[:, 0]] = uvs[pairs[:, 1]]  # numpy: last assignment wins, pairs are sorted
    px = np.clip(vert_uvs * F32(atlas_res), 0, atlas_res - 1).astype(np.int64)
    return np.stack((px[:, 1], px[:, 0]), 1)


def adjacency_csr(n_vertices, faces):
    """kaolin adjacency_matrix: unique directed edges of every face, both directions."""
    f = np.asarray(faces, dtype=np.int64)
    r = np.roll(f, 1, axis=-1)
    ind = np.concatenate([np.stack([f, r], -1), np.stack([r, f], -1)], 1).reshape(-1, 2)
    ind = np.unique(ind, axis=0)
    rowptr = np.zeros(n_vertices + 1, dtype=np.int64)
    np.add.at(rowptr, ind[:, 0] + 1, 1)
    return np.cumsum(rowptr), ind[:, 1]


def colour_by_neighbours(colors, has_color, rowptr, colidx, max_rounds=10000):
    """unproject.py:137-172: Jacobi rounds of neighbour averaging over the never-coloured
    vertices until no new vertex gets a colour, then as many smoothing rounds again."""
    colors = np.array(colors, dtype=F32)
    V = colors.shape[0]
    invalid = np.nonzero(~has_color)[0]
    deg = (rowptr[1:] - rowptr[:-1])
    with np.errstate(divide="ignore"):
        w = (F32(1) / deg.astype(F32)).astype(F32)
    w[deg == 0] = 0  # kaolin: NaN rows (0/0) are zeroed
    count = np.ones(V, dtype=F32)
    count[invalid] = 0
    maxdeg = int(deg[invalid].max()) if len(invalid) else 0
    nb = np.full((len(invalid), maxdeg), -1, dtype=np.int64)
    for k in range(maxdeg):
        has = deg[invalid] > k
        nb[has, k] = colidx[rowptr[invalid][has] + k]
    wi = w[invalid]
    total = count.sum()
    coloring_round = 0
    stage = "uncolored"
    rounds = 0
    while stage == "uncolored" or coloring_round > 0:
        nc = np.zeros((len(invalid), 3), dtype=F32)
        nn = np.zeros(len(invalid), dtype=F32)
        for k in range(maxdeg):
            j = nb[:, k]
            ok = j >= 0
            jj = np.where(ok, j, 0)
            cj = colors[jj] * count[jj][:, None]
            nc = np.where(ok[:, None], nc + wi[:, None] * cj, nc).astype(F32)
            nn = np.where(ok, nn + wi * count[jj], nn).astype(F32)
        pos = nn > 0
        with np.errstate(divide="ignore", invalid="ignore"):
            avg = (nc / nn[:, None]).astype(F32)
        colors[invalid] = np.where(pos[:, None], avg, colors[invalid])
        count[invalid] = pos.astype(F32)
        new_total = count.sum()
        rounds += 1
        if new_total > total:
            total = new_total
            coloring_round += 1
        else:
            stage = "colored"
            coloring_round -= 1
        if coloring_round > max_rounds:
            break
    assert not np.isnan(colors).any()
    return colors, rounds


def paint_invisible_areas_by_neighbors(vertices, faces, uvs, face_uv_idx, to_inpaint_face_id,
                                       atlas_img, atlas_inpainted_mask):
    """unproject.py:93-196 (use_atlas=True).  atlas_img [R,R,3] f32, atlas_inpainted_mask [R,R]
    bool.  Returns (atlas [R,R,3] f32, tie mask of the final nearest fill, rounds)."""
    R = atlas_inpainted_mask.shape[1]
    v = np.asarray(vertices, dtype=F32)
    f = np.asarray(faces, dtype=np.int64)
    uv = np.asarray(uvs, dtype=F32)
    fuv = np.asarray(face_uv_idx, dtype=np.int64)
    fid = np.asarray(to_inpaint_face_id, dtype=np.int64)
    for _ in range(2):  # the SAME face ids are reused on the renumbered faces (unproject.py:112-114)
        v, f, uv, fuv = subdivide_with_uv(v, f, fuv, uv, face_index=fid)
    pix = vertex_pixels(f, fuv, uv, len(v), R)
    atlas = np.array(atlas_img, dtype=F32)
    mask = np.array(atlas_inpainted_mask, dtype=bool)
    colors = atlas[pix[:, 0], pix[:, 1]]
    has = mask[pix[:, 0], pix[:, 1]]
    rowptr, colidx = adjacency_csr(len(v), f)
    colors, rounds = colour_by_neighbours(colors, has, rowptr, colidx)
    atlas[pix[:, 0], pix[:, 1]] = colors  # numpy: the highest vertex index wins a shared texel
    mask[pix[:, 0], pix[:, 1]] = True
    out, tie = ofill.nearest_fill(np.ascontiguousarray(atlas.transpose(2, 0, 1)), mask)
    return out.transpose(1, 2, 0).copy(), tie, rounds
