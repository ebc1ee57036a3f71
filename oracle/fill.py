"""Oracle (TEST INFRASTRUCTURE): nearest-valid-pixel fill.

Restates ours_utils.py:610-643 `naive_inpainting(method='nearest')` (scipy griddata ->
cKDTree 1-NN of every pixel over the valid pixels; valid pixels map to themselves) and
unproject.py:480-504 `dilate_atlas`.

scipy's tie rule between equidistant sources is the kd-tree traversal order (UNPINNED).
Canonical rule used by oracle and CUDA: minimum squared distance, then the lowest linear
index (row*W + col) of the source pixel.  `tie_mask` reports where more than one source is
at the minimum distance so tests can exclude those pixels when comparing with scipy.
"""
import numpy as np


def _column_nearest(valid):
    """per pixel: distance to nearest valid pixel above-or-at (up) and at-or-below (dn) in the
    same column; a large sentinel when none."""
    H, W = valid.shape
    BIG = 1 << 20
    up = np.full((H, W), BIG, dtype=np.int64)
    dn = np.full((H, W), BIG, dtype=np.int64)
    last = np.full(W, -BIG, dtype=np.int64)
    for y in range(H):
        last = np.where(valid[y], y, last)
        up[y] = np.minimum(y - last, BIG)
    nxt = np.full(W, 3 * BIG, dtype=np.int64)
    for y in range(H - 1, -1, -1):
        nxt = np.where(valid[y], y, nxt)
        dn[y] = np.minimum(nxt - y, BIG)
    return up, dn


def nearest_source(valid):
    """For every pixel the (row, col) of its nearest valid pixel under the canonical rule.
    Returns src_row[H,W], src_col[H,W] (int64; -1 where `valid` is empty) and tie_mask[H,W]."""
    H, W = valid.shape
    BIG = 1 << 20
    up, dn = _column_nearest(valid)
    best_d = np.full((H, W), np.iinfo(np.int64).max, dtype=np.int64)
    best_r = -np.ones((H, W), dtype=np.int64)
    best_c = -np.ones((H, W), dtype=np.int64)
    ties = np.zeros((H, W), dtype=np.int64)
    rows = np.arange(H)[:, None]
    cols = np.arange(W)[None, :]

    def consider(d2, r, c, ok):
        nonlocal best_d, best_r, best_c, ties
        lin = r * W + c
        blin = best_r * W + best_c
        better = ok & ((d2 < best_d) | ((d2 == best_d) & (lin < blin)))
        same = ok & (d2 == best_d) & (lin != blin)
        ties = np.where(ok & (d2 < best_d), 0, ties)
        ties = ties + same.astype(np.int64)
        best_d = np.where(better, d2, best_d)
        best_r = np.where(better, r, best_r)
        best_c = np.where(better, c, best_c)

    order = [0]
    for a in range(1, W):
        order += [-a, a]
    for dx in order:
        if dx * dx > best_d.max():
            break  # every remaining column is farther than the current best everywhere
        c = cols + dx
        inb = (c >= 0) & (c < W)
        cc = np.clip(c, 0, W - 1)
        u = np.take_along_axis(up, np.broadcast_to(cc, (H, W)), 1)
        d = np.take_along_axis(dn, np.broadcast_to(cc, (H, W)), 1)
        cfull = np.broadcast_to(cc, (H, W))
        consider(dx * dx + u * u, rows - u, cfull, inb & (u < BIG))
        # the below candidate coincides with the above one when the pixel itself is valid (d == 0)
        consider(dx * dx + d * d, rows + d, cfull, inb & (d < BIG) & (d > 0))
    return best_r, best_c, ties > 0


def nearest_fill(img, known):
    """img [C,H,W] fp32, known [H,W] bool -> filled [C,H,W] (exact copies of source pixels)."""
    r, c, tie = nearest_source(known)
    if (r < 0).any():
        raise ValueError("nearest_fill: no valid pixel")
    return img[:, r, c].copy(), tie


def naive_inpainting_nearest(img, mask2):
    """ours_utils.py:610-643 with method='nearest': uses channel 0 of the mask."""
    known = mask2[0].astype(bool)
    return nearest_fill(img, known)


def dilate_atlas(atlas, mask):
    """unproject.py:480-504: atlas [R,R,3], mask [1,R,R,1] -> atlas [R,R,3]."""
    known = mask[0, :, :, 0].astype(bool)
    out, tie = nearest_fill(np.ascontiguousarray(atlas.transpose(2, 0, 1)), known)
    return out.transpose(1, 2, 0).copy(), tie
