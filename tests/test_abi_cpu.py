"""CPU-side checks of the boundary: libpdr.so builds, loads without a GPU and exports every symbol
declared in include/pdr.h; argument validation works without touching a device; host-side helpers
agree with the oracle."""
import ctypes

import numpy as np
import pytest

from pointdreamer_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    from pointdreamer_b200 import build
    build.build()
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    names = _lib.declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.pdr_version() == 100


def test_argument_errors_do_not_need_a_gpu(lib):
    rc = lib.pdr_conv_tc(None, None, None, None, None, None, 1, 8, 8, 64, 0, 128, 9, 0, None)
    assert rc < 0 and b"null" in lib.pdr_last_error()
    rc = lib.pdr_unproject(*([None] * 3), 1, 1, *([None] * 4), 1, None, 1, None, None,
                           ctypes.c_double(0.0), 0, None, None, None, 1, 0, *([None] * 8))
    assert rc < 0


def test_unet_engine_plans_without_gpu(lib):
    """The static planner is pure host code: the full 256x256 model needs a 1.9 GB arena at B=8."""
    from pointdreamer_b200.unet import DEFAULT_MODEL_CONFIG, PdrUnetConfig, param_shapes
    cfg = DEFAULT_MODEL_CONFIG
    c = PdrUnetConfig()
    c.image_size, c.in_channels, c.model_channels = 256, 3, 256
    c.out_channels, c.num_res_blocks, c.n_mult = 6, 2, 6
    for i, m in enumerate(cfg["channel_mult"]):
        c.channel_mult_x2[i] = 2 * m
    c.n_attn_ds = 3
    for i, d in enumerate([8, 16, 32]):
        c.attn_ds[i] = d
    c.num_head_channels = 64
    h = ctypes.c_void_p()
    assert lib.pdr_unet_create(ctypes.byref(c), ctypes.byref(h)) == 0
    need = ctypes.c_size_t(0)
    # missing parameters are reported by name
    assert lib.pdr_unet_workspace_bytes(h, 8, ctypes.byref(need)) < 0
    assert b"was not provided" in lib.pdr_last_error()
    # register dummy (non-null) pointers with the right byte sizes
    shapes = param_shapes(cfg)
    emb_rows = 0
    for name, shp in shapes.items():
        n = int(np.prod(shp))
        if "emb_layers" in name:
            if name.endswith(".bias"):
                emb_rows += n
            continue
        torso_conv = len(shp) in (3, 4) and not name.startswith("out.")
        nbytes = n * (2 if (torso_conv and name.endswith(".weight")) else 4)
        assert lib.pdr_unet_set_param(h, name.encode(), ctypes.c_void_p(1024), ctypes.c_size_t(nbytes)) == 0
    # channel-changing ResBlocks: out_layers.3 and skip_connection as ONE weight matrix along K
    for name, shp in shapes.items():
        if name.endswith(".skip_connection.weight"):
            p = name[:-len(".skip_connection.weight")]
            cout, cin = shp[0], shp[1]
            lib.pdr_unet_set_param(h, (p + ".out_layers.3_skip.weight").encode(), ctypes.c_void_p(1024),
                                   ctypes.c_size_t(cout * (9 * cout + cin) * 2))
            lib.pdr_unet_set_param(h, (p + ".out_layers.3_skip.bias").encode(), ctypes.c_void_p(1024),
                                   ctypes.c_size_t(cout * 4))
    lib.pdr_unet_set_param(h, b"emb_all.weight", ctypes.c_void_p(1024), ctypes.c_size_t(emb_rows * 1024 * 4))
    lib.pdr_unet_set_param(h, b"emb_all.bias", ctypes.c_void_p(1024), ctypes.c_size_t(emb_rows * 4))
    rc = lib.pdr_unet_workspace_bytes(h, 8, ctypes.byref(need))
    assert rc == 0, lib.pdr_last_error()
    assert 1.0e9 < need.value < 4.0e9
    lib.pdr_unet_destroy(h)


def test_step_table_matches_oracle():
    from oracle import ddnm as oddnm
    from pointdreamer_b200.ddnm_inpainting import DEFAULT_DDNM_CONFIG, step_table
    ts, c = step_table(DEFAULT_DDNM_CONFIG)
    ts2, c2 = oddnm.step_table()
    assert np.array_equal(ts, ts2) and np.array_equal(c, c2)
    assert ts[0] == 990 and ts[-1] == 0 and len(ts) == 100
    assert c[-1, 3] == 0 and c[-1, 2] == 1  # last step: gamma = 0, sqrt(alpha_next) = 1


def test_param_shapes_match_reference_count():
    from pointdreamer_b200.unet import param_shapes
    total = sum(int(np.prod(s)) for s in param_shapes().values())
    assert total == 552814086  # reference create_model(**imagenet_256.yml)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PdrError):
        _lib.load()


def test_texture_psnr_metric():
    import numpy as np
    from pointdreamer_b200 import metrics
    a = np.zeros((8, 8, 3), dtype=np.uint8)
    b = a.copy()
    assert metrics.calculate_psnr(a, b) == float("inf")
    b[...] = 1
    assert abs(metrics.calculate_psnr(a, b) - 20 * np.log10(255.0)) < 1e-9
    atlas = np.linspace(-0.1, 1.1, 8 * 8 * 3, dtype=np.float32).reshape(8, 8, 3)
    q = metrics.atlas_to_uint8(atlas)
    assert q.dtype == np.uint8 and q[-1, 0, 0] == 0 and q[0, -1, -1] == 255


def test_workspace_queries_are_host_functions(lib):
    """The *_workspace_bytes entry points are pure host code: callable without a GPU, growing with the
    problem, and large enough for the arrays pdr.h says they carve (the hidden-point-removal one holds
    three [V, N] double4 arrays - points, sorted survivors - two [V, N] double2 arrays and the sort
    buckets)."""
    for name in ("pdr_hidden_point_removal_workspace_bytes", "pdr_rasterize_workspace_bytes",
                 "pdr_sparse_images_workspace_bytes", "pdr_nearest_fill_workspace_bytes",
                 "pdr_unproject_workspace_bytes"):
        getattr(lib, name).restype = ctypes.c_size_t
    V, N = 8, 30000
    hpr = lib.pdr_hidden_point_removal_workspace_bytes(V, N)
    assert hpr >= V * N * (2 * 32 + 2 * 16 + 4 + 2 + 2)
    assert hpr < 64 << 20  # a production call stays a small slice of HBM
    assert lib.pdr_hidden_point_removal_workspace_bytes(V, 2 * N) > hpr
    assert lib.pdr_hidden_point_removal_workspace_bytes(2 * V, N) > hpr
    assert lib.pdr_hidden_point_removal_workspace_bytes(1, 1) % 256 == 0
    r = lib.pdr_rasterize_workspace_bytes(8, 20000, 512)
    assert r > 0 and lib.pdr_rasterize_workspace_bytes(8, 40000, 512) > r
    assert lib.pdr_nearest_fill_workspace_bytes(1, 1024, 1024) >= 1024 * 1024 * 2
    assert lib.pdr_sparse_images_workspace_bytes(8, 256) > 0
    assert lib.pdr_unproject_workspace_bytes(1024, 1) > 0
