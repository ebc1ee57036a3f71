"""Host-side formats of the file-level flow ("next" row N3) that need no GPU: PLY cloud, OBJ/MTL."""
import numpy as np

from pointdreamer_b200 import io_utils


def test_ply_round_trip_and_wire_format(tmp_path):
    rng = np.random.default_rng(0)
    xyz = rng.normal(size=(100, 3)).astype(np.float32)
    rgb8 = rng.integers(0, 256, size=(100, 3), dtype=np.uint8)
    p = str(tmp_path / "c.ply")
    io_utils.save_colored_pc_ply(xyz, rgb8.astype(np.float32) / 255.0, p)
    raw = open(p, "rb").read()
    header, body = raw.split(b"end_header\n", 1)
    assert header.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 100\n")
    assert len(body) == 100 * 15  # 3 x float32 + 3 x uchar per vertex (utils/other_utils.py:155-162)
    x2, c2 = io_utils.read_ply_xyzrgb(p)
    assert np.array_equal(x2, xyz) and np.array_equal(c2, rgb8)


def test_normalize_cloud_matches_demo_py():
    xyz = np.array([[0, 0, 0], [2, 1, 4], [1, 3, 2]], dtype=np.float32)
    n = io_utils.normalize_cloud(xyz.copy())
    assert np.allclose(n.max(0) + n.min(0), 0, atol=1e-6)       # centred on the bbox centre
    assert np.isclose((n.max(0) - n.min(0)).max(), 1.0)          # largest extent == 1


def test_obj_mtl_text_and_reader(tmp_path):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float32)
    vt = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [0.5, 0.5]], dtype=np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3]])
    ft = np.array([[0, 1, 2], [0, 2, 4]])
    obj = str(tmp_path / "models" / "model_normalized.obj")
    (tmp_path / "models").mkdir()
    io_utils.savemeshtes2(v, vt, f, ft, obj)
    text = open(obj).read().splitlines()
    assert text[0] == "mtllib model_normalized.mtl"
    assert text[1] == "v 0.000000 0.000000 0.000000" and text[5] == "vt 0.000000 0.000000"
    assert text[10] == "usemtl material_0" and text[11] == "f 1/1 2/2 3/3" and text[12] == "f 1/1 3/3 4/5"
    mtl = open(str(tmp_path / "models" / "model_normalized.mtl")).read()
    assert mtl == ("newmtl material_0\nKd 1 1 1\nKa 0 0 0\nKs 0.4 0.4 0.4\nNs 10\nillum 2\n"
                   "map_Kd model_normalized.png\n")
    v2, vt2, f2, ft2 = io_utils.loadobjtex(obj)
    assert np.array_equal(f2, f) and np.array_equal(ft2, ft)
    assert np.allclose(v2, v) and np.allclose(vt2, vt)


def test_obj_reader_splits_quads_and_handles_missing_uvs(tmp_path):
    p = str(tmp_path / "q.obj")
    open(p, "w").write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n")
    v, vt, f, ft = io_utils.loadobjtex(p)
    assert vt is None and ft is None
    assert f.tolist() == [[0, 1, 2], [0, 2, 3]]


def test_load_config_and_keys(tmp_path):
    from pointdreamer_b200 import demo
    y = tmp_path / "c.yaml"
    y.write_text("view_num: 4\nres: 128\ntexture_gen_method: nearest\n")
    cfg = dict(demo.DEFAULT_CONFIG)
    cfg.update(demo.load_config(str(y)))
    assert cfg["view_num"] == 4 and cfg["texture_gen_method"] == "nearest"
    assert all(k in cfg for k in demo.PATH_CONFIG_KEYS)


def test_ply_reader_accepts_what_plyfile_accepts(tmp_path):
    """CRLF header, an element before the vertices, a face element with a list property, extra
    vertex properties, big-endian and ASCII bodies (the reference reads clouds with plyfile)."""
    import struct
    xyz = np.array([[0.5, -1.25, 2.0], [3.0, 4.0, -5.5]], dtype=np.float32)
    rgb = np.array([[1, 2, 3], [250, 128, 0]], dtype=np.uint8)

    def write(path, fmt, nl):
        end = ">" if fmt == "binary_big_endian" else "<"
        hdr = ["ply", f"format {fmt} 1.0", "comment made by a test", "element camera 1",
               "property float fx", "property list uchar int ids", "element vertex 2",
               "property float x", "property float y", "property float z", "property float nx",
               "property uchar red", "property uchar green", "property uchar blue",
               "element face 1", "property list uchar int vertex_indices", "end_header"]
        with open(path, "wb") as f:
            f.write((nl.join(hdr) + nl).encode())
            if fmt == "ascii":
                f.write(b"1.5 2 7 8\n")
                for p, c in zip(xyz, rgb):
                    f.write(("%r %r %r 0.0 %d %d %d\n" % (float(p[0]), float(p[1]), float(p[2]),
                                                          c[0], c[1], c[2])).encode())
                f.write(b"3 0 1 0\n")
            else:
                f.write(struct.pack(end + "fBii", 1.5, 2, 7, 8))
                for p, c in zip(xyz, rgb):
                    f.write(struct.pack(end + "ffffBBB", p[0], p[1], p[2], 0.0, *[int(v) for v in c]))
                f.write(struct.pack(end + "Biii", 3, 0, 1, 0))

    for fmt, nl in [("binary_little_endian", "\n"), ("binary_little_endian", "\r\n"),
                    ("binary_big_endian", "\n"), ("ascii", "\n")]:
        p = str(tmp_path / f"{fmt}_{len(nl)}.ply")
        write(p, fmt, nl)
        x2, c2 = io_utils.read_ply_xyzrgb(p)
        assert np.array_equal(x2, xyz) and np.array_equal(c2, rgb), (fmt, nl)


def test_ply_reader_errors(tmp_path):
    import pytest
    p = str(tmp_path / "bad.ply")
    open(p, "wb").write(b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\n")
    with pytest.raises(ValueError, match="end_header"):
        io_utils.read_ply_xyzrgb(p)
    open(p, "wb").write(b"ply\nformat binary_little_endian 1.0\nelement vertex 1\n"
                        b"property quaternion x\nend_header\n")
    with pytest.raises(ValueError, match="unsupported PLY property type"):
        io_utils.read_ply_xyzrgb(p)
    open(p, "wb").write(b"ply\nformat binary_little_endian 1.0\nelement vertex 1\n"
                        b"property float x\nproperty float y\nproperty float z\nend_header\n" + b"\0" * 12)
    with pytest.raises(ValueError, match="'red' missing"):
        io_utils.read_ply_xyzrgb(p)


def test_reference_demo_cloud_fixture():
    """tests/golden/clock.ply is the reference's dataset/demo_data/clock.ply (BASELINE configs[0])."""
    import os
    from golden_util import GOLDEN_DIR
    xyz, rgb = io_utils.read_ply_xyzrgb(os.path.join(GOLDEN_DIR, "clock.ply"))
    assert xyz.shape == (30000, 3) and rgb.shape == (30000, 3) and rgb.dtype == np.uint8
