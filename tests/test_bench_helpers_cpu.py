"""bench.py host-side helpers: the algorithmic-byte model of the HBM-bound stages reproduces
SURVEY.md section 8(d)'s figures, and every BASELINE.json config index maps to a workload."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_algorithmic_bytes_match_survey_8d():
    c = dict(bench.CONFIGS[1])
    b = bench.hbm_algorithmic_bytes(c, n_points=30000, n_verts=5000, n_faces=10000)
    assert abs(b["project_splat"] / 1e6 - 20.6) < 0.2     # N*24 + V*N*4 + 3*V*3*res^2*4
    assert abs(b["raster"] / 1e6 - 19.0) < 0.2            # (Vm+F)*12 + V*cam_res^2*9
    assert abs(b["unproject_nbf"] / 1e6 - 50.4) < 0.3     # R^2*17 + V*cam^2*4 + V*3*res^2*4 + F*12 + R^2*17
    assert abs(b["fill_atlas"] / 1e6 - 33.6) < 0.1        # 2 * R^2 * 16
    assert b["fill_views"] == 0                           # DDNM config: no nearest fill of the views
    c0 = dict(bench.CONFIGS[0])
    b0 = bench.hbm_algorithmic_bytes(c0, 30000, 5000, 10000)
    assert b0["fill_views"] == c0["V"] * c0["res"] ** 2 * 32
    assert b0["total"] == sum(v for k, v in b0.items() if k != "total")


def test_every_baseline_config_has_a_workload():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert sorted(bench.CONFIGS) == list(range(len(base["configs"])))
    assert "clock.ply" in bench.CONFIGS[0]["workload"] and bench.CONFIGS[0]["method"] == "nearest"
    assert bench.CONFIGS[1]["V"] == 8 and bench.CONFIGS[1]["res"] == 256
    assert bench.CONFIGS[3]["S"] == 8 and bench.CONFIGS[4]["V"] == 16 and bench.CONFIGS[4]["res"] == 512
    for i, c in bench.CONFIGS.items():
        cfg, flow = bench.path_config(c)
        assert cfg["view_num"] == c["V"] and cfg["res"] == c["res"] and cfg["texture_gen_method"] == c["method"]
        assert flow == c.get("flow", "path")
        if flow == "path":
            assert cfg["complete_unseen_by"] == "unproject" and cfg["optimize_from"] is None
        else:
            assert cfg["complete_unseen_by"] == "neighbor" and cfg["optimize_from"] == "ours"
