"""Production-size parity of the CUDA geometry path against the reference's OWN code on the
reference's five demo clouds (30 000 points, 8 views, 256^2 / 512^2, atlas 1024^2, NBF [21], HPR on):
SHA-256 of every boundary tensor must equal the digest recorded by
tests/golden/make_golden_production.py (which ran the reference functions through the stub loader).
Everything here is INT or an exact fp32 copy / fixed-order fp32 expression, so the bar is bit-exact."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from golden_util import GOLDEN_DIR
from test_geometry_gpu import _run_pipeline

pytestmark = pytest.mark.gpu

sys.path.insert(0, GOLDEN_DIR)
GOLD = json.load(open(os.path.join(GOLDEN_DIR, "production_digests.json")))


@pytest.mark.parametrize("name", sorted(GOLD["clouds"]))
def test_demo_cloud_production_size(cuda, name):
    from make_golden_production import production_scene
    cfg = GOLD["config"]
    sc = production_scene(name)
    got = _run_pipeline(cfg, sc, cuda)
    want = GOLD["clouds"][name]["digests"]
    bad = []
    for k, w in sorted(want.items()):
        a = np.ascontiguousarray(got[k].astype(np.dtype(w["dtype"]), copy=False))
        if list(a.shape) != w["shape"]:
            bad.append(f"{k}: shape {a.shape} != {w['shape']}")
        elif hashlib.sha256(a.tobytes()).hexdigest() != w["sha256"]:
            bad.append(f"{k}: digest differs")
    assert not bad, f"{name}: " + "; ".join(bad)
    info = GOLD["clouds"][name]["info"]
    assert int(got["atlas_painted_mask"].sum()) == info["painted_texels"]
