"""The numpy oracle at PRODUCTION size (30 000 points, 8 views, 256^2 / 512^2, atlas 1024^2, NBF [21])
against the digests of the reference's own run on the reference's clock.ply
(tests/golden/make_golden_production.py): every boundary tensor bit-exact."""
import hashlib
import json
import os
import sys

import numpy as np

from oracle.pipeline import run_path as oracle_pipeline

from golden_util import GOLDEN_DIR

sys.path.insert(0, GOLDEN_DIR)


def test_oracle_matches_reference_digests_clock():
    from make_golden_production import production_scene
    gold = json.load(open(os.path.join(GOLDEN_DIR, "production_digests.json")))
    got = oracle_pipeline(gold["config"], production_scene("clock"))
    bad = []
    for k, w in sorted(gold["clouds"]["clock"]["digests"].items()):
        a = np.ascontiguousarray(np.asarray(got[k]).astype(np.dtype(w["dtype"]), copy=False))
        if list(a.shape) != w["shape"]:
            bad.append(f"{k}: shape {a.shape} != {w['shape']}")
        elif hashlib.sha256(a.tobytes()).hexdigest() != w["sha256"]:
            bad.append(f"{k}: digest differs")
    assert not bad, "; ".join(bad)
