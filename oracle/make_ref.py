"""Oracle (TEST INFRASTRUCTURE): populate baseline/_ref/ with the UNMODIFIED reference files of
the hot path so that the reference's own code can run on the GPU box (which has no
/root/reference).  baseline/_ref/ is git-ignored (never part of the history) but travels with
the gpurun snapshot, like the built libpdr.so.

    python -m oracle.make_ref            (also called by __graft_entry__.build() when
                                          /root/reference is present)

Copied verbatim, tree layout preserved (SURVEY.md §8c "Reference files the oracle executes"):
  models/DDNM/**            sampler, U-Net, configs (diffusion.py:459-570, unet.py, nn.py, ...)
  pointdreamer/ours_utils.py, pointdreamer/unproject.py
  utils/*.py                camera_utils, utils_2d, other_utils, mesh_utils, metric_utils/...
  models/get3d/**           extract_texture_map.py, get3d_utils/utils_3d.py
  configs/*.yaml, demo.py
  dataset/demo_data/*.ply, dataset/NBF_demo_data/*.ply      (BASELINE configs[0], [2])
Consumers: oracle/ref_loader.py (REFERENCE_ROOT falls back to baseline/_ref), the `-m gpu`
chain-parity test, and bench.py's `gpu_baseline` / `--impl reference` legs.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("PDR_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

TREES = ["models/DDNM", "models/get3d", "utils", "configs", "dataset/demo_data", "dataset/NBF_demo_data"]
FILES = ["pointdreamer/ours_utils.py", "pointdreamer/unproject.py",
         "demo.py"]
KEEP_EXT = (".py", ".yml", ".yaml", ".ply")


def populate(verbose=False):
    """Copy the listed reference files into baseline/_ref/.  Returns the number of files."""
    if not os.path.isdir(os.path.join(SRC, "pointdreamer")):
        raise RuntimeError(f"reference not present at {SRC}")
    n = 0
    todo = list(FILES)
    for tree in TREES:
        for dirpath, _, names in os.walk(os.path.join(SRC, tree)):
            for name in names:
                if name.endswith(KEEP_EXT):
                    todo.append(os.path.relpath(os.path.join(dirpath, name), SRC))
    for rel in sorted(set(todo)):
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src) or \
                os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copy2(src, dst)
        n += 1
        if verbose:
            print(rel)
    return n


if __name__ == "__main__":
    print(f"{populate('-v' in sys.argv)} reference files under {DST}")
