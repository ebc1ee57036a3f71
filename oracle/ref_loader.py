"""Oracle (TEST INFRASTRUCTURE): import the reference's OWN source under a stub loader.

Usable where /root/reference exists (the build container) or where oracle/make_ref.py has put
its verbatim copy of the hot path's files under baseline/_ref/ (the GPU box).  It is how the
numpy restatement in this package is pinned: `tests/golden/make_golden*.py` run the reference
functions through this loader and commit their outputs as fixtures; on the GPU box
`tests/test_ddnm_reference_chain_gpu.py` and bench.py's `gpu_baseline` leg run the reference's
own sampler + UNetModel through it next to the CUDA path.

The reference imports kaolin, nvdiffrast, open3d, trimesh, munch, kiui, matplotlib, ... which
are not installed.  They are replaced by empty stub modules, and the FOUR third-party
arithmetic calls the hot path makes are replaced by shims that implement the canonical rules
documented in oracle/camera.py and oracle/project.py:
   kaolin.render.camera.Camera            -> TorchCamera (same fp32 op order as oracle.camera)
   kaolin.metrics.pointcloud.sided_distance -> exact argmin, lowest index on ties
   nvdiffrast.torch.rasterize / interpolate -> oracle.project.rasterize / interpolate
   kaolin.render.mesh.texture_mapping     -> F.grid_sample (kaolin's published bilinear branch)
   kaolin.ops.mesh.face_normals           -> normalised cross product
   xatlas.parametrize                     -> result supplied by the test (third-party, absent)
   kaolin.ops.mesh.uniform_laplacian      -> dense adjacency / degree (kaolin's published code)
   trimesh.grouping.unique_rows, trimesh.geometry.faces_to_edges -> oracle.neighbors restatements
   open3d ... hidden_point_removal        -> scipy ConvexHull on the spherically flipped cloud
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

_REPO_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                         "baseline", "_ref")
# /root/reference in the build container; on the GPU box the verbatim copy of the hot path's
# files that oracle/make_ref.py placed under baseline/_ref/ (git-ignored, shipped by gpurun)
REFERENCE_ROOT = os.environ.get("PDR_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/pointdreamer") else _REPO_REF)

_STUBS = [
    "kaolin", "kaolin.render", "kaolin.render.camera", "kaolin.metrics",
    "kaolin.metrics.pointcloud", "kaolin.ops", "kaolin.ops.mesh", "kaolin.render.mesh",
    "nvdiffrast", "nvdiffrast.torch", "open3d", "trimesh", "xatlas", "munch", "kiui",
    "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "plyfile",
    "skimage", "skimage.metrics", "seaborn",
    "torch_cluster", "torch_geometric", "mcubes", "pymeshlab", "vtk", "lpips", "imageio", "pytz",
]


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pointdreamer"))


class _Munch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @classmethod
    def fromDict(cls, d):
        if isinstance(d, dict):
            return cls({k: cls.fromDict(v) for k, v in d.items()})
        if isinstance(d, list):
            return [cls.fromDict(v) for v in d]
        return d


def _make_torch_camera():
    import torch

    from . import camera as ocam

    class TorchCamera:
        """kaolin Camera shim; elementwise fp32 torch ops in oracle.camera.transform's order."""

        def __init__(self, eye, at, up, fov, width, height, device="cpu"):
            self.params = torch.from_numpy(ocam.view_params(eye, at, up, fov))
            self.width = int(width)
            self.height = int(height)

        @classmethod
        def from_args(cls, eye, at, up, fov, width, height, device="cpu", **kw):
            return cls(np.asarray(eye), np.asarray(at), np.asarray(up), fov, width, height)

        def transform(self, pts):
            batched = pts.dim() == 3
            q = pts.reshape(-1, 3).float()
            p = self.params
            x, y, z = q[:, 0], q[:, 1], q[:, 2]
            cx = ((p[0] * x + p[1] * y) + p[2] * z) + p[9]
            cy = ((p[3] * x + p[4] * y) + p[5] * z) + p[10]
            cz = ((p[6] * x + p[7] * y) + p[8] * z) + p[11]
            d = -cz
            out = torch.stack([(cx * p[12]) / d, (cy * p[12]) / d, p[13] - p[14] / d], 1)
            return out.unsqueeze(0) if batched else out

    return TorchCamera


def _sided_distance(p1, p2):
    import torch
    a = p1[0].long()
    b = p2[0].long()
    if b.shape[0] == 0:
        raise ValueError("sided_distance: empty p2 (view without valid points)")
    idx = torch.empty(a.shape[0], dtype=torch.long)
    dist = torch.empty(a.shape[0], dtype=torch.float32)
    for s in range(0, a.shape[0], 512):
        d = ((a[s:s + 512, None, :] - b[None, :, :]) ** 2).sum(-1)
        m, i = d.min(1)
        # torch.min returns the first minimal index on CPU; make it explicit
        first = (d == m[:, None]).float().argmax(1)
        idx[s:s + 512] = first
        dist[s:s + 512] = m.float()
    return dist[None], idx[None]


def _rasterize(glctx, pos, tri, resolution, grad_db=False):
    import torch

    from . import project as oproj
    depth, fidx, mask, bary = oproj.rasterize(pos.detach().cpu().numpy(), tri.cpu().numpy(),
                                              int(resolution[0]), return_bary=True)
    V, H, W = depth.shape
    rast = np.zeros((V, H, W, 4), dtype=np.float32)
    rast[..., 0:2] = bary
    rast[..., 2] = depth
    rast[..., 3] = (fidx + 1).astype(np.float32)
    return torch.from_numpy(rast), None


def _interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    """nvdiffrast.torch.interpolate shim -> oracle.project.interpolate (barycentrics from
    rast[..., 0:2], triangle id + 1 in rast[..., 3])."""
    import torch

    from . import project as oproj
    a = attr.detach().cpu().numpy()
    if a.ndim == 3:
        a = a[0]
    r = rast.detach().cpu().numpy()
    fidx = r[..., 3].astype(np.int64) - 1
    out = oproj.interpolate(np.ascontiguousarray(r[..., 0:2]), fidx, a, tri.cpu().numpy())
    return torch.from_numpy(out), None


def _texture_mapping(texture_coordinates, texture_maps, mode="nearest"):
    """kaolin 0.15.0 render.mesh.texture_mapping, bilinear branch (published algorithm;
    call site ours_utils.py:1721)."""
    if mode != "bilinear":
        raise NotImplementedError("texture_mapping shim: bilinear only")
    from . import optimize as oopt
    return oopt.texture_mapping_bilinear(texture_coordinates, texture_maps)


def _face_normals(face_vertices, unit=False):
    import torch
    v0, v1, v2 = face_vertices[..., 0, :], face_vertices[..., 1, :], face_vertices[..., 2, :]
    n = torch.cross(v1 - v0, v2 - v0, dim=-1)
    if unit:
        n = torch.nn.functional.normalize(n, dim=-1)
    return n


def _uniform_laplacian(num_vertices, faces):
    """kaolin 0.15.0 ops.mesh.uniform_laplacian (published algorithm; call site
    unproject.py:144): dense binary adjacency / vertex degree, diagonal -1, NaN -> 0."""
    import torch
    fwd = torch.stack([faces, torch.roll(faces, 1, dims=-1)], dim=-1)
    bwd = torch.stack([torch.roll(faces, 1, dims=-1), faces], dim=-1)
    ind = torch.cat([fwd, bwd], dim=1).reshape(-1, 2).unique(dim=0)
    adj = torch.zeros(num_vertices, num_vertices)
    adj[ind[:, 0], ind[:, 1]] = 1.0
    L = torch.div(adj, torch.sum(adj, dim=1).view(-1, 1))
    torch.diagonal(L)[:] = -1
    L[torch.isnan(L)] = 0
    return L


_xatlas_result = None


def set_xatlas_parametrization(vmapping, indices, uvs):
    """xatlas.parametrize is third-party (and absent): tests supply its result."""
    global _xatlas_result
    _xatlas_result = (np.asarray(vmapping), np.asarray(indices), np.asarray(uvs))


def _xatlas_parametrize(vertices, faces):
    if _xatlas_result is None:
        raise RuntimeError("xatlas shim: call set_xatlas_parametrization first")
    return _xatlas_result


@contextlib.contextmanager
def cuda_literals_to_cpu():
    """optimize_color hard-codes device='cuda' in one torch.ones call (ours_utils.py:1670) whose
    result is never used on the nvdiffrast branch; let it run on a CPU-only machine."""
    import torch
    orig = torch.ones

    def ones(*a, **k):
        if k.get("device") == "cuda" and not torch.cuda.is_available():
            k["device"] = "cpu"
        return orig(*a, **k)

    torch.ones = ones
    try:
        yield
    finally:
        torch.ones = orig


class _Dummy:
    """Placeholder for any name imported from a stubbed package; using it is an error."""

    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise RuntimeError(f"stubbed third-party symbol {self._name} was called")

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Dummy(self._name + "." + k)


class _StubModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Dummy(self.__name__ + "." + k)


class _StubFinder:
    """meta-path finder: any submodule of a stubbed top-level package resolves to a stub."""

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        top = fullname.split(".")[0]
        if top in {n.split(".")[0] for n in _STUBS}:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m._pdr_stub = True
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _O3dPointCloud:
    """open3d.geometry.PointCloud shim exposing only hidden_point_removal (ours_utils.py:209-214)."""

    def __init__(self, points=None):
        self.points = np.asarray(points, dtype=np.float64)

    def hidden_point_removal(self, camera_location, radius):
        return None, list(hidden_point_removal_scipy(self.points, camera_location, radius))


def install_stubs():
    """Register stub modules + shims; idempotent."""
    if "kaolin" in sys.modules and getattr(sys.modules["kaolin"], "_pdr_stub", False):
        return
    for name in _STUBS:
        m = _StubModule(name)
        m._pdr_stub = True
        m.__path__ = []
        sys.modules[name] = m
    for name in _STUBS:
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])
    sys.modules["munch"].Munch = _Munch
    kiui = sys.modules["kiui"]
    kiui.lo = lambda *a, **k: None
    kiui.seed_everything = lambda *a, **k: None
    sys.modules["kaolin.render.camera"].Camera = _make_torch_camera()
    sys.modules["kaolin.metrics.pointcloud"].sided_distance = _sided_distance
    sys.modules["nvdiffrast.torch"].rasterize = _rasterize
    sys.modules["nvdiffrast.torch"].interpolate = _interpolate
    import torch as _torch
    krm = sys.modules["kaolin.render.mesh"]
    krm.texture_mapping = _texture_mapping
    krm.prepare_vertices = lambda *a, **k: (None, None, None)  # result unused (ours_utils.py:1655)
    krc = sys.modules["kaolin.render.camera"]
    krc.generate_transformation_matrix = lambda *a, **k: _torch.zeros(1)   # results unused
    krc.generate_perspective_projection = lambda *a, **k: _torch.zeros(3, 1)
    krc.perspective_camera = lambda *a, **k: None
    sys.modules["kaolin.ops.mesh"].face_normals = _face_normals
    sys.modules["xatlas"].parametrize = _xatlas_parametrize
    sys.modules["kaolin.ops.mesh"].uniform_laplacian = _uniform_laplacian
    from . import neighbors as _onb
    for sub_name in ("trimesh.grouping", "trimesh.geometry"):
        mod = _StubModule(sub_name)
        mod._pdr_stub = True
        mod.__path__ = []
        sys.modules[sub_name] = mod
        setattr(sys.modules["trimesh"], sub_name.split(".")[1], mod)
    sys.modules["trimesh.grouping"].unique_rows = _onb.unique_rows      # restated trimesh
    sys.modules["trimesh.geometry"].faces_to_edges = _onb.faces_to_edges
    sys.modules["nvdiffrast.torch"].RasterizeCudaContext = lambda *a, **k: None
    o3d = sys.modules["open3d"]
    o3d.geometry = types.SimpleNamespace(PointCloud=_O3dPointCloud)
    o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: np.asarray(a, dtype=np.float64))
    plt = sys.modules["matplotlib.pyplot"]
    plt.figure = plt.imshow = plt.show = plt.axis = lambda *a, **k: None
    sys.meta_path.append(_StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def patch_determinism():
    """Determinism switches the reference needs: deterministic index_put winner,
    pinned-torchvision Resize semantics (antialias=False)."""
    import torch
    torch.use_deterministic_algorithms(True)
    import torchvision.transforms.transforms as T
    if not getattr(T.Resize, "_pdr_patched", False):
        orig = T.Resize.__init__

        def init(self, size, *a, **k):
            k["antialias"] = False
            orig(self, size, *a, **k)

        T.Resize.__init__ = init
        T.Resize._pdr_patched = True


@contextlib.contextmanager
def quiet():
    """The reference prints tensors mid-function (unproject.py:363-365)."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield


_mods = {}


def load(name):
    """Import a reference module (e.g. 'pointdreamer.ours_utils') under the stubs, with cwd at
    the reference root (some modules open relative paths)."""
    if name in _mods:
        return _mods[name]
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    install_stubs()
    patch_determinism()
    import importlib
    import warnings
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with quiet():
                _mods[name] = importlib.import_module(name)
    finally:
        os.chdir(cwd)
    return _mods[name]


def hidden_point_removal_scipy(points, eye, radius):
    """open3d hidden_point_removal shim (see oracle/hpr.py)."""
    from .hpr import hidden_point_removal
    return hidden_point_removal(points, eye, radius)
