"""Host side of the native ADM U-Net engine (libpdr.so: unet_engine.cu).

Takes the arguments of the reference's `create_model` (script_util.py:130-185, values from
models/DDNM/configs/imagenet_256.yml) and a state_dict with the REFERENCE's parameter names
(so `256x256_diffusion_uncond.pt` loads unchanged), re-lays the weights out once for the
kernels (conv weights -> fp16 [Cout][tap][Cin]; everything the reference keeps in fp32 stays
fp32), and exposes `forward(x, timesteps)` like UNetModel.forward (unet.py:635-664).
"""
import ctypes

import torch

from . import _lib

DEFAULT_MODEL_CONFIG = dict(  # models/DDNM/configs/imagenet_256.yml:14-33
    image_size=256, in_channels=3, model_channels=256, out_channels=6, num_res_blocks=2,
    attention_resolutions="32,16,8", channel_mult=(1, 1, 2, 2, 4, 4), num_head_channels=64,
    num_heads=4, use_scale_shift_norm=True, resblock_updown=True, use_fp16=True,
    use_new_attention_order=False)


class PdrUnetConfig(ctypes.Structure):
    _fields_ = [("image_size", ctypes.c_int), ("in_channels", ctypes.c_int),
                ("model_channels", ctypes.c_int), ("out_channels", ctypes.c_int),
                ("num_res_blocks", ctypes.c_int), ("n_mult", ctypes.c_int),
                ("channel_mult_x2", ctypes.c_int * 8), ("n_attn_ds", ctypes.c_int),
                ("attn_ds", ctypes.c_int * 8), ("num_head_channels", ctypes.c_int)]


def channel_mult_for(image_size):
    """script_util.py:149-160."""
    return {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4),
            64: (1, 2, 3, 4)}[image_size]


def _res_block_names(cfg):
    """Prefixes of every ResBlock in module construction order (unet.py:482-611)."""
    mult = cfg["channel_mult"]
    nrb = cfg["num_res_blocks"]
    attn_ds = [cfg["image_size"] // int(r) for r in cfg["attention_resolutions"].split(",")]
    names = []
    blk, ds = 1, 1
    for level in range(len(mult)):
        for _ in range(nrb):
            names.append(f"input_blocks.{blk}.0")
            blk += 1
        if level != len(mult) - 1:
            names.append(f"input_blocks.{blk}.0")
            blk += 1
            ds *= 2
    names += ["middle_block.0", "middle_block.2"]
    blk = 0
    for level in reversed(range(len(mult))):
        for i in range(nrb + 1):
            names.append(f"output_blocks.{blk}.0")
            sub = 1
            if ds in attn_ds:
                sub += 1
            if level and i == nrb:
                names.append(f"output_blocks.{blk}.{sub}")
                ds //= 2
            blk += 1
    return names


class UNetEngine:
    """B200-native UNetModel: `engine = UNetEngine(state_dict, cfg); y = engine(x, t)`."""

    def __init__(self, state_dict, cfg=None, device="cuda"):
        cfg = dict(DEFAULT_MODEL_CONFIG if cfg is None else cfg)
        if not cfg.get("use_scale_shift_norm", True) or not cfg.get("resblock_updown", True) \
                or cfg.get("use_new_attention_order", False):
            raise NotImplementedError("only the imagenet_256.yml variant of the ADM U-Net "
                                      "(scale-shift norm, resblock up/down, legacy attention)")
        self.cfg = cfg
        self.device = torch.device(device)
        self.lib = _lib.load()
        c = PdrUnetConfig()
        c.image_size = cfg["image_size"]
        c.in_channels = cfg["in_channels"]
        c.model_channels = cfg["model_channels"]
        c.out_channels = cfg["out_channels"]
        c.num_res_blocks = cfg["num_res_blocks"]
        mult = cfg["channel_mult"]
        c.n_mult = len(mult)
        for i, m in enumerate(mult):
            c.channel_mult_x2[i] = int(round(m * 2))
        ads = [cfg["image_size"] // int(r) for r in cfg["attention_resolutions"].split(",")]
        c.n_attn_ds = len(ads)
        for i, d in enumerate(ads):
            c.attn_ds[i] = d
        c.num_head_channels = cfg["num_head_channels"]
        self._handle = ctypes.c_void_p()
        _lib.check(self.lib.pdr_unet_create(ctypes.byref(c), ctypes.byref(self._handle)),
                   "pdr_unet_create")
        self._tensors = {}  # keeps the device copies alive
        self._load(state_dict)
        self._workspace = None
        self._batch = 0

    def __del__(self):
        try:
            if self._handle:
                self.lib.pdr_unet_destroy(self._handle)
        except Exception:
            pass

    # ---- parameters ----
    def _set(self, name, t):
        t = t.contiguous()
        self._tensors[name] = t
        _lib.check(self.lib.pdr_unet_set_param(self._handle, name.encode(), _lib.ptr(t),
                                               ctypes.c_size_t(t.numel() * t.element_size())),
                   "pdr_unet_set_param")

    def _load(self, sd):
        dev = self.device
        emb_w, emb_b = [], []
        res_names = _res_block_names(self.cfg)
        emb_keys = {n + ".emb_layers.1.weight" for n in res_names} | \
                   {n + ".emb_layers.1.bias" for n in res_names}
        for name, v in sd.items():
            if name.startswith("module."):
                name = name[len("module."):]  # DataParallel checkpoints (diffusion.py:456)
            v = v.detach()
            if name in emb_keys:
                continue
            if name.endswith(".weight") and v.dim() in (3, 4) and not name.startswith("out."):
                # torso conv (Conv2d [Co,Ci,kh,kw] or Conv1d [Co,Ci,1]) -> fp16 [Co][tap][Ci]
                if v.dim() == 3:
                    v = v[..., None]
                w = v.to(dev).half().permute(0, 2, 3, 1).reshape(v.shape[0], -1)
                self._set(name, w)
            elif name.startswith("out.2"):
                self._set(name, v.to(dev).float())
            else:
                t = v.to(dev)
                is_torso_bias = name.endswith(".bias") and (name[:-5] + ".weight") in sd and \
                    sd[name[:-5] + ".weight"].dim() in (3, 4) and not name.startswith("out.")
                if is_torso_bias:
                    t = t.half()  # the reference stores torso conv biases in fp16
                self._set(name, t.float())
        for n in res_names:
            emb_w.append(sd[n + ".emb_layers.1.weight"].detach().to(dev).float())
            emb_b.append(sd[n + ".emb_layers.1.bias"].detach().to(dev).float())
        # channel-changing ResBlocks: skip_connection (1x1) is accumulated inside out_layers.3's
        # GEMM -> weights concatenated along K, biases summed in fp32 (unet.py:219-222, 256)
        for n in res_names:
            if n + ".skip_connection.weight" in self._tensors:
                w3 = self._tensors[n + ".out_layers.3.weight"]
                ws = self._tensors[n + ".skip_connection.weight"]
                self._set(n + ".out_layers.3_skip.weight", torch.cat([w3, ws], 1))
                self._set(n + ".out_layers.3_skip.bias",
                          self._tensors[n + ".out_layers.3.bias"] +
                          self._tensors[n + ".skip_connection.bias"])
        # stem on the tensor cores: [C][27] -> zero-padded [C][64] (K index = (ky*3+kx)*3 + c)
        stem = self._tensors["input_blocks.0.0.weight"]
        if stem.shape[0] % 64 == 0:
            w64 = torch.zeros(stem.shape[0], 64, dtype=torch.float16, device=dev)
            w64[:, :27] = stem
            self._set("input_blocks.0.0_tc.weight", w64)
            self._set("input_blocks.0.0_tc.bias", self._tensors["input_blocks.0.0.bias"])
        self._set("emb_all.weight", torch.cat(emb_w, 0))
        self._set("emb_all.bias", torch.cat(emb_b, 0))

    # ---- planning / execution ----
    def plan(self, batch):
        if batch == self._batch:
            return
        need = ctypes.c_size_t(0)
        _lib.check(self.lib.pdr_unet_workspace_bytes(self._handle, int(batch), ctypes.byref(need)),
                   "pdr_unet_workspace_bytes")
        self._workspace = None  # release the previous arena before allocating the next
        ws = torch.empty(need.value + 1024, dtype=torch.uint8, device=self.device)
        off = (-ws.data_ptr()) % 1024
        self._workspace = ws
        _lib.check(self.lib.pdr_unet_plan(self._handle, int(batch),
                                          ctypes.c_void_p(ws.data_ptr() + off),
                                          ctypes.c_size_t(need.value)), "pdr_unet_plan")
        self._batch = batch
        self.workspace_bytes = need.value

    @property
    def handle(self):
        return self._handle

    def forward(self, x, timesteps, n_out=None):
        """x [B,3,S,S] fp32 cuda, timesteps [B] -> [B,n_out,S,S] fp32 (default all channels)."""
        B = x.shape[0]
        self.plan(B)
        n_out = self.cfg["out_channels"] if n_out is None else n_out
        S = self.cfg["image_size"]
        out = torch.empty(B, n_out, S, S, device=self.device)
        _lib.call("pdr_unet_forward", self._handle, x.float().contiguous(),
                  timesteps.float().contiguous(), out, int(n_out))
        return out

    __call__ = forward


def param_shapes(cfg=None):
    """(name -> shape) of every parameter of the reference UNetModel for `cfg`, in module
    construction order (unet.py:441-617)."""
    cfg = dict(DEFAULT_MODEL_CONFIG if cfg is None else cfg)
    mc = cfg["model_channels"]
    mult = cfg["channel_mult"]
    nrb = cfg["num_res_blocks"]
    ted = 4 * mc
    attn_ds = [cfg["image_size"] // int(r) for r in cfg["attention_resolutions"].split(",")]
    shapes = {}

    def lin(n, cin, cout):
        shapes[n + ".weight"] = (cout, cin)
        shapes[n + ".bias"] = (cout,)

    def conv(n, cin, cout, k, dims=2):
        shapes[n + ".weight"] = (cout, cin) + (k,) * dims
        shapes[n + ".bias"] = (cout,)

    def gn(n, ch):
        shapes[n + ".weight"] = (ch,)
        shapes[n + ".bias"] = (ch,)

    def res(p, cin, cout):
        gn(p + ".in_layers.0", cin)
        conv(p + ".in_layers.2", cin, cout, 3)
        lin(p + ".emb_layers.1", ted, 2 * cout)
        gn(p + ".out_layers.0", cout)
        conv(p + ".out_layers.3", cout, cout, 3)
        if cin != cout:
            conv(p + ".skip_connection", cin, cout, 1)

    def attn(p, ch):
        gn(p + ".norm", ch)
        conv(p + ".qkv", ch, 3 * ch, 1, dims=1)
        conv(p + ".proj_out", ch, ch, 1, dims=1)

    lin("time_embed.0", mc, ted)
    lin("time_embed.2", ted, ted)
    ch = int(mult[0] * mc)
    conv("input_blocks.0.0", cfg["in_channels"], ch, 3)
    chans = [ch]
    blk, ds = 1, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            res(f"input_blocks.{blk}.0", ch, int(m * mc))
            ch = int(m * mc)
            if ds in attn_ds:
                attn(f"input_blocks.{blk}.1", ch)
            chans.append(ch)
            blk += 1
        if level != len(mult) - 1:
            res(f"input_blocks.{blk}.0", ch, ch)
            chans.append(ch)
            blk += 1
            ds *= 2
    res("middle_block.0", ch, ch)
    attn("middle_block.1", ch)
    res("middle_block.2", ch, ch)
    blk = 0
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            res(f"output_blocks.{blk}.0", ch + ich, int(mc * m))
            ch = int(mc * m)
            sub = 1
            if ds in attn_ds:
                attn(f"output_blocks.{blk}.{sub}", ch)
                sub += 1
            if level and i == nrb:
                res(f"output_blocks.{blk}.{sub}", ch, ch)
                ds //= 2
            blk += 1
    gn("out.0", ch)
    conv("out.2", ch, cfg["out_channels"], 3)
    return shapes


def random_state_dict(cfg=None, seed=1234, device="cuda"):
    """Seeded random-init weights of the reference architecture, generated on `device`.

    Used when `256x256_diffusion_uncond.pt` is absent (no network in the build/bench
    environment).  The modules the reference zero-initialises (`zero_module`, unet.py:210-212,
    294, 616) get small random weights too, otherwise the network's output is identically 0."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith(".bias"):
            sd[name] = torch.randn(shape, generator=g, device=device) * 0.02
        elif len(shape) == 1:  # GroupNorm weight
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            gain = 0.5 if (name.endswith("out_layers.3.weight") or "proj_out" in name
                           or name.startswith("out.2")) else 1.0
            if "emb_layers" in name:
                gain = 0.3
            sd[name] = torch.randn(shape, generator=g, device=device) * (gain / fan_in ** 0.5)
    return sd


OP_CLASSES = ("conv_tc", "gn_stats", "gn_apply", "resample", "attention", "linear", "stem", "head")


def profile_begin(engine, every=10, max_forwards=16):
    """Start sampling per-kernel-class CUDA-event timings of `engine` (see pdr_unet_profile_begin)."""
    _lib.check(engine.lib.pdr_unet_profile_begin(engine.handle, int(every), int(max_forwards)),
               "pdr_unet_profile_begin")


def profile_end(engine):
    """Stop sampling; returns {class: dict(ms, flops, launches)} and the number of sampled forwards."""
    n = len(OP_CLASSES)
    ms = (ctypes.c_double * n)()
    fl = (ctypes.c_double * n)()
    la = (ctypes.c_longlong * n)()
    fw = ctypes.c_longlong(0)
    _lib.check(engine.lib.pdr_unet_profile_end(engine.handle, ms, fl, la, ctypes.byref(fw)),
               "pdr_unet_profile_end")
    return {OP_CLASSES[i]: dict(ms=ms[i], flops=fl[i], launches=int(la[i])) for i in range(n)}, int(fw.value)
