"""Full-size DDNM chain parity study (shared by tests/test_ddnm_reference_chain_gpu.py and
tools/chain_parity_report.py): the reference's OWN sampler + UNetModel (oracle/reference_ddnm.py,
unmodified code from baseline/_ref) next to the CUDA path on the same B200, same weights, same
sparse images / masks, same noise stream (seed 42, offset 0).

Reported per configuration:
  * free-running chains: per-step max-abs distance of x_t to the reference fp16 chain for
      ours, the reference re-run with another cuDNN algorithm choice (cudnn.benchmark=True) and
      the reference in fp32 (TF32 off) - i.e. the CUDA path's drift next to the reference's own
      run-to-run / precision envelope;
  * teacher-forced forwards: eps(x_t of the reference chain) of our engine (every step) and of
      the fp32 reference model (every `tf_every`-th step) against the reference's fp16 eps;
  * final images: max-abs, mean-abs, PSNR; known pixels exact.
"""
import ctypes

import numpy as np
import torch


def _psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else float(10 * np.log10(1.0 / mse))


def scene_inputs(dev, n_views, res, seed=0, n_points=30000):
    """Sparse images + masks of the synthetic bench scene through the CUDA geometry path."""
    from pointdreamer_b200 import demo, ours_utils, synthetic
    cam_res = 2 * res
    cfg = dict(demo.DEFAULT_CONFIG, view_num=n_views, res=res, cam_res=cam_res)
    sc = synthetic.make_scene(n_points, seed=seed, atlas_res=256)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cam = demo.prepare_cameras(cfg, dev)
    (hm, _, depths, _, _, _, _, puv, pdepth) = ours_utils.get_rendered_hard_mask_and_face_idx_batch(
        cam["cams"], t(sc["vertices"]), t(sc["faces"]), t(sc["xyz"]), rescale=True, padding=0.05)
    hm = ours_utils.resize_hard_masks(hm, res)
    pv, _ = ours_utils.get_point_validation_by_depth(cam_res, puv, pdepth, depths, offset=0.0001)
    pp = ours_utils.get_point_pixels(puv, res)
    sparse, m0, m2, _ = ours_utils.get_sparse_images(pp, t(sc["rgb"]), pv, hm, None, n_views, res,
                                                     1, 1, 0.82)
    return sparse, m2


def ours_stepwise(inp, sparse, mask, record=True):
    """The CUDA sampler driven step by step through the C ABI (pdr_ddnm_prepare / pdr_unet_forward
    / pdr_ddnm_step / pdr_ddnm_final) so that every x_t and eps can be read; bit-identical to the
    single-call pdr_ddnm_sample (asserted by the test)."""
    from pointdreamer_b200 import _lib
    V, _, S, _ = sparse.shape
    dev = sparse.device
    steps = len(inp.ts)
    x = torch.empty(V, 3, S, S, device=dev)
    y = torch.empty_like(x)
    out = torch.empty_like(x)
    seed, base, dpc = ctypes.c_ulonglong(inp.seed), ctypes.c_ulonglong(inp.offset), \
        ctypes.c_ulonglong(steps + 1)
    sp, mk = sparse.float().contiguous(), mask.float().contiguous()
    _lib.call("pdr_ddnm_prepare", sp, mk, V, S, seed, base, dpc, 0, y, x)
    xs, ets = [], []
    for s in range(steps):
        t = torch.full((V,), float(inp.ts[s]), device=dev)
        et = inp.model.forward(x, t)
        if record:
            xs.append(x.clone())
            ets.append(et[:, :3].clone())
        c = np.ascontiguousarray(inp.coefs[s])
        _lib.call("pdr_ddnm_step", x, et, et.shape[1], y, mk, V, S,
                  c.ctypes.data_as(ctypes.c_void_p), seed, base, dpc, 0, 1 + s)
    _lib.call("pdr_ddnm_final", x, ctypes.c_longlong(x.numel()), out)
    return out, xs, ets


def study(dev, model_cfg, n_views, T, tf_every=10, with_fp32=True, with_benchmark=True, seed=42):
    """Runs everything; returns a dict of plain floats / lists (JSON-serialisable)."""
    from oracle import reference_ddnm as rd
    from pointdreamer_b200.ddnm_inpainting import DEFAULT_DDNM_CONFIG, Inpainter
    from pointdreamer_b200.unet import random_state_dict
    S = model_cfg["image_size"]
    sparse, m2 = scene_inputs(dev, n_views, S)
    mask = m2[:, 0].contiguous()
    sd = random_state_dict(model_cfg, seed=1234, device=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = dict(views=n_views, T_sampling=T, image_size=S,
               known_fraction=float(mask.mean()))

    def ref_chain(model, benchmark):
        torch.backends.cudnn.benchmark = benchmark
        rec = rd.Recorder(model)
        runner = rd.build_runner(dev, image_size=S, T_sampling=T)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            out = rd.run_views(rd.reference_inpainter(runner, rec), sparse, m2, seed=seed)
        e1.record()
        torch.cuda.synchronize()
        torch.backends.cudnn.benchmark = False
        # [T][V,3,S,S]: the reference runs view after view; regroup by step
        xs = [torch.cat([rec.xs[v * T + s] for v in range(n_views)]) for s in range(T)]
        ets = [torch.cat([rec.ets[v * T + s] for v in range(n_views)]) for s in range(T)]
        return out, xs, ets, e0.elapsed_time(e1) / 1e3

    m16 = rd.build_model(model_cfg, sd, dev, fp16=True)
    ref, rxs, rets, secs = ref_chain(m16, False)
    res["reference_fp16_seconds"] = secs
    known = (mask[:, None] > 0).expand_as(ref)
    res["reference_known_pixel_max_err"] = float((ref - sparse)[known].abs().max())

    inp = Inpainter(dev, state_dict=sd, model_config=model_cfg,
                    ddnm_config=dict(DEFAULT_DDNM_CONFIG, T_sampling=T), seed=seed, offset=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    inp.inpaint_batch(sparse, mask, chain0=0)  # warm-up (plan + graph capture)
    e0.record()
    ours_one_call = inp.inpaint_batch(sparse, mask, chain0=0)
    e1.record()
    torch.cuda.synchronize()
    res["ours_seconds"] = e0.elapsed_time(e1) / 1e3
    ours, oxs, _ = ours_stepwise(inp, sparse, mask)
    res["stepwise_equals_single_call"] = bool(torch.equal(ours, ours_one_call))
    res["ours_known_pixel_max_err"] = float((ours - sparse)[known].abs().max())
    res["x_T_equal"] = bool(torch.equal(oxs[0], rxs[0]))

    def curve(xs):
        return [float((a - b).abs().max()) for a, b in zip(xs, rxs)]

    def final(a, name):
        d = (a - ref).abs()
        res[name] = dict(max_abs=float(d.max()), mean_abs=float(d.mean()), psnr_db=_psnr(a, ref),
                         frac_gt_1e3=float((d > 1e-3).float().mean()),
                         psnr_8bit_db=_psnr((a * 255).floor() / 255, (ref * 255).floor() / 255))

    res["drift_ours"] = curve(oxs)
    final(ours, "final_ours_vs_ref16")
    # teacher-forced: our eps on the reference chain's own x_t, every step
    tf = []
    for s in range(T):
        t = torch.full((n_views,), float(inp.ts[s]), device=dev)
        e = inp.model.forward(rxs[s], t)[:, :3]
        tf.append(float((e - rets[s]).abs().max()))
    res["teacher_forced_ours_vs_ref16"] = tf
    res["eps_abs_max"] = float(max(float(e.abs().max()) for e in rets))
    res["eps_std"] = float(torch.stack([e.std() for e in rets]).mean())
    del oxs

    if with_benchmark:
        refb, bxs, _, secs = ref_chain(m16, True)
        res["reference_fp16_benchmark_seconds"] = secs
        res["drift_ref16_cudnn_benchmark"] = curve(bxs)
        final(refb, "final_ref16_cudnn_benchmark_vs_ref16")
        del bxs
    if with_fp32:
        m32 = rd.build_model(model_cfg, sd, dev, fp16=False)
        ref32, fxs, _, secs = ref_chain(m32, False)
        res["reference_fp32_seconds"] = secs
        res["drift_ref32"] = curve(fxs)
        final(ref32, "final_ref32_vs_ref16")
        d = (ours - ref32).abs()
        res["final_ours_vs_ref32"] = dict(max_abs=float(d.max()), mean_abs=float(d.mean()),
                                          psnr_db=_psnr(ours, ref32))
        del fxs
        tf32 = {}
        with torch.no_grad():
            for s in range(0, T, tf_every):
                t = torch.full((1,), float(inp.ts[s]), device=dev)
                e = torch.cat([m32(rxs[s][v:v + 1], t)[:, :3] for v in range(n_views)])
                tf32[s] = float((e - rets[s]).abs().max())
        res["teacher_forced_ref32_vs_ref16"] = tf32
    return res
