"""bench.py — shapes/s of the project -> DDNM-inpaint -> unproject hot path on B200.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank/GPU)
  python bench.py --impl reference ...                   (CPU oracle arm, rank 0 only)

A "step" is one pass of the whole path over one shape per rank (weak scaling): a synthetic
30 000-point coloured cloud + ~10k-triangle mesh + 1024^2 atlas, 8 views, DDNM 100 steps at
256^2 (BASELINE.json configs[1]); random-init weights of the reference architecture.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

V, RES, CAM_RES, ATLAS_RES, N_POINTS, T_STEPS = 8, 256, 512, 1024, 30000, 100
F_UNET = 2.2397e12  # FLOPs of one U-Net forward at 256^2, batch 1
WORKLOAD = "30k-pt synthetic cloud, view_num=8, DDNM_inpaint 256^2 (T_sampling=100), atlas 1024^2"


def path_config(flow="path"):
    from pointdreamer_b200 import demo
    # "path": configs/default.yaml with the two post-path "next" rows disabled (SURVEY §8d config 3,
    # the north-star metric); "default": configs/default.yaml as shipped (complete_unseen_by:
    # neighbor + optimize_from: ours run after the path)
    cfg = dict(demo.DEFAULT_CONFIG, view_num=V, res=RES, cam_res=CAM_RES, xatlas_texture_res=ATLAS_RES)
    if flow == "path":
        cfg.update(complete_unseen_by="unproject", optimize_from=None)
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1364.0), d.get("hbm_gbs", 6556.5), "measured"
    return 1400.0, 6650.0, "fallback"


def cpu_unet_forward_seconds(n_forwards=2):
    """Reference U-Net restatement (oracle, fp32) on the host cores, batch 1 at 256^2."""
    import torch
    from oracle import unet as ounet
    from pointdreamer_b200.unet import random_state_dict
    sd = random_state_dict(ounet.DEFAULT_CONFIG, seed=1234, device="cpu")
    o = ounet.UNetOracle(sd, ounet.DEFAULT_CONFIG, emulate_fp16=False)
    x = torch.randn(1, 3, RES, RES)
    t = torch.tensor([500.0])
    times = []
    for _ in range(n_forwards):
        t0 = time.time()
        o.forward(x, t)
        times.append(time.time() - t0)
    return min(times), times


def cpu_geometry_seconds(scene, cfg):
    """Oracle project/splat/unproject for ONE view, scaled by V by the caller."""
    import numpy as np
    from oracle import camera as ocam, project as oproj, unproject as ounproj
    cams, base_dirs, _, _ = ocam.create_cameras(V, 1.6, CAM_RES)
    params = [c.params for c in cams][:1]
    t0 = time.time()
    pr = oproj.project_vertices_points(params, scene["vertices"], scene["xyz"], True, 0.05)
    depth, fidx, mask = oproj.rasterize(pr["pos"], scene["faces"], CAM_RES)
    hm = oproj.resize_mask_half_any(mask, RES)
    vis, _ = oproj.point_validation_by_depth(CAM_RES, pr["point_uvs"], pr["point_depths"], depth, 1e-4)
    pp = oproj.point_pixels(pr["point_uvs"], RES)
    sparse, m0, m2, scales = oproj.get_sparse_images(pp, scene["rgb"], vis, hm, 1, RES, 1, 1, 0.82)
    xa = scene["xatlas_dict"]
    ounproj.unproject(sparse, scene["f_normals"], RES, params, CAM_RES, base_dirs[:1], xa["gb_pos"],
                      xa["mask"], xa["per_atlas_pixel_face_id"], pr["uv_centers"], pr["uv_scales"],
                      0.05, scales, depth, cfg["edge_dilate_kernels"], True)
    return time.time() - t0


def run_reference(args):
    """--impl reference: the reference algorithm (oracle restatement; the reference itself is
    pure Python + third-party CUDA packages that are not installable here) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from pointdreamer_b200 import synthetic
    cores = torch.get_num_threads()
    cfg = path_config()
    scene = synthetic.make_scene(N_POINTS, seed=0, atlas_res=ATLAS_RES)
    t_geom_view = cpu_geometry_seconds(scene, cfg)
    from oracle import unet as ounet
    from pointdreamer_b200.unet import random_state_dict
    sd = random_state_dict(ounet.DEFAULT_CONFIG, seed=1234, device="cpu")
    o = ounet.UNetOracle(sd, ounet.DEFAULT_CONFIG, emulate_fp16=False)
    x = torch.randn(1, 3, RES, RES)
    t = torch.tensor([500.0])
    for _ in range(min(args.warmup, 1)):
        o.forward(x, t)
    times = []
    for _ in range(args.steps):
        t0 = time.time()
        o.forward(x, t)
        times.append(time.time() - t0)
    t_fwd = sum(times) / len(times)
    per_shape = V * T_STEPS * t_fwd + V * t_geom_view
    value = 1.0 / per_shape
    sample = (f"each step = 1 U-Net forward (batch 1, fp32, 256^2) of the {V * T_STEPS} per shape; "
              f"+ oracle project/unproject of 1 of {V} views timed once ({t_geom_view:.1f} s); "
              f"shapes/s extrapolated = 1/({V * T_STEPS}*t_fwd + {V}*t_geom_view)")
    print(json.dumps({
        "impl": "reference", "metric": "shapes/sec", "value": value, "unit": "shapes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_fwd * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "weights": "random-init ADM 256x256 architecture"},
        "cpu_baseline": {"value": value, "unit": "shapes/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "shapes/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shapes-per-gpu", type=int, default=1,
                    help="shapes per step per GPU (default 1 = BASELINE configs[1]; 8 with --gpus 8 "
                         "= configs[3], all chains of a GPU in one U-Net batch)")
    ap.add_argument("--flow", default="path", choices=["path", "default"],
                    help="path = project->inpaint->unproject (the metric); default = configs/"
                         "default.yaml as shipped, i.e. + neighbour completion + optimize_color")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from pointdreamer_b200 import _lib, demo, synthetic
    from pointdreamer_b200 import dist as pdist
    from pointdreamer_b200.ddnm_inpainting import Inpainter
    from pointdreamer_b200.unet import profile_begin, profile_end

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    cfg = path_config(args.flow)
    # ---- setup (untimed): inputs in pinned host memory, weights, cameras ----
    S = args.shapes_per_gpu
    scene_np = synthetic.make_scene(N_POINTS, seed=rank * S, atlas_res=ATLAS_RES)
    extra_dev = []
    for i in range(1, S):
        e = synthetic.make_scene(N_POINTS, seed=rank * S + i, atlas_res=ATLAS_RES)
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in e.items() if k != "xatlas_dict"}
        d["xatlas_dict"] = {k: torch.from_numpy(v).to(dev) for k, v in e["xatlas_dict"].items()}
        extra_dev.append(d)

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    scene_host = {k: pin(v) for k, v in scene_np.items() if k != "xatlas_dict"}
    scene_host["xatlas_dict"] = {k: pin(v) for k, v in scene_np["xatlas_dict"].items()}
    scene_dev = {k: v.to(dev) for k, v in scene_host.items() if k != "xatlas_dict"}
    xa_dev = {k: v.to(dev) for k, v in scene_host["xatlas_dict"].items()}
    inpainter = Inpainter(dev, seed=42, offset=0, allow_random_weights=True)
    cam_info = demo.prepare_cameras(cfg, dev)
    keys = {k: cfg[k] for k in demo.PATH_CONFIG_KEYS}

    def step_device():
        inpainter.chains_done = 0
        if S > 1:
            first = dict(scene_dev, xatlas_dict=xa_dev)
            atlas = torch.stack(demo.colorize_batch([first] + extra_dev, cam_info, cfg, inpainter, dev))
            if world > 1:
                atlas = pdist.gather_stacked(atlas, world * S)
            return atlas
        out = demo.colorize_one_mesh(scene_dev["xyz"], scene_dev["rgb"], scene_dev["vertices"],
                                     scene_dev["faces"], scene_dev["f_normals"], xa_dev, cam_info,
                                     device=dev, save_img_path=None, inpainter=inpainter,
                                     glctx=None, logger=None, **keys)
        atlas = out[4]
        if world > 1:  # the one collective of the path: assemble every rank's atlas
            atlas = pdist.gather_stacked(atlas[None], world)
        return atlas

    def step_e2e():
        inpainter.chains_done = 0
        atlas_host, h2d, d2h = demo.colorize_from_host(scene_host, cam_info, cfg, inpainter, dev)
        return atlas_host, h2d, d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- leg 1: inputs resident in HBM ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    profile_begin(inpainter.model, every=25, max_forwards=4 * args.steps + 4)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        atlas = step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - l0
    prof, n_fwd = profile_end(inpainter.model)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * S * 1000.0 / ms_per_step

    # ---- leg 2: end to end through the public API with HOST buffers ----
    step_e2e()
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        atlas_host, h2d, d2h = step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.time() - t0) * 1e3) / args.steps
    e2e_value = world * 1000.0 / e2e_ms if S == 1 else None  # the host-buffer leg runs one shape

    # ---- per-stage breakdown of one more (untimed) step: CUDA events at the stage boundaries ----
    stage_ms = None
    if S == 1:
        evs = []
        inpainter.chains_done = 0
        demo.colorize_one_mesh(scene_dev["xyz"], scene_dev["rgb"], scene_dev["vertices"],
                               scene_dev["faces"], scene_dev["f_normals"], xa_dev, cam_info,
                               device=dev, save_img_path=None, inpainter=inpainter, glctx=None,
                               logger=None, stage_events=evs, **keys)
        torch.cuda.synchronize()
        stage_ms = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(evs[:-1], evs[1:])}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_tf, peak_hbm, peak_src = measured_peaks()
    conv = prof["conv_tc"]
    conv_tflops = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    total_prof_ms = sum(v["ms"] for v in prof.values())
    traffic, traffic_src = None, None
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_conv_traffic.json")))
    tp = cands[-1] if cands else ""
    if tp:  # DRAM bytes per conv launch from the newest committed ncu capture
        tj = json.load(open(tp))
        traffic, traffic_src = tj["avg_traffic_bytes_per_launch"], tj["source"]
    roofline = {
        "bound": "tensor", "achieved": conv_tflops, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": conv_tflops / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
        "kernel": "conv_halo_kernel / conv_tc2_kernel / conv_tc_kernel (tcgen05 implicit-GEMM conv: halo-tile 3x3, 2-CTA and 1-CTA)",
        "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
        "launch_avg_ms": conv["ms"] / max(conv["launches"], 1),
        "flops_per_launch_avg": conv["flops"] / max(conv["launches"], 1),
        "share_of_unet_time": conv["ms"] / total_prof_ms if total_prof_ms else None,
        "sampled_forwards": n_fwd,
        "per_class_ms_per_forward": {k: v["ms"] / max(n_fwd, 1) for k, v in prof.items()},
        "whole_path_frac_of_tensor_roofline":
            (S * V * T_STEPS * F_UNET / (ms_per_step * 1e-3) / 1e12) / peak_tf,
    }
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        t_fwd, all_t = cpu_unet_forward_seconds(2)
        t_geom = cpu_geometry_seconds(scene_np, cfg)
        per_shape = V * T_STEPS * t_fwd + V * t_geom
        cpu_baseline = {
            "value": 1.0 / per_shape, "unit": "shapes/s", "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": (f"2 oracle U-Net forwards (fp32, batch 1, 256^2: {t_fwd:.2f} s best) of the "
                       f"{V * T_STEPS} per shape + oracle project/unproject of 1 of {V} views "
                       f"({t_geom:.1f} s); extrapolated to a whole shape")}
    line = {
        "metric": "shapes/sec", "value": value, "unit": "shapes/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "shapes_per_step_per_gpu": S, "flow": args.flow,
                   "weights": "random-init ADM 256x256 architecture (552.8M params)",
                   "point_validation_by_o3d": cfg["point_validation_by_o3d"],
                   "complete_unseen_by": cfg["complete_unseen_by"],
                   "optimize_from": cfg["optimize_from"], "edge_dilate_kernels": cfg["edge_dilate_kernels"],
                   "l2": "working set (1.9 GB U-Net arena per step) is far larger than the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "shapes/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "stage_ms": stage_ms,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
