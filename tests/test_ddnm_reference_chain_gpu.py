"""North-star FP parity on the REAL chain: the reference's own, unmodified sampler
(Diffusion.simplified_ddnm_inpainting, diffusion.py:459-570) + UNetModel (unet.py:396-664) in
stock PyTorch fp16/cuDNN - loaded from baseline/_ref through oracle/reference_ddnm.py - next to
Inpainter.inpaint_batch on the same B200: 100 steps, 256^2, the full 552.8M-parameter model, 8 views
of the synthetic bench scene, same weights, same noise stream (seed 42, offset 0).

What was measured on B200 (profiles/r02_chain_parity.json; `tools/chain_parity_report.py` also runs
the reference in fp32 and with cudnn.benchmark=True, which is too slow for the test-suite):
  * x_T and every known pixel: bit-identical to the reference.
  * ONE forward on the reference chain's own x_t (teacher forced, all 100 steps x 8 views):
        ours vs reference-fp16  max-abs 2.6e-3 (eps reaches |3.6|, std 0.33);
        reference-fp32 vs reference-fp16  1.6e-3 .. 2.7e-3   -> the engine sits inside the
        reference's own fp16 rounding envelope.
  * the free-running chain is chaotic in fp16 rounding (random-init weights are not a denoiser):
        max|x_t - x_t(ref16)| grows 0 -> 0.58 for ours and 0 -> 1.15 for the reference's own fp32 run;
        final images vs ref16:  ours  mean-abs 2.1e-5, PSNR 58.8 dB, 5.6e-4 of the pixels off by more
        than 1e-3, max-abs 0.22;  reference-fp32  mean-abs 3.5e-5, PSNR 54.7 dB, 5.9e-4, max-abs 0.30.
        The reference re-run with cudnn.benchmark=True is bit-identical to itself.
  So the north-star bound (1e-3 abs per channel) holds for 99.94 % of the pixels and cannot hold for
  the rest for ANY implementation that is not bit-identical to cuDNN's fp16 convolutions - the
  reference's own fp32 run misses it by more.  The asserts below are 1.5 x the observed values.
"""
import json
import os

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not ref_loader.available(),
                    reason="baseline/_ref not populated (python -m oracle.make_ref in the build container)")
def test_full_chain_vs_reference_sampler(cuda):
    from chain_parity import study
    from pointdreamer_b200.unet import DEFAULT_MODEL_CONFIG
    r = study(cuda, dict(DEFAULT_MODEL_CONFIG), n_views=8, T=100, with_fp32=False, with_benchmark=False)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(r, open(os.path.join(out_dir, "chain_parity_test.json"), "w"), indent=1)
    f = r["final_ours_vs_ref16"]
    tf = max(r["teacher_forced_ours_vs_ref16"])
    print(f"reference {r['reference_fp16_seconds']:.1f} s, ours {r['ours_seconds']:.2f} s; teacher-forced "
          f"max-abs {tf:.3e}; final max-abs {f['max_abs']:.3e} mean-abs {f['mean_abs']:.3e} "
          f"PSNR {f['psnr_db']:.1f} dB, > 1e-3: {f['frac_gt_1e3']:.2e}; drift end {r['drift_ours'][-1]:.3f}")
    # exact parts
    assert r["x_T_equal"] and r["stepwise_equals_single_call"]
    assert r["ours_known_pixel_max_err"] <= 1.5e-8 and r["reference_known_pixel_max_err"] <= 1.5e-8
    assert r["drift_ours"][0] == 0.0
    # one forward on identical inputs, every step of the real chain: observed max 2.59e-3
    assert tf < 3.9e-3
    # free-running chain, final images: observed max-abs 0.223, mean-abs 2.08e-5, PSNR 58.8 dB,
    # 5.65e-4 of the pixels beyond 1e-3 (the reference's own fp32 run: 0.297 / 3.46e-5 / 54.7 / 5.88e-4)
    assert f["max_abs"] < 0.34
    assert f["mean_abs"] < 3.2e-5
    assert f["psnr_db"] > 57.0 and f["psnr_8bit_db"] > 57.0
    assert f["frac_gt_1e3"] < 8.5e-4
    assert r["drift_ours"][-1] < 0.88
