"""Oracle (TEST INFRASTRUCTURE): DDNM inpainting sampler.

Restates models/DDNM/guided_diffusion/diffusion.py:459-570 (simplified_ddnm_inpainting),
:770-791 (get_schedule_jump), :809-812 (compute_alpha), :46-76 (get_beta_schedule, linear) and
models/DDNM/datasets/__init__.py:208-234 (data_transform / inverse_data_transform) with the
constants of models/DDNM/configs/imagenet_256.yml and ddnm_inpainting.py:20-24
(sigma_y = 0, eta = 0.85).  Quirks kept on purpose: gamma_t uses
alpha_bar SQUARED; a noise tensor is drawn on the last step although its weight is 0.

Pinned against the reference's own sampler executed on CPU (tests/golden/make_golden_ddnm.py).
"""
import numpy as np
import torch

F32 = np.float32


def get_schedule_jump(T_sampling, travel_length, travel_repeat):
    """diffusion.py:770-791."""
    jumps = {}
    for j in range(0, T_sampling - travel_length, travel_length):
        jumps[j] = travel_repeat - 1
    t = T_sampling
    ts = []
    while t >= 1:
        t = t - 1
        ts.append(t)
        if jumps.get(t, 0) > 0:
            jumps[t] = jumps[t] - 1
            for _ in range(travel_length):
                t = t + 1
                ts.append(t)
    ts.append(-1)
    return ts


def alphas_cumprod_table(beta_start=1e-4, beta_end=0.02, n=1000):
    """compute_alpha's table: cumprod(1 - cat([0], betas)) with betas float32.
    (torch's CPU cumprod accumulates float32 inputs in float64 and rounds every output.)"""
    betas = np.linspace(beta_start, beta_end, n, dtype=np.float64).astype(F32)
    one_minus = (F32(1) - np.concatenate([np.zeros(1, F32), betas])).astype(F32)
    return np.cumprod(one_minus.astype(np.float64)).astype(F32)  # index t+1


def step_table(T_sampling=100, num_timesteps=1000, eta=0.85, sigma_y=0.0, travel_length=1,
               travel_repeat=1, beta_start=1e-4, beta_end=0.02):
    """Per-step (t, coefficients) exactly as the sampler's loop computes them in fp32.
    Returns ts float32 [steps] and coef float32 [steps,7] =
    (sqrt(1-at), sqrt(at), sqrt(at_next), gamma_t, c1, c2, lambda_t)."""
    if travel_repeat != 1:
        raise NotImplementedError("time-travel (travel_repeat > 1) is not used by the path")
    skip = num_timesteps // T_sampling
    times = get_schedule_jump(T_sampling, travel_length, travel_repeat)
    acp = alphas_cumprod_table(beta_start, beta_end, num_timesteps)
    sigma_y = F32(2 * sigma_y)  # diffusion.py:469
    ts, coefs = [], []
    for i, j in zip(times[:-1], times[1:]):
        i, j = i * skip, j * skip
        if j < 0:
            j = -1
        assert j < i
        at = acp[i + 1]
        at_next = acp[j + 1]
        sigma_t = np.sqrt(F32(1) - at_next ** F32(2), dtype=F32)
        if sigma_t >= at_next * sigma_y:
            lambda_t = F32(1.0)
            gamma_t = np.sqrt(sigma_t ** F32(2) - (at_next * sigma_y) ** F32(2), dtype=F32)
        else:
            lambda_t = F32(sigma_t / (at_next * sigma_y))
            gamma_t = F32(0.0)
        c1 = np.sqrt(F32(1) - at_next, dtype=F32) * F32(eta)
        c2 = np.sqrt(F32(1) - at_next, dtype=F32) * F32((1 - eta ** 2) ** 0.5)
        ts.append(F32(i))
        coefs.append([np.sqrt(F32(1) - at, dtype=F32), np.sqrt(at, dtype=F32),
                      np.sqrt(at_next, dtype=F32), gamma_t, c1, c2, lambda_t])
    return np.asarray(ts, dtype=F32), np.asarray(coefs, dtype=F32)


def ddnm_step(xt, et, y, mask, c, noise):
    """One reverse step, numpy float32, one rounding per reference op (diffusion.py:533-552).
    xt, et, y, noise [V,3,H,W]; mask [V,1,H,W] or broadcastable; c = 7 coefficients."""
    s1m, sat, satn, gamma, c1, c2, lam = [F32(v) for v in c]
    x0_t = (xt - et * s1m) / sat
    x0_hat = x0_t - lam * ((x0_t * mask - y) * mask)
    return (satn * x0_hat + gamma * (c1 * noise + c2 * et)).astype(F32)


def sample(model_fn, sparse, mask, noise_fn, T_sampling=100, num_timesteps=1000, eta=0.85):
    """Full chain for V views (each view an independent chain).

    model_fn(x [V,3,H,W] float32 numpy, t [V] float32) -> eps [V,>=3,H,W]
    sparse [V,3,H,W] in [0,1]; mask [V,H,W] (1 = known)
    noise_fn(chain v, draw d, shape) -> standard normal draw d of chain v (d=0: x_T, 1+s: step s)
    Returns [V,3,H,W] in [0,1]."""
    V = sparse.shape[0]
    ts, coefs = step_table(T_sampling, num_timesteps, eta)
    m = mask[:, None].astype(F32)
    x_orig = (F32(2) * sparse.astype(F32) - F32(1)).astype(F32)
    y = x_orig * m
    x = np.stack([noise_fn(v, 0, sparse.shape[1:]) for v in range(V)]).astype(F32)
    for s in range(len(ts)):
        t = np.full((V,), ts[s], dtype=F32)
        et = np.asarray(model_fn(x, t), dtype=F32)[:, :3]
        noise = np.stack([noise_fn(v, 1 + s, sparse.shape[1:]) for v in range(V)]).astype(F32)
        x = ddnm_step(x, et, y, m, coefs[s], noise)
    return np.clip((x + F32(1)) / F32(2), F32(0), F32(1)).astype(F32)


def torch_cpu_noise_stream(seed, V, draws_per_chain, shape):
    """The reference's draw order on ONE generator: chain after chain, x_T then one per step."""
    g = torch.Generator().manual_seed(seed)
    table = {}
    for v in range(V):
        for d in range(draws_per_chain):
            table[(v, d)] = torch.randn(1, *shape, generator=g)[0].numpy()
    return lambda v, d, shp: table[(v, d)]
