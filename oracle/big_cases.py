"""ORACLE (test infrastructure) -- the BASELINE-size trajectory cases: definitions + seeded inputs.

Shared by oracle/make_golden.py (which runs the reference on them in the build container) and by
tests/test_sampler_gpu.py (which regenerates the same inputs from the seed on the GPU box, so the fixtures only
hold the reference's OUTPUTS). numpy's legacy RandomState stream is bit-stable across versions.

A case runs `n_steps` consecutive steps t = t_first, t_first-1, ... of the reference's step function
(nicediffusion/diffusion.py:266-369):
  * start == 'noise'   : x = x_T ~ N(0,1) at t_first = R-1 -- the steps where sqrt(1/abar_t) is largest (641 for the
                         cosine schedule at R=250) and 16-bit operand error is amplified most before the clamp;
  * start == 'diffused': x = q-sample of a synthetic image to step t_first (diffusion.py:132-153), then
                         `denoise(x, steps_to_do=t_first+1)` to t = 0 -- the partial-denoise / --start_img path.
"""
import numpy as np
import torch

BIG_CASES = {
    # name: (config, diffusion kwargs, batch, start, n_steps)
    'in64_ddpm_first10': ('in64', dict(original_num_steps=1000, rescaled_num_steps=250, beta_schedule='cosine',
                                       sampling_var_type='learned_interpolation'), 1, 'noise', 10),
    'in64_ddpm_last10': ('in64', dict(original_num_steps=1000, rescaled_num_steps=250, beta_schedule='cosine',
                                      sampling_var_type='learned_interpolation'), 1, 'diffused', 10),
    # default_args.py:4-6 (EMNIST preset): cosine, R=25, classifier-free guidance 0.8 -- the whole 25-step chain
    'emnist_cfg_full': ('emnist', dict(original_num_steps=1000, rescaled_num_steps=25, beta_schedule='cosine',
                                       sampling_var_type='learned_interpolation', guidance_method='classifier_free',
                                       guidance_strength=0.8), 2, 'noise', 25),
    # BASELINE config 3: 128x128, linear schedule, DDIM eta=0, R=50, classifier-free guidance
    'in128_ddim_cfg_first10': ('in128', dict(original_num_steps=1000, rescaled_num_steps=50, beta_schedule='linear',
                                             sampling_var_type='learned_interpolation', use_ddim=True, ddim_eta=0.0,
                                             guidance_method='classifier_free', guidance_strength=0.8),
                               1, 'noise', 10),
}

WEIGHT_SEED = 4321


def big_case_inputs(name):
    """-> dict(x_start, noises [n_steps, B, C, H, W] (row i = draw of the i-th executed step), y, t_first, x0_img)."""
    from . import CONFIGS
    cname, dkw, B, start, n_steps = BIG_CASES[name]
    cfg = CONFIGS[cname]
    R = dkw['rescaled_num_steps']
    shape = (B, cfg['in_channels'], cfg['resolution'], cfg['resolution'])
    rs = np.random.RandomState(abs(hash_name(name)) % (2 ** 31))
    noises = torch.from_numpy(rs.standard_normal((n_steps,) + shape).astype(np.float32))
    y = torch.from_numpy(rs.randint(1, cfg['num_classes'], size=(B,)).astype(np.int64))
    out = dict(noises=noises, y=y, x0_img=None)
    if start == 'noise':
        out['x_start'] = torch.from_numpy(rs.standard_normal(shape).astype(np.float32))
        out['t_first'] = R - 1
    else:
        # a smooth synthetic "image" in [-1, 1] plus the q-sample noise; the q-sample itself is done by the caller
        # (reference / oracle / product each with their own diffuse(), which is bit-exact across the three)
        res = cfg['resolution']
        yy, xx = np.meshgrid(np.linspace(-1, 1, res), np.linspace(-1, 1, res), indexing='ij')
        img = np.stack([np.sin(3 * xx + c) * np.cos(2 * yy - c) for c in range(cfg['in_channels'])])
        out['x0_img'] = torch.from_numpy(np.broadcast_to(img, shape).astype(np.float32).copy())
        out['q_noise'] = torch.from_numpy(rs.standard_normal(shape).astype(np.float32))
        out['t_first'] = n_steps - 1
    return out


def hash_name(name):
    """Stable small hash (python's hash() is salted per process)."""
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000003
    return h
