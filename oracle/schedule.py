"""ORACLE (test infrastructure, not product code) -- noise schedule, respacing and
posterior tables of the reference sampler, restated in numpy fp64.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.

Follows (all paths relative to /root/reference):
  * nicediffusion/diffusion.py:445-475   get_beta_schedule
  * nicediffusion/diffusion.py:87-130    Diffusion.__init__ (respacing + tables)

Pinned against the reference itself by ``oracle/make_golden.py`` (the reference
has no tests or golden vectors of its own -- SURVEY.md section 4) and against the
known-answer values listed in SURVEY.md section 8(c).
"""
import math

import numpy as np


def beta_schedule(method, num_steps, beta_0, beta_T):
    """diffusion.py:445-475. linear / constant / cosine (IDDPM eq. 17, clipped at 0.999)."""
    if method == 'linear':
        return np.linspace(beta_0, beta_T, num_steps, dtype=np.float64)
    if method == 'constant':
        return beta_0 * np.ones(num_steps, dtype=np.float64)
    if method == 'cosine':
        def f(u):
            return math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2
        out = []
        for s in range(num_steps):
            out.append(min(1 - f((s + 1) / num_steps) / f(s / num_steps), 0.999))
        return np.array(out)
    raise NotImplementedError("unimplemented variance scheduling method: {}".format(method))


class Tables:
    """All per-step constants the sampler reads (diffusion.py:94-130)."""

    def __init__(self, original_num_steps, rescaled_num_steps, beta_schedule_name='linear', betas=None):
        T, R = original_num_steps, rescaled_num_steps
        if betas is None:
            # endpoints scaled by 1000/T (diffusion.py:88-89)
            betas = beta_schedule(beta_schedule_name, T, 0.0001 * 1000 / T, 0.02 * 1000 / T)
        else:
            assert len(betas) == T, 'betas must be the right length!'
            betas = np.array(betas, dtype=np.float64)

        abar_full = np.cumprod(1.0 - betas, axis=0)
        # diffusion.py:97-99 -- note: when R does not divide T the list is NOT of length R (kept as is)
        kept = list(range(T // (2 * R), T + T // (2 * R), T // R))
        kept_set = set(kept)
        last = 1.0
        new_betas = []
        for i, a in enumerate(abar_full):
            if i in kept_set:
                new_betas.append(1.0 - a / last)
                last = a
        betas = np.array(new_betas)
        assert (betas > 0).all() and (betas <= 1).all(), 'betas in invalid range'

        self.timestep_map = np.array(kept, dtype=np.int64)
        self.betas = betas
        alphas = 1.0 - betas
        self.alphas_cumprod = abar = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = abar_prev = np.append(1.0, abar[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(abar)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - abar)
        self.sqrt_reciprocal_alphas_cumprod = np.sqrt(1.0 / abar)
        self.sqrt_reciprocal_alphas_minus_one_cumprod = np.sqrt(1.0 / abar - 1)
        self.posterior_mean_coef_x0 = np.sqrt(abar_prev) * betas / (1.0 - abar)
        self.posterior_mean_coef_xt = np.sqrt(alphas) * (1.0 - abar_prev) / (1.0 - abar)
        self.posterior_variance = betas * (1.0 - abar_prev) / (1.0 - abar)
        self.log_posterior_var_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))

    NAMES = ('betas', 'alphas_cumprod', 'alphas_cumprod_prev', 'sqrt_alphas_cumprod',
             'sqrt_one_minus_alphas_cumprod', 'sqrt_reciprocal_alphas_cumprod',
             'sqrt_reciprocal_alphas_minus_one_cumprod', 'posterior_mean_coef_x0',
             'posterior_mean_coef_xt', 'posterior_variance', 'log_posterior_var_clipped')
