"""ORACLE (test infrastructure, not product code) -- CPU restatement of the reference's training loss.

Follows (paths relative to /root/reference):
  * nicediffusion/diffusion.py:375-406  Diffusion.loss (simple / KL / KL_rescaled / hybrid)
  * nicediffusion/diffusion.py:408-438  variational_lower_bound
  * nicediffusion/diffusion.py:499-508  kl_div
  * nicediffusion/diffusion.py:511-519  approx_cdf (Page 1977)
  * nicediffusion/diffusion.py:522-543  log_likelihood (discretised Gaussian)
  * nicediffusion/diffusion.py:546-549  mean_flat
Pinned bit-exact against the reference by oracle/make_golden.py (`loss` cases -> tests/golden/loss_*.npz).
"""
import math

import numpy as np
import torch

from .sampler import Sampler, extract


def mean_flat(x):
    return torch.mean(x, dim=list(range(1, x.dim())))


def kl_div(mean_1, log_var_1, mean_2, log_var_2):
    return ((log_var_2 - log_var_1) + torch.exp(log_var_1 - log_var_2) +
            ((mean_1 - mean_2) ** 2) * torch.exp(-log_var_2) - 1.0) / 2


def approx_cdf(x):
    y = math.sqrt(2.0 / math.pi) * (x + 0.0444715 * (x ** 3))
    return 0.5 * (1.0 + torch.tanh(y))


def log_likelihood(target, mean, log_var):
    std_recip = torch.exp(-0.5 * log_var)
    centered = target - mean
    plus = (centered + 1.0 / 255.0) * std_recip
    minus = (centered - 1.0 / 255.0) * std_recip
    cdf_minus, cdf_plus = approx_cdf(minus), approx_cdf(plus)
    cdf_delta = cdf_plus - cdf_minus
    log_cdf_plus = torch.log(cdf_plus.clamp(min=1e-12))
    log_one_minus_cdf_minus = torch.log((1.0 - cdf_minus).clamp(min=1e-12))
    return torch.where(target < -0.999, log_cdf_plus,
                       torch.where(target > 0.999, log_one_minus_cdf_minus, torch.log(cdf_delta.clamp(min=1e-12))))


class LossOracle(Sampler):
    """Sampler tables + the loss; `model` must be differentiable if gradients are wanted."""

    def __init__(self, model, loss_type, **kw):
        super().__init__(model, **kw)
        assert loss_type in ('simple', 'KL', 'KL_rescaled', 'hybrid')
        self.loss_type = loss_type

    def vlb(self, x_0, x_t, t, eps, log_var):
        tab = self.tab
        true_mean = extract(tab.posterior_mean_coef_x0, t, x_0.shape) * x_0 + \
            extract(tab.posterior_mean_coef_xt, t, x_t.shape) * x_t
        true_log_var = extract(tab.log_posterior_var_clipped, t, x_0.shape)
        pred_x0 = extract(tab.sqrt_reciprocal_alphas_cumprod, t, x_t.shape) * x_t - \
            extract(tab.sqrt_reciprocal_alphas_minus_one_cumprod, t, eps.shape) * eps
        mean = extract(tab.posterior_mean_coef_x0, t, pred_x0.shape) * pred_x0 + \
            extract(tab.posterior_mean_coef_xt, t, x_t.shape) * x_t
        kl = mean_flat(kl_div(true_mean, true_log_var, mean, log_var)) / np.log(2.0)
        nll = mean_flat(-log_likelihood(x_0, mean, log_var)) / np.log(2.0)
        return torch.where((t == 0), nll, kl)

    def loss_from_output(self, x_0, x_t, t, out, noise):
        """The loss given the raw model output (what the fused kernel is compared with)."""
        tab = self.tab
        if self.var_type in ('learned', 'learned_interpolation'):
            eps, v = torch.split(out, int(out.shape[1] / 2), dim=1)
            if self.var_type == 'learned':
                log_var = v
            else:
                frac = (v + 1) / 2
                log_var = frac * extract(np.log(tab.betas), t, x_t.shape) + \
                    (1 - frac) * extract(tab.log_posterior_var_clipped, t, x_t.shape)
        else:
            eps = out
            fixed = np.log(np.append(tab.posterior_variance[1], tab.betas[1:])) if self.var_type == 'large' else \
                np.log(np.maximum(tab.posterior_variance, 1e-20))
            log_var = extract(fixed, t, x_t.shape)
        if self.loss_type == 'simple':
            return mean_flat((eps - noise) ** 2)
        if self.loss_type in ('KL', 'KL_rescaled'):
            loss = self.vlb(x_0, x_t, t, eps, log_var)
            return loss * self.rescaled_num_steps if self.loss_type == 'KL_rescaled' else loss
        loss_simple = mean_flat((eps - noise) ** 2)
        loss_vlb = self.vlb(x_0, x_t, t, eps.detach(), log_var) * self.rescaled_num_steps
        return loss_simple + 0.001 * loss_vlb

    def loss(self, x_0, t, kwargs, noise):
        x_t = self.diffusion_step(x_0, t, noise)
        out = self.model(x_t, self.timestep_map[t.long()], **kwargs)
        return self.loss_from_output(x_0, x_t, t, out, noise)
