"""ORACLE (test infrastructure, not product code) -- CPU restatement of the reference sampler.

``model`` is any callable ``model(x_t, timestep_original_index, y=...) -> [B, out_ch, H, W]``
(normally a closure over oracle.unet.unet_forward). Noise is always *injected* (``noise_fn``)
so trajectories can be compared step by step with the CUDA path on identical noise.

Follows (paths relative to /root/reference):
  * nicediffusion/diffusion.py:155-226  denoise (loop order, timestep tensor)
  * nicediffusion/diffusion.py:242-264  get_eps_and_log_var (4 variance types)
  * nicediffusion/diffusion.py:266-316  denoising_step (DDPM, CFG, classifier guidance)
  * nicediffusion/diffusion.py:318-369  ddim_denoising_step
  * nicediffusion/diffusion.py:232-240  diffusion_step (q-sample)
  * nicediffusion/diffusion.py:478-496  extract (fp64 table -> fp32 gather)
"""
import numpy as np
import torch

from .schedule import Tables


def extract(a, t, shape):
    """diffusion.py:478-496: the fp64 table is cast to fp32 *before* the gather."""
    r = torch.gather(torch.from_numpy(np.asarray(a)).float(), 0, t.long())
    while r.dim() < len(shape):
        r = r[..., None]
    return r.expand(shape)


class Sampler:
    def __init__(self, model, original_num_steps, rescaled_num_steps, sampling_var_type,
                 beta_schedule='linear', betas=None, guidance_method=None, guidance_strength=None,
                 classifier=None, use_ddim=False, ddim_eta=None):
        if guidance_method not in (None, 'classifier', 'classifier_free'):
            raise NotImplementedError(guidance_method)
        if sampling_var_type not in ('small', 'large', 'learned', 'learned_interpolation'):
            raise NotImplementedError(sampling_var_type)
        if use_ddim:
            assert ddim_eta is not None, 'please supply eta if you want to use ddim'
        self.model = model
        self.tab = Tables(original_num_steps, rescaled_num_steps, beta_schedule, betas)
        self.timestep_map = torch.from_numpy(self.tab.timestep_map)
        self.rescaled_num_steps = rescaled_num_steps
        self.var_type = sampling_var_type
        self.guidance = guidance_method
        self.strength = guidance_strength
        self.classifier = classifier
        self.use_ddim = use_ddim
        self.ddim_eta = ddim_eta

    # diffusion.py:232-240
    def diffusion_step(self, x_0, t, noise):
        return extract(self.tab.sqrt_alphas_cumprod, t, x_0.shape) * x_0 + \
            extract(self.tab.sqrt_one_minus_alphas_cumprod, t, x_0.shape) * noise

    # diffusion.py:132-153
    def diffuse(self, x_0, steps_to_do=None, noise=None):
        if steps_to_do is None or steps_to_do > self.rescaled_num_steps:
            steps_to_do = self.rescaled_num_steps
        t = (steps_to_do - 1) * torch.ones(x_0.shape[0])
        return self.diffusion_step(x_0, t, noise)

    # diffusion.py:242-264
    def eps_and_log_var(self, x_t, t, kwargs):
        tab = self.tab
        out = self.model(x_t, self.timestep_map[t.long()], **kwargs)
        if self.var_type == 'learned':
            eps, log_var = torch.split(out, int(out.shape[1] / 2), dim=1)
        elif self.var_type == 'learned_interpolation':
            eps, log_var = torch.split(out, int(out.shape[1] / 2), dim=1)
            min_log = extract(tab.log_posterior_var_clipped, t, x_t.shape)
            max_log = extract(np.log(tab.betas), t, x_t.shape)
            frac = (log_var + 1) / 2
            log_var = frac * max_log + (1 - frac) * min_log
        elif self.var_type == 'large':
            eps = out
            log_var = extract(np.log(np.append(tab.posterior_variance[1], tab.betas[1:])), t, x_t.shape)
        else:  # small
            eps = out
            log_var = extract(np.log(np.maximum(tab.posterior_variance, 1e-20)), t, x_t.shape)
        return eps, log_var

    def _null_eps(self, x_t, t, n):
        base = self.model(x_t, self.timestep_map[t.long()], y=torch.zeros(n, dtype=torch.long))
        if self.var_type in ('learned', 'learned_interpolation'):
            base, _ = torch.split(base, int(base.shape[1] / 2), dim=1)
        return base

    def _classifier_grad(self, x_t, t, y):
        with torch.enable_grad():
            x = x_t.detach().requires_grad_(True)
            lp = torch.log_softmax(self.classifier(x, t), dim=-1)
            grabbed = lp[range(lp.shape[0]), torch.flatten(y)]
            return torch.autograd.grad(grabbed.sum(), x)[0]

    # diffusion.py:266-316
    def ddpm_step(self, x_t, t, kwargs, noise, clip_x=True):
        tab = self.tab
        eps, log_var = self.eps_and_log_var(x_t, t, kwargs)
        if self.guidance == 'classifier_free':
            base = self._null_eps(x_t, t, eps.shape[0])
            eps = (1 + self.strength) * eps - self.strength * base
        x0 = extract(tab.sqrt_reciprocal_alphas_cumprod, t, x_t.shape) * x_t - \
            extract(tab.sqrt_reciprocal_alphas_minus_one_cumprod, t, eps.shape) * eps
        if clip_x:
            x0 = torch.clamp(x0, -1, 1)
        mean = extract(tab.posterior_mean_coef_x0, t, x0.shape) * x0 + \
            extract(tab.posterior_mean_coef_xt, t, x_t.shape) * x_t
        if self.guidance == 'classifier':
            # NB the classifier receives the *respaced* index t (diffusion.py:301)
            grad = self._classifier_grad(x_t, t, kwargs['y'])
            mean = mean + self.strength * grad * torch.exp(log_var)
        mask = (1.0 - (t == 0).float()).reshape((x_t.shape[0],) + (1,) * (x_t.dim() - 1))
        sample = mean + mask * torch.exp(0.5 * log_var) * noise
        return sample.float(), x0

    # diffusion.py:318-369
    def ddim_step(self, x_t, t, kwargs, noise, clip_x=True):
        tab = self.tab
        eps = self.model(x_t, self.timestep_map[t.long()], **kwargs)
        if self.var_type in ('learned', 'learned_interpolation'):
            eps, _ = torch.split(eps, int(eps.shape[1] / 2), dim=1)
        if self.guidance == 'classifier':
            grad = self._classifier_grad(x_t, t, kwargs['y'])
            eps = eps - self.strength * grad * extract(tab.sqrt_one_minus_alphas_cumprod, t, x_t.shape)
        elif self.guidance == 'classifier_free':
            base = self._null_eps(x_t, t, eps.shape[0])
            eps = (1 + self.strength) * eps - self.strength * base
        x0 = extract(tab.sqrt_reciprocal_alphas_cumprod, t, x_t.shape) * x_t - \
            extract(tab.sqrt_reciprocal_alphas_minus_one_cumprod, t, x_t.shape) * eps
        if clip_x:
            x0 = torch.clamp(x0, -1, 1)
        abar = extract(tab.alphas_cumprod, t, x_t.shape)
        abar_prev = extract(tab.alphas_cumprod_prev, t, x_t.shape)
        var = self.ddim_eta ** 2 * (1.0 - abar_prev) * (1.0 - abar / abar_prev) / (1.0 - abar)
        # raw eps (not eps re-derived from the clipped x0) -- diffusion.py:360
        mean = x0 * torch.sqrt(abar_prev) + torch.sqrt(1 - abar_prev - var) * eps
        mask = (1.0 - (t == 0).float()).reshape((x_t.shape[0],) + (1,) * (x_t.dim() - 1))
        sample = mean + mask * torch.sqrt(var) * noise
        return sample.float(), x0

    # diffusion.py:155-226
    @torch.no_grad()
    def denoise(self, x, kwargs, noise_fn, start_step=None, steps_to_do=None, trajectory=None):
        """noise_fn(step_index_t) -> noise tensor shaped like x. ``trajectory`` (list) collects
        (t, x_{t-1}, pred_x0) per step when given."""
        if start_step is None:
            start_step = self.rescaled_num_steps
        if steps_to_do is None or steps_to_do > start_step:
            steps_to_do = start_step
        for t in reversed(range(steps_to_do)):
            timestep = t * torch.ones(x.shape[0])
            step = self.ddim_step if self.use_ddim else self.ddpm_step
            x, x0 = step(x, timestep, kwargs, noise_fn(t))
            if trajectory is not None:
                trajectory.append((t, x.clone(), x0.clone()))
        return x
