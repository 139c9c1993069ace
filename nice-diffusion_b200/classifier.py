"""`NoisyClassifier` -- a noisy-image classifier p(y | x_t, t) for classifier guidance.

The reference defines the interface only -- `classifier(x, t) -> [B, num_classes]` logits, called with the RESPACED step
index as a float tensor (nicediffusion/diffusion.py:72-75, 298-305, 330-337) -- ships no model and its CLI refuses one
(utils.py:171-172). This is the ADM-style encoder: the UNet's down path and middle block (the same parameter containers,
so OpenAI classifier checkpoints map onto it the way UNet checkpoints do) followed by GroupNorm - SiLU - attention pooling.

Guidance needs d log p(y | x_t) / d x_t every step, i.e. a backward pass through this network with respect to its INPUT.
It is evaluated with differentiable torch ops (train_forward.py's block functions) -- exactly the kind of module the
reference API accepts -- and `Diffusion.denoise` captures its forward + backward inside the per-step CUDA graph, between
the UNet plan and the fused sampler kernel that consumes the gradient (diffusion._StepRunner).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .model import AttentionBlock, ResidualBlock, UsesStepsSequential
from .train_forward import _attn, _res, _sinusoid


class AttentionPool2d(nn.Module):
    """Attention pooling (CLIP / ADM): the mean token attends over all positions; its output is the feature."""

    def __init__(self, spatial, channels, num_head_channels, out_channels):
        super().__init__()
        self.positional_embedding = nn.Parameter(torch.randn(channels, spatial ** 2 + 1) / channels ** 0.5)
        self.qkv_proj = nn.Conv1d(channels, 3 * channels, 1)
        self.c_proj = nn.Conv1d(channels, out_channels, 1)
        self.num_heads = channels // num_head_channels

    def forward(self, x):
        B, C = x.shape[:2]
        x = x.reshape(B, C, -1)
        x = torch.cat([x.mean(dim=-1, keepdim=True), x], dim=-1) + self.positional_embedding[None].to(x.dtype)
        q, k, v = self.qkv_proj(x).reshape(B * self.num_heads, 3, C // self.num_heads, -1).unbind(1)
        scale = 1 / math.sqrt(math.sqrt(q.shape[1]))
        w = torch.softmax(torch.einsum('bct,bcs->bts', q * scale, k * scale), dim=-1)
        a = torch.einsum('bts,bcs->bct', w, v).reshape(B, C, -1)
        return self.c_proj(a)[:, :, 0]


class NoisyClassifier(nn.Module):
    def __init__(self, resolution, in_channels, model_channels, num_classes, num_res_blocks, attention_resolutions,
                 channel_mult=(1, 2, 3, 4), num_head_channels=64, dropout=0.0, use_adaptive_gn=True, steps_scale=1.0):
        super().__init__()
        self.resolution, self.in_channels, self.model_channels = resolution, in_channels, model_channels
        self.num_classes = num_classes
        # the sampler hands over the respaced index t (diffusion.py:301); a classifier trained on original timesteps
        # rescales it here (steps_scale = original_num_steps / rescaled_num_steps)
        self.steps_scale = steps_scale
        E = 4 * model_channels
        self.step_embed = nn.Sequential(nn.Linear(model_channels, E), nn.SiLU(), nn.Linear(E, E))
        cur = int(model_channels * channel_mult[0])
        res = resolution
        self.downsampling = nn.ModuleList([UsesStepsSequential(nn.Conv2d(in_channels, cur, 3, padding=1))])
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResidualBlock(cur, E, dropout, out_channels=int(model_channels * mult), use_adaptive_gn=use_adaptive_gn)]
                cur = int(model_channels * mult)
                if res in attention_resolutions:
                    layers.append(AttentionBlock(cur, num_head_channels=num_head_channels))
                self.downsampling.append(UsesStepsSequential(*layers))
            if level != len(channel_mult) - 1:
                self.downsampling.append(UsesStepsSequential(
                    ResidualBlock(cur, E, dropout, downsample=True, out_channels=cur, use_adaptive_gn=use_adaptive_gn)))
                res //= 2
        self.middle_block = UsesStepsSequential(
            ResidualBlock(cur, E, dropout, use_adaptive_gn=use_adaptive_gn), AttentionBlock(cur, num_head_channels=num_head_channels),
            ResidualBlock(cur, E, dropout, use_adaptive_gn=use_adaptive_gn))
        self.out = nn.Sequential(nn.GroupNorm(32, cur), nn.SiLU(), AttentionPool2d(res, cur, num_head_channels, num_classes))
        # the encoder's residual branches start at zero (zero_module); give the head something to classify from the start
        for mod in self.modules():
            if isinstance(mod, ResidualBlock):
                nn.init.normal_(mod.out_conv.weight, std=(9 * mod.out_channels) ** -0.5)
            elif isinstance(mod, AttentionBlock):
                nn.init.normal_(mod.proj_out.weight, std=mod.channels ** -0.5)

    def forward(self, x, t):
        emb = self.step_embed(_sinusoid(t.float() * self.steps_scale, self.model_channels))
        h = x
        for seq in list(self.downsampling) + [self.middle_block]:
            for layer in seq:
                if isinstance(layer, ResidualBlock):
                    h = _res(layer, h, emb, self.training)
                elif isinstance(layer, AttentionBlock):
                    h = _attn(layer, h)
                else:
                    h = layer(h)
        return self.out(h)
