"""Differentiable forward of `DiffusionModel` for the TRAINING step.

The sampling forward (`engine.py`) is a static plan of hand-written kernels with no backward. Training needs gradients
w.r.t. every parameter, so `DiffusionModel.forward` routes here when the module is in training mode with autograd
enabled: a walk over the same parameter containers that evaluates each block with differentiable torch ops, in the
reference's arithmetic (fp32, NCHW; nicediffusion/model.py:451-476 for the trunk, :188-211 ResidualBlock, :260-291
AttentionBlock, :76-80 / :107-112 up / down-sampling, :514-523 timestep embedding). torch's autograd supplies the backward;
the loss itself (forward and gradient) is a fused kernel (`ndb_diffusion_loss`), and the data-parallel gradient exchange is
`trainer.GradBuckets`.

Blocks are recomputed in the backward pass when the model was built with `use_grad_checkpoints=True`
(the reference wraps ResidualBlock that way and AttentionBlock always: model.py:186, 257) -- here through
torch.utils.checkpoint, and only when requested: a B200 has the memory not to.
"""
import math

import torch
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint as _ckpt

from .model import AttentionBlock, Downsample, ResidualBlock, Upsample, timestep_embedding_freqs


_FREQS = {}


def _sinusoid(timestep, dim):
    # the frequency vector is built on the CPU in fp32 like the reference's (model.py:515-517) but uploaded once per
    # (dim, device): the reference's per-call host-to-device copy would also be illegal inside a CUDA graph capture
    key = (dim, timestep.device)
    f = _FREQS.get(key)
    if f is None:
        f = _FREQS[key] = timestep_embedding_freqs(dim).to(timestep.device)
    ang = timestep[:, None].float() * f[None]
    e = torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)
    return F.pad(e, (0, 1, 0, 0)) if dim % 2 == 1 else e


def _res(rb, x, emb, training):
    h = F.silu(rb.in_norm(x))
    if rb.mode == 'up':
        h, x = (F.interpolate(v, scale_factor=2.0, mode='nearest') for v in (h, x))
    elif rb.mode == 'down':
        h, x = (F.avg_pool2d(v, kernel_size=2, stride=2) for v in (h, x))
    h = rb.in_conv(h)
    e = rb.step_embedding(F.silu(emb))[:, :, None, None]
    if rb.use_adaptive_gn:
        scale, shift = e.chunk(2, dim=1)
        h = rb.out_norm(h) * (1 + scale) + shift
    else:
        h = rb.out_norm(h + e)
    h = F.dropout(F.silu(h), p=rb.dropout_p, training=training)
    return rb.out_conv(h) + rb.skip(x)


def _attn(ab, x):
    B, C, H, W = x.shape
    heads = ab.num_heads
    d = C // heads
    flat = x.reshape(B, C, H * W)
    qkv = ab.qkv_nin(ab.norm(flat))                        # [B, 3C, T]
    if ab.split_qkv_first:
        q, k, v = qkv.reshape(B, 3, heads, d, -1).unbind(1)            # channels laid out [3][heads][d]
    else:
        q, k, v = qkv.reshape(B, heads, 3, d, -1).unbind(2)            # legacy: [heads][3][d]
    w = torch.softmax(torch.einsum('bhdt,bhds->bhts', q, k) * (d ** -0.5), dim=-1)
    o = torch.einsum('bhts,bhds->bhdt', w, v).reshape(B, C, -1)
    return (flat + ab.proj_out(o)).reshape(B, C, H, W)


def _layer(layer, x, emb, training, ckpt):
    if isinstance(layer, ResidualBlock):
        return _ckpt(_res, layer, x, emb, training, use_reentrant=False) if ckpt else _res(layer, x, emb, training)
    if isinstance(layer, AttentionBlock):
        return _ckpt(_attn, layer, x, use_reentrant=False) if ckpt else _attn(layer, x)
    if isinstance(layer, Upsample):
        x = F.interpolate(x, scale_factor=2.0, mode='nearest')
        return layer.conv(x) if layer.with_conv else x
    if isinstance(layer, Downsample):
        return layer.conv(x) if layer.with_conv else F.avg_pool2d(x, kernel_size=2, stride=2)
    return layer(x)                                        # the stem nn.Conv2d


def train_forward(model, x, timestep, y=None):
    """-> [B, out_channels, H, W] with an autograd graph over model.parameters()."""
    training, ckpt = model.training, model.use_grad_checkpoints
    emb = model.step_embed(_sinusoid(timestep, model.model_channels))
    if model.conditional:
        emb = emb + model.class_embedding(y)
    skips = []
    for seq in model.downsampling:
        for layer in seq:
            x = _layer(layer, x, emb, training, ckpt)
        skips.append(x)
    for layer in model.middle_block:
        x = _layer(layer, x, emb, training, ckpt)
    for seq in model.upsampling:
        x = torch.cat([x, skips.pop()], dim=1)
        for layer in seq:
            x = _layer(layer, x, emb, training, ckpt)
    return model.out(x)


def count_flops(model, batch):
    """Algorithmic forward FLOPs of one training batch (2 * MACs of every conv / linear + 4 h T^2 d per attention), for
    bench.py; the training step costs ~3x this (forward + two backward GEMMs per layer)."""
    total = 0
    res = model.resolution

    def conv(c, hw):
        return 2 * c.out_channels * c.in_channels * c.kernel_size[0] * c.kernel_size[-1] * hw

    def walk(layer, r):
        nonlocal total
        if isinstance(layer, ResidualBlock):
            r_out = r * 2 if layer.mode == 'up' else (r // 2 if layer.mode == 'down' else r)
            total += conv(layer.in_conv, r_out * r_out) + conv(layer.out_conv, r_out * r_out)
            if not isinstance(layer.skip, torch.nn.Identity):
                total += conv(layer.skip, r_out * r_out)
            total += 2 * layer.step_embedding.in_features * layer.step_embedding.out_features
            return r_out
        if isinstance(layer, AttentionBlock):
            C, T = layer.channels, r * r
            total += 2 * T * C * 3 * C + 2 * T * C * C + 4 * T * T * C
            return r
        if isinstance(layer, Upsample):
            if layer.with_conv:
                total += conv(layer.conv, 4 * r * r)
            return 2 * r
        if isinstance(layer, Downsample):
            if layer.with_conv:
                total += conv(layer.conv, ((r - 1) // 2 + 1) ** 2)
            return (r - 1) // 2 + 1 if layer.with_conv else r // 2
        total += conv(layer, r * r)
        return r
    for seq in list(model.downsampling) + [model.middle_block] + list(model.upsampling):
        for layer in seq:
            res = walk(layer, res)
    total += conv(model.out[2], res * res)
    E = 4 * model.model_channels
    total += 2 * model.model_channels * E + 2 * E * E
    return total * batch
