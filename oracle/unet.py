"""ORACLE (test infrastructure, not product code) -- CPU restatement of the reference UNet
forward as a pure function of (config, state_dict, x, timestep, y).

The reference's arithmetic lives in PyTorch (un-pinned dependency, setup.py:13-21); this file
restates the reference's *algorithm* -- which torch op is applied to what, in which order -- on
top of the same torch CPU ops, so that on identical weights it reproduces the reference
bit-for-bit on CPU (checked by oracle/make_golden.py against /root/reference itself; the
reference ships no tests / golden vectors -- SURVEY.md section 4).

Follows (paths relative to /root/reference):
  * nicediffusion/model.py:322-449  DiffusionModel.__init__ (block topology)
  * nicediffusion/model.py:451-476  DiffusionModel.forward
  * nicediffusion/model.py:188-211  ResidualBlock._forward
  * nicediffusion/model.py:260-291  AttentionBlock._forward (both qkv channel orders)
  * nicediffusion/model.py:51-112   Upsample / Downsample
  * nicediffusion/model.py:514-523  timestep_embedding

``operand_cast`` is a study hook (oracle/precision_study.py): a callable applied to every
tensor-core operand (conv / linear / matmul inputs and weights) to emulate reduced-precision
operands with fp32 accumulation. It is ``None`` for the oracle proper.
"""
import math

import torch
import torch.nn.functional as F


def topology(cfg):
    """Enumerate the blocks exactly as model.py:363-449 constructs them.

    Returns dict(down=[[layer,...],...], mid=[...], up=[[...],...], head_channels=int) where each
    layer is a tuple:
      ('conv', cin, cout)                      stem conv3x3
      ('res', cin, cout, mode)                 mode in {'same','up','down'}
      ('attn', channels, heads)
      ('upsample', cin, cout, with_conv) / ('downsample', cin, cout, with_conv)
    """
    mc = cfg['model_channels']
    mult = tuple(cfg.get('channel_mult', (1, 2, 4, 8)))
    nrb = cfg['num_res_blocks']
    attn_res = tuple(cfg['attention_resolutions'])
    updown = cfg.get('resblock_updown', False)
    conv_resample = cfg.get('conv_resample', True)
    num_heads = cfg.get('num_heads', 1)
    nhc = cfg.get('num_head_channels', None)

    def heads_of(c):
        if nhc is None:
            return num_heads
        assert c % nhc == 0
        return c // nhc

    cur = first = int(mc * mult[0])
    res = cfg['resolution']
    down = [[('conv', cfg['in_channels'], cur)]]
    chans = [cur]
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [('res', cur, int(mc * m), 'same')]
            cur = int(mc * m)
            if res in attn_res:
                layers.append(('attn', cur, heads_of(cur)))
            chans.append(cur)
            down.append(layers)
        cur = int(mc * m)
        if level != len(mult) - 1:
            if updown:
                down.append([('res', cur, cur, 'down')])
            else:
                down.append([('downsample', cur, cur, conv_resample)])
            chans.append(cur)
            res //= 2
    mid = [('res', cur, cur, 'same'), ('attn', cur, heads_of(cur)), ('res', cur, cur, 'same')]
    up = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            skip = chans.pop()
            layers = [('res', cur + skip, int(mc * m), 'same')]
            cur = int(mc * m)
            if res in attn_res:
                layers.append(('attn', cur, heads_of(cur)))
            if level != 0 and i == nrb:
                if updown:
                    layers.append(('res', cur, cur, 'up'))
                else:
                    layers.append(('upsample', cur, cur, conv_resample))
                res *= 2
            up.append(layers)
    return dict(down=down, mid=mid, up=up, head_in=first, head_norm=cur)


def timestep_embedding(timesteps, dim, max_period=10000):
    """model.py:514-523 -- frequency vector is built on the CPU in fp32."""
    half = dim // 2
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_period) / half))
    args = timesteps[:, None].float() * freqs[None].to(timesteps.device)
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class _Ops:
    def __init__(self, operand_cast):
        self.c = operand_cast if operand_cast is not None else (lambda t: t)

    def conv2d(self, x, w, b, **kw):
        return F.conv2d(self.c(x), self.c(w), b, **kw)

    def conv1d(self, x, w, b):
        return F.conv1d(self.c(x), self.c(w), b)

    def linear(self, x, w, b):
        return F.linear(self.c(x), self.c(w), b)

    def matmul(self, a, b):
        return self.c(a) @ self.c(b)


def _resblock(ops, sd, p, x, emb, mode, adaptive):
    """model.py:188-211."""
    h = F.silu(F.group_norm(x, 32, sd[p + 'in_norm.weight'], sd[p + 'in_norm.bias'], eps=1e-5))
    if mode == 'up':
        h = F.interpolate(h, scale_factor=2.0, mode='nearest')
        x = F.interpolate(x, scale_factor=2.0, mode='nearest')
    elif mode == 'down':
        h = F.avg_pool2d(h, kernel_size=(2, 2), stride=(2, 2))
        x = F.avg_pool2d(x, kernel_size=(2, 2), stride=(2, 2))
    h = ops.conv2d(h, sd[p + 'in_conv.weight'], sd[p + 'in_conv.bias'], padding=1)
    e = ops.linear(F.silu(emb), sd[p + 'step_embedding.weight'], sd[p + 'step_embedding.bias'])[:, :, None, None]
    if adaptive:
        scale, shift = torch.chunk(e, 2, dim=1)
        h = F.group_norm(h, 32, sd[p + 'out_norm.weight'], sd[p + 'out_norm.bias'], eps=1e-5) * (1 + scale) + shift
    else:
        h = h + e
        h = F.group_norm(h, 32, sd[p + 'out_norm.weight'], sd[p + 'out_norm.bias'], eps=1e-5)
    h = F.silu(h)  # dropout is identity in eval mode
    h = ops.conv2d(h, sd[p + 'out_conv.weight'], sd[p + 'out_conv.bias'], padding=1)
    if (p + 'skip.weight') in sd:
        w = sd[p + 'skip.weight']
        x = ops.conv2d(x, w, sd[p + 'skip.bias'], padding=(w.shape[-1] // 2))
    return h + x


def _attention(ops, sd, p, x, heads, split_qkv_first):
    """model.py:260-291."""
    B, C, H, W = x.shape
    T = H * W
    d = C // heads
    scale = d ** -0.5
    x = x.reshape(B, C, T)
    qkv = F.group_norm(x, 32, sd[p + 'norm.weight'], sd[p + 'norm.bias'], eps=1e-5)
    qkv = ops.conv1d(qkv, sd[p + 'qkv_nin.weight'], sd[p + 'qkv_nin.bias'])  # b, 3c, t
    if split_qkv_first:
        qkv = qkv.permute(0, 2, 1).reshape(B, T, 3, heads, d).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)  # b, heads, t, d
        w = ops.matmul(q, k.transpose(2, 3)) * scale
        w = torch.softmax(w, dim=-1)
        h = ops.matmul(w, v).transpose(1, 2).reshape(B, T, C).permute(0, 2, 1)
    else:
        q, k, v = qkv.reshape(B * heads, 3 * d, T).split(d, dim=1)
        w = ops.matmul(q.transpose(1, 2), k) * scale
        w = torch.softmax(w, dim=2)
        h = ops.matmul(v, w.transpose(1, 2)).reshape(B, -1, T)  # einsum('bts,bcs->bct')
    h = ops.conv1d(h, sd[p + 'proj_out.weight'], sd[p + 'proj_out.bias'])
    return (h + x).reshape(B, C, H, W)


def _run_layers(ops, cfg, sd, prefix, layers, x, emb):
    adaptive = cfg.get('use_adaptive_gn', False)
    sqf = cfg.get('split_qkv_first', True)
    for j, layer in enumerate(layers):
        p = '{}{}.'.format(prefix, j)
        kind = layer[0]
        if kind == 'conv':
            x = ops.conv2d(x, sd[p + 'weight'], sd[p + 'bias'], padding=1)
        elif kind == 'res':
            x = _resblock(ops, sd, p, x, emb, layer[3], adaptive)
        elif kind == 'attn':
            x = _attention(ops, sd, p, x, layer[2], sqf)
        elif kind == 'upsample':
            x = F.interpolate(x, scale_factor=2.0, mode='nearest')
            if layer[3]:
                x = ops.conv2d(x, sd[p + 'conv.weight'], sd[p + 'conv.bias'], padding=1)
        elif kind == 'downsample':
            if layer[3]:
                x = ops.conv2d(x, sd[p + 'conv.weight'], sd[p + 'conv.bias'], stride=2, padding=1)
            else:
                x = F.avg_pool2d(x, kernel_size=(2, 2), stride=(2, 2))
        else:
            raise AssertionError(kind)
    return x


@torch.no_grad()
def unet_forward(cfg, sd, x, timestep, y=None, operand_cast=None):
    """model.py:451-476. ``sd`` maps the reference's state-dict names to fp32 CPU tensors."""
    ops = _Ops(operand_cast)
    conditional = cfg.get('num_classes', None) is not None
    assert (y is not None) == conditional, 'pass y iff class-conditional model'
    assert x.shape[2] == cfg['resolution'] and x.shape[3] == cfg['resolution'], \
        'incorrect resolution: {}'.format(x.shape[2:])
    topo = topology(cfg)
    emb = timestep_embedding(timestep, cfg['model_channels'])
    emb = ops.linear(emb, sd['step_embed.0.weight'], sd['step_embed.0.bias'])
    emb = ops.linear(F.silu(emb), sd['step_embed.2.weight'], sd['step_embed.2.bias'])
    if conditional:
        emb = emb + F.embedding(y, sd['class_embedding.weight'])
    xs = []
    for i, layers in enumerate(topo['down']):
        x = _run_layers(ops, cfg, sd, 'downsampling.{}.'.format(i), layers, x, emb)
        xs.append(x)
    x = _run_layers(ops, cfg, sd, 'middle_block.', topo['mid'], x, emb)
    for i, layers in enumerate(topo['up']):
        x = torch.cat([x, xs.pop()], dim=1)
        x = _run_layers(ops, cfg, sd, 'upsampling.{}.'.format(i), layers, x, emb)
    x = F.silu(F.group_norm(x, 32, sd['out.0.weight'], sd['out.0.bias'], eps=1e-5))
    return ops.conv2d(x, sd['out.2.weight'], sd['out.2.bias'], padding=1)


def param_shapes(cfg):
    """Ordered {name: shape} of the reference state dict (model.py ctor order), used to create
    seed-fixed weights without importing the reference."""
    topo = topology(cfg)
    mc = cfg['model_channels']
    emb_dim = 4 * mc
    adaptive = cfg.get('use_adaptive_gn', False)
    shapes = {}

    def add(name, *shape):
        shapes[name] = tuple(shape)

    add('step_embed.0.weight', emb_dim, mc); add('step_embed.0.bias', emb_dim)
    add('step_embed.2.weight', emb_dim, emb_dim); add('step_embed.2.bias', emb_dim)
    if cfg.get('num_classes', None) is not None:
        add('class_embedding.weight', cfg['num_classes'], emb_dim)

    def add_layers(prefix, layers):
        for j, layer in enumerate(layers):
            p = '{}{}.'.format(prefix, j)
            kind = layer[0]
            if kind == 'conv':
                add(p + 'weight', layer[2], layer[1], 3, 3); add(p + 'bias', layer[2])
            elif kind == 'res':
                cin, cout = layer[1], layer[2]
                if cin != cout:
                    add(p + 'skip.weight', cout, cin, 1, 1); add(p + 'skip.bias', cout)
                add(p + 'in_norm.weight', cin); add(p + 'in_norm.bias', cin)
                add(p + 'in_conv.weight', cout, cin, 3, 3); add(p + 'in_conv.bias', cout)
                add(p + 'out_norm.weight', cout); add(p + 'out_norm.bias', cout)
                add(p + 'out_conv.weight', cout, cout, 3, 3); add(p + 'out_conv.bias', cout)
                n = 2 * cout if adaptive else cout
                add(p + 'step_embedding.weight', n, emb_dim); add(p + 'step_embedding.bias', n)
            elif kind == 'attn':
                c = layer[1]
                add(p + 'qkv_nin.weight', 3 * c, c, 1); add(p + 'qkv_nin.bias', 3 * c)
                add(p + 'norm.weight', c); add(p + 'norm.bias', c)
                add(p + 'proj_out.weight', c, c, 1); add(p + 'proj_out.bias', c)
            elif kind in ('upsample', 'downsample'):
                if layer[3]:
                    add(p + 'conv.weight', layer[2], layer[1], 3, 3); add(p + 'conv.bias', layer[2])

    for i, layers in enumerate(topo['down']):
        add_layers('downsampling.{}.'.format(i), layers)
    add_layers('middle_block.', topo['mid'])
    for i, layers in enumerate(topo['up']):
        add_layers('upsampling.{}.'.format(i), layers)
    add('out.0.weight', topo['head_norm']); add('out.0.bias', topo['head_norm'])
    add('out.2.weight', cfg['out_channels'], topo['head_in'], 3, 3); add('out.2.bias', cfg['out_channels'])
    return shapes
