"""ORACLE (test infrastructure) -- seed-fixed synthetic weights for parity tests.

The reference zero-initialises every ResBlock out_conv, attention proj_out and the head conv
(model.py:177,253,448,507-510), so a freshly constructed model outputs exact zeros and parity on it
is vacuous. Parity tests therefore use fully re-randomised tensors:

  conv / linear weights   N(0, 1/fan_in)
  biases                  N(0, 0.02^2)
  GroupNorm weight        1 + 0.1 N(0,1);  GroupNorm bias 0.1 N(0,1)
  class embedding         N(0, 1)

drawn from numpy's legacy ``RandomState(seed)`` stream (bit-stable across numpy versions), one
tensor after the other in state-dict order.
"""
import numpy as np
import torch

from .unet import param_shapes


def random_state_dict(cfg, seed=0):
    rs = np.random.RandomState(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        n = int(np.prod(shape))
        z = rs.standard_normal(n).astype(np.float32).reshape(shape)
        if name.endswith('norm.weight') or name == 'out.0.weight':
            v = 1.0 + 0.1 * z
        elif name.endswith('norm.bias') or name == 'out.0.bias':
            v = 0.1 * z
        elif name == 'class_embedding.weight':
            v = z
        elif name.endswith('.bias'):
            v = 0.02 * z
        else:
            fan_in = int(np.prod(shape[1:]))
            v = z / np.sqrt(fan_in)
        sd[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd
