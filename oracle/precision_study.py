"""ORACLE study (test infrastructure): which tensor-core operand format meets the parity bar?

Emulates reduced-precision GEMM operands with fp32 accumulation by rounding every conv / linear /
matmul operand (oracle.unet ``operand_cast`` hook) and reports max|d|/max|ref| per config.
    python -m oracle.precision_study
"""
import numpy as np
import torch

from oracle import CONFIGS
from oracle.unet import unet_forward
from oracle.weights import random_state_dict
from oracle.make_golden import make_inputs


def main():
    torch.set_num_threads(8)
    casts = {
        'bf16': lambda t: t.to(torch.bfloat16).float(),
        'fp16': lambda t: t.to(torch.float16).float(),
        'tf32(trunc)': lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32),
    }
    for cname, batch in (('emnist', 4), ('in64', 1)):
        cfg = CONFIGS[cname]
        sd = random_state_dict(cfg, seed=1234)
        x, t, y = make_inputs(cfg, batch, seed=99)
        ref = unet_forward(cfg, sd, x, t, y)
        for name, c in casts.items():
            out = unet_forward(cfg, sd, x, t, y, operand_cast=c)
            err = float((out - ref).abs().max() / ref.abs().max())
            print('{:8s} {:12s} max|d|/max|ref| = {:.3e}'.format(cname, name, err))


if __name__ == '__main__':
    main()
