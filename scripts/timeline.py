#!/usr/bin/env python
"""In-graph per-kernel timeline of one UNet forward (UNetEngine.timeline): warm caches, no host launch latency.

    python scripts/timeline.py [--config in64|in128|emnist] [--batch 32] [--out profiles/timeline.txt]
"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='in64')
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--out', default=None)
    ap.add_argument('--raw', action='store_true', help='print every launch in order instead of the grouped table')
    a = ap.parse_args()
    from nice_diffusion_b200 import default_args as da
    from nice_diffusion_b200.model import DiffusionModel
    from nice_diffusion_b200.utils import randomize_zero_init_
    margs = {'in64': da.OPENAI_64_MODEL_ARGS, 'in128': da.OPENAI_128_MODEL_ARGS, 'emnist': da.EMNIST_MODEL_ARGS}[a.config]
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    model = DiffusionModel(**margs)
    randomize_zero_init_(model, 0)
    model.to(dev).eval()
    B = a.batch
    x = torch.randn(B, margs['in_channels'], margs['resolution'], margs['resolution'], device=dev)
    t = torch.full((B,), 500, device=dev)
    y = torch.randint(1, margs['num_classes'], (B,), device=dev)
    model(x, t, y)                                   # builds the plan, packs weights, warms up
    eng = model.engine_for(x)
    rows, stamp = eng.timeline()
    total = sum(r[2] for r in rows)
    lines = ['in-graph timeline ({} B={}): {} launches, sum {:.1f} us, stamp cost {:.2f} us'.format(a.config, B, len(rows), total, stamp)]
    if a.raw:
        for i, (name, desc, us, _f, _b) in enumerate(rows):
            lines.append('{:4d} {:12s} {:46s} {:8.1f}'.format(i, name, desc, us))
    else:
        agg = collections.OrderedDict()
        by_kernel = collections.Counter()
        for name, desc, us, fl, nb in rows:
            d = agg.setdefault((name, desc), [0, 0.0, 0, 0])
            d[0] += 1
            d[1] += us
            d[2] += fl
            d[3] += nb
            by_kernel[name] += us
        lines.append(str({k: round(v / 1e3, 3) for k, v in by_kernel.items()}))
        lines.append('{:12s} {:46s} {:>3s} {:>9s} {:>9s} {:>8s} {:>7s}'.format('kernel', 'shape', 'n', 'ms/step', 'us/launch', 'TFLOP/s', 'GB/s'))
        for (name, desc), (cnt, us, fl, nb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append('{:12s} {:46s} {:3d} {:9.3f} {:9.1f} {:8.1f} {:7.0f}'.format(name, desc, cnt, us / 1e3, us / cnt, fl / max(us, 1e-9) / 1e6, nb / max(us, 1e-9) / 1e3))
    text = '\n'.join(lines)
    print(text)
    if a.out:
        with open(a.out, 'w') as f:
            f.write(text + '\n')


if __name__ == '__main__':
    main()
