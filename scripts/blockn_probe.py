"""Sweep BLOCK_N for a few conv shapes (CUDA graph of 20 launches, CUDA events)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nice_diffusion_b200 import ops, _lib as L
dev = torch.device('cuda:0')
# N, H, W, Cin, Cout, k, residual, stats, fp16 out
shapes = [(32, 8, 8, 768, 768, 3, 1, 1, 0), (32, 8, 8, 1536, 768, 3, 0, 1, 0), (32, 16, 16, 576, 576, 3, 1, 1, 0),
          (32, 16, 16, 1152, 576, 3, 0, 1, 0), (32, 32, 32, 384, 384, 3, 1, 1, 0), (32, 8, 8, 768, 2304, 1, 0, 0, 1),
          (32, 8, 8, 768, 768, 1, 1, 1, 0), (32, 16, 16, 576, 1728, 1, 0, 0, 1), (32, 32, 32, 384, 1152, 1, 0, 0, 1),
          (32, 64, 64, 192, 192, 3, 1, 1, 0)]
for (N, H, W, Cin, Cout, k, res, stats, o16) in shapes:
    x16 = torch.randn(N, H, W, Cin, device=dev).half()
    w = torch.randn(Cout, Cin, k, k, device=dev) / math.sqrt(Cin * k * k)
    wp = ops.pack_conv_weight(w)
    b = torch.randn(Cout, device=dev)
    r = torch.randn(N, H, W, Cout, device=dev) if res else None
    out = torch.empty(N, H, W, Cout, device=dev, dtype=torch.float16 if o16 else torch.float32)
    line = 'N%d %dx%d %d->%d k%d res%d st%d o16=%d:' % (N, H, W, Cin, Cout, k, res, stats, o16)
    for bn in (256, 192, 128, 96, 64):
        if Cout % bn:
            continue
        st = ops.conv_stats_alloc(N, H, W, Cout, dev, Cin * k * k, block_n=bn) if stats else None
        f = lambda: ops.conv_gemm([(x16, wp)], N, H, W, Cout, bias=b, residual=r, residual_mode=L.RES_SAME, out=out, stats=st,
                                  out_kind=L.OUT_F16_NHWC if o16 else L.OUT_F32_NHWC, block_n=bn)
        for _ in range(3): f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): f()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        fl = 2.0 * N * H * W * Cout * Cin * k * k
        line += '  bn%d %.1fus (%.0f TF)' % (bn, us, fl / us / 1e6)
    print(line, flush=True)
