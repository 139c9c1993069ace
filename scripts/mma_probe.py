"""MMA-rate probe: same conv, different BLOCK_N (run with NDB_DBG=53: no TMA traffic, empty epilogue)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nice_diffusion_b200 import ops, _lib as L
dev = torch.device('cuda:0')
N, H, W, Cin, k = 32, 64, 64, 384, 3
for Cout, bn in ((256, 256), (256, 128), (256, 64), (192, 192), (192, 96), (192, 48)):
    x16 = torch.randn(N, H, W, Cin, device=dev).half()
    wp = ops.pack_conv_weight(torch.randn(Cout, Cin, k, k, device=dev) / math.sqrt(Cin * 9))
    out = torch.empty(N, H, W, Cout, device=dev)
    f = lambda: ops.conv_gemm([(x16, wp)], N, H, W, Cout, out=out, block_n=bn)
    for _ in range(3): f()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * N * H * W * Cout * Cin * 9
    n_mma = (N * H * W // 128) * (Cout // bn) * (Cin * 9 // 16)      # 128-row MMAs in total
    per_sm = n_mma / 148
    print('cout %d block_n %3d: %7.1f us  %7.1f TFLOP/s   %.0f ns per 128xNx16 MMA per SM' % (Cout, bn, us, fl / us / 1e6, us * 1e3 / per_sm))
