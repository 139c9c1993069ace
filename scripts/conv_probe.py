"""Time single conv shapes (CUDA graph of 20 launches, CUDA events) under the env switches NDB_DISABLE_2CTA / NDB_DBG."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nice_diffusion_b200 import ops, _lib as L
if os.environ.get('NDB_LIB'):          # e.g. a -DNDB_BRINGUP build kept next to the product library
    L.LIB_PATH = os.environ['NDB_LIB']
dev = torch.device('cuda:0')
# N, H, W, Cin, Cout, k, residual, stats, fp16 out
shapes = [(32, 32, 32, 384, 1152, 1, 0, 0, 1), (32, 32, 32, 384, 384, 1, 1, 1, 0), (32, 8, 8, 768, 768, 3, 1, 1, 0),
          (32, 64, 64, 192, 192, 3, 1, 1, 0), (32, 64, 64, 384, 192, 3, 0, 1, 0), (32, 64, 64, 192, 192, 3, 0, 1, 0),
          (32, 64, 64, 192, 192, 3, 0, 1, 1), (32, 16, 16, 576, 1728, 1, 0, 0, 1), (32, 16, 16, 576, 576, 1, 1, 1, 0),
          # + a fused 1x1 skip segment over C2 channels (10th field): the up-path out_conv + skip(concat) layers
          (32, 64, 64, 192, 192, 3, 0, 1, 0, 384), (32, 32, 32, 384, 384, 3, 0, 1, 0, 768), (32, 16, 16, 576, 576, 3, 0, 1, 0, 1152),
          (32, 16, 16, 1152, 576, 3, 0, 1, 1), (32, 16, 16, 1152, 192, 3, 0, 1, 1)]
if len(sys.argv) > 1:
    shapes = [shapes[int(a)] for a in sys.argv[1:]]
for shp in shapes:
    (N, H, W, Cin, Cout, k, res, stats, o16), C2 = shp[:9], (shp[9] if len(shp) > 9 else 0)
    x16 = torch.randn(N, H, W, Cin, device=dev).half()
    w = torch.randn(Cout, Cin, k, k, device=dev) / math.sqrt(Cin * k * k)
    wp = ops.pack_conv_weight(w)
    b = torch.randn(Cout, device=dev)
    r = torch.randn(N, H, W, Cout, device=dev) if res else None
    out = torch.empty(N, H, W, Cout, device=dev, dtype=torch.float16 if o16 else torch.float32)
    segs = [(x16, wp)]
    if C2:
        segs.append((torch.randn(N, H, W, C2, device=dev).half(), ops.pack_conv_weight(torch.randn(Cout, C2, 1, 1, device=dev) / math.sqrt(C2))))
    st = ops.conv_stats_alloc(N, H, W, Cout, dev, Cin * k * k + C2) if stats else None
    f = lambda: ops.conv_gemm(segs, N, H, W, Cout, bias=b, residual=r, residual_mode=L.RES_SAME, out=out, stats=st,
                              out_kind=L.OUT_F16_NHWC if o16 else L.OUT_F32_NHWC)
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * N * H * W * Cout * (Cin * k * k + C2)
    print('N%d %dx%d %d->%d k%d res%d st%d o16=%d skip%d: %8.1f us  %7.1f TFLOP/s' % (N, H, W, Cin, Cout, k, res, stats, o16, C2, us, fl / us / 1e6))
