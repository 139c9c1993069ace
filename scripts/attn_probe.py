"""Time the attention kernel alone (CUDA graph of 10 launches, CUDA events). NDB_LIB selects another build of the library
(e.g. a -DNDB_BRINGUP one, whose NDB_ATTN_DBG switches remove / duplicate the first-pass TMEM reads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nice_diffusion_b200 import ops, _lib as L
if os.environ.get('NDB_LIB'):
    L.LIB_PATH = os.environ['NDB_LIB']
dev = torch.device('cuda:0')
for (B, T, heads, d) in [(32, 1024, 6, 64), (32, 256, 9, 64), (32, 64, 12, 64), (16, 4096, 4, 64)]:
    qkv = torch.randn(B, T, 3 * heads * d, device=dev).half()
    out = torch.empty(B, T, heads * d, device=dev, dtype=torch.float16)
    f = lambda: ops.attention(qkv, B, T, heads, d, order=0, out=out)
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    fl = 4.0 * B * heads * T * T * d
    print('B%d T%d h%d d%d: %8.1f us  %7.1f TFLOP/s' % (B, T, heads, d, us, fl / us / 1e6))
