"""Diagnostics for bring-up on the GPU box: tiny GEMM/attention cases with printed error structure."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from nice_diffusion_b200 import ops, _lib as L

dev = torch.device('cuda:0')
print('device', torch.cuda.get_device_name(0), 'supported', L.load().ndb_device_supported())


def report(name, got, ref):
    err = (got.float() - ref.float()).abs()
    rel = float(err.max() / ref.abs().max())
    print('{:40s} rel {:.3e} max|ref| {:.3f} nan {}'.format(name, rel, float(ref.abs().max()), bool(torch.isnan(got).any())))
    if rel > 1e-2 or rel != rel:
        flat = err.flatten()
        idx = int(flat.argmax())
        print('   worst idx', idx, 'got', float(got.flatten()[idx]), 'ref', float(ref.flatten()[idx]))
        print('   got[0,:4,:8]\n', got.reshape(-1, got.shape[-1])[:4, :8])
        print('   ref[0,:4,:8]\n', ref.reshape(-1, ref.shape[-1])[:4, :8])
        bad_rows = (err.reshape(-1, err.shape[-1]).max(1).values > 1e-2 * float(ref.abs().max())).nonzero().flatten()
        bad_cols = (err.reshape(-1, err.shape[-1]).max(0).values > 1e-2 * float(ref.abs().max())).nonzero().flatten()
        print('   bad rows', bad_rows[:40].tolist(), '... total', len(bad_rows))
        print('   bad cols', bad_cols[:40].tolist(), '... total', len(bad_cols))
    return rel


g = torch.Generator().manual_seed(0)
# 1x1 GEMM M=128 (1 tile), K=64, N=64
for (N, H, W, Cin, Cout, k) in [(1, 2, 64, 64, 64, 1), (1, 2, 64, 128, 64, 1), (1, 2, 64, 64, 64, 3), (2, 16, 16, 128, 192, 3),
                                (2, 8, 8, 64, 64, 3), (2, 28, 28, 64, 64, 3)]:
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).to(dev)
    x16 = x.permute(0, 2, 3, 1).contiguous().half()
    wp = ops.pack_conv_weight(w)
    out = ops.conv_gemm([(x16, wp)], N, H, W, Cout)
    torch.cuda.synchronize()
    ref = F.conv2d(x.half().float(), w.half().float(), None, padding=k // 2).permute(0, 2, 3, 1)
    report('conv N{} {}x{} {}->{} k{}'.format(N, H, W, Cin, Cout, k), out, ref)

for (B, T, heads, d, order) in [(1, 128, 1, 64, 0), (1, 64, 1, 64, 0), (1, 256, 2, 64, 0), (1, 128, 1, 32, 0), (1, 128, 1, 128, 0)]:
    C = heads * d
    qkv = torch.randn(B, T, 3 * C, generator=g).to(dev).half()
    out = ops.attention(qkv, B, T, heads, d, order=order)
    torch.cuda.synchronize()
    f = qkv.float()
    q, k, v = f.reshape(B, T, 3, heads, d).permute(2, 0, 3, 1, 4).unbind(0)
    s = (q @ k.transpose(2, 3)) * d ** -0.5
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, C)
    report('attn B{} T{} h{} d{}'.format(B, T, heads, d), out, ref)
print('probe done')
