"""SM clock / power while one conv shape runs back to back for ~3 s (is the tensor loop power-throttled?)."""
import math, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nice_diffusion_b200 import ops, _lib as L
dev = torch.device('cuda:0')
N, H, W, Cin, Cout, k = 32, 64, 64, 384, 192, 3
x16 = torch.randn(N, H, W, Cin, device=dev).half()
wp = ops.pack_conv_weight(torch.randn(Cout, Cin, k, k, device=dev) / math.sqrt(Cin * 9))
out = torch.empty(N, H, W, Cout, device=dev)
f = lambda: ops.conv_gemm([(x16, wp)], N, H, W, Cout, out=out)
for _ in range(3): f()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(50): f()
torch.cuda.synchronize()
p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap', '--format=csv,noheader', '-lms', '100'],
                     stdout=subprocess.PIPE, text=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); times = []
while time.time() - t0 < 3.0:
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / 50 * 1e3)
p.terminate()
rows = p.stdout.read().strip().split('\n')
print('us per launch: first %.1f  median %.1f  last %.1f' % (times[0], sorted(times)[len(times) // 2], times[-1]))
print('clock samples:', rows[:3], '...', rows[-4:])
