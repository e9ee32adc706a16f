"""Times forward + backward of the 2D module (training seam): python tools/time_bwd.py [B H W N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cspn_b200
from cspn_b200.synth import make_inputs
B, H, W, N = [int(a) for a in sys.argv[1:5]] if len(sys.argv) > 4 else (8, 228, 304, 24)
g, d, s = [t.cuda() for t in make_inputs(0, B, 1, H, W)]
g.requires_grad_(True); d.requires_grad_(True)
m = cspn_b200.Affinity_Propagate(N, 3, '8sum')
go = torch.rand_like(d)
def step():
    g.grad = None; d.grad = None
    out = m(g, d, s)
    out.backward(go)
for _ in range(3): step()
evs = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record(); evs.append((a, b))
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in evs)[5]
with torch.no_grad():
    evs = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); m(g, d, s); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
fwd = sorted(a.elapsed_time(b) for a, b in evs)[5]
print(f'2D train step B={B} H={H} W={W} N={N}: forward+backward {ms:.3f} ms (forward alone {fwd:.3f} ms), {B*H*W/ms/1e3:.0f} Mpx/s', flush=True)
