"""Times the 3D forward: python tools/time_3d.py [B D H W N]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cspn_b200
from cspn_b200.synth import make_inputs_3d
B, D, H, W, N = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (8, 64, 96, 312, 12)
g, f = [t.cuda() for t in make_inputs_3d(0, B, 1, D, H, W)]
for mode in ('26sum_abs', 'paddle'):
    m = cspn_b200.Affinity_Propagate3D(N, 3, mode)
    for _ in range(2): m(g, f)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): m(g, f)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    vox = B * D * H * W
    print(f'3D {B}x{D}x{H}x{W} N={N} {mode}: {dt*1e3:.2f} ms  {vox/dt/1e6:.0f} Mvox/s  {112*vox/dt/1e9:.0f} GB/s algorithmic '
          f'({(108*N+108+112)*vox/dt/1e9:.0f} GB/s of streamed weights+volumes)')
