"""Times the 2D forward for arbitrary shapes: python tools/time_shape.py algo B H W N [B H W N ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
try:
    import pynvml
    pynvml.nvmlInit(); _h = pynvml.nvmlDeviceGetHandleByIndex(0)
    def clk(): return pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
except Exception:
    def clk(): return -1

import cspn_b200
from cspn_b200.synth import make_inputs

algo = {'auto': 0, 'generic': 1, 'cluster': 2}[sys.argv[1]]
args = [int(a) for a in sys.argv[2:]]
for i in range(0, len(args), 4):
    B, H, W, N = args[i:i + 4]
    g, d, s = [t.cuda() for t in make_inputs(0, min(B, 8), 1, H, W)]
    if B > 8:   # replicate (timing only)
        rep = (B + 7) // 8
        g, d, s = [t.repeat(rep, 1, 1, 1)[:B].contiguous() for t in (g, d, s)]
    for _ in range(3):
        cspn_b200.propagate2d(g, d, s, N, '8sum', algo)
    evs = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cspn_b200.propagate2d(g, d, s, N, '8sum', algo); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    c_mhz = clk()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    px = B * H * W
    print(f'B={B} H={H} W={W} N={N}: median {ms[5]*1e3:.1f} us  min {ms[0]*1e3:.1f} us  {px/ms[5]/1e3:.0f} Mpx/s  '
          f'{44*px/ms[5]/1e6:.0f} GB/s algorithmic  sm_clk={c_mhz} MHz | {cspn_b200.describe_plan(B, 1, H, W, N, algo)}', flush=True)
