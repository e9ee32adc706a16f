"""Times the planner's choice against forced alternatives (single pass, fixed cluster size):
python tools/sweep_plans.py   -> one line per (shape, steps, variant)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cspn_b200
from cspn_b200.synth import make_inputs


def time_call(g, d, s, N):
    for _ in range(3):
        cspn_b200.propagate2d(g, d, s, N, '8sum')
    evs = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cspn_b200.propagate2d(g, d, s, N, '8sum'); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[5] * 1e3


CASES = [  # B, H, W, steps, [(passes, cs) forced variants; 0 = planner's choice]
    (64, 228, 304, 4, [(0, 0), (1, 6)]),
    (64, 228, 304, 8, [(0, 0), (1, 6)]),
    (64, 228, 304, 16, [(0, 0), (1, 6), (2, 6)]),
    (64, 228, 304, 24, [(0, 0), (1, 6), (2, 6)]),
    (64, 228, 304, 48, [(0, 0), (1, 6), (2, 6), (3, 6), (4, 6)]),
    (32, 352, 1216, 24, [(0, 0), (2, 9)]),
    (32, 352, 1216, 32, [(0, 0), (1, 9)]),
    (32, 352, 1216, 48, [(0, 0), (1, 9), (3, 9)]),
    (8, 1080, 1920, 24, [(0, 0), (1, 0)]),
]
for B, H, W, N, variants in CASES:
    g, d, s = [t.cuda() for t in make_inputs(0, min(B, 8), 1, H, W)]
    if B > 8:
        rep = (B + 7) // 8
        g, d, s = [t.repeat(rep, 1, 1, 1)[:B].contiguous() for t in (g, d, s)]
    for passes, cs in variants:
        for k, v in (('CSPN_B200_FORCE_PASSES', passes), ('CSPN_B200_FORCE_CS', cs)):
            if v:
                os.environ[k] = str(v)
            else:
                os.environ.pop(k, None)
        try:
            us = time_call(g, d, s, N)
            print(f'B={B} H={H} W={W} N={N} passes={passes or "auto"} cs={cs or "auto"}: {us:.1f} us  '
                  f'{B*H*W/us:.0f} Mpx/s | {cspn_b200.describe_plan(B, 1, H, W, N)}', flush=True)
        except Exception as e:  # a forced variant may be infeasible
            print(f'B={B} H={H} W={W} N={N} passes={passes} cs={cs}: {e}', flush=True)
os.environ.pop('CSPN_B200_FORCE_PASSES', None)
os.environ.pop('CSPN_B200_FORCE_CS', None)
