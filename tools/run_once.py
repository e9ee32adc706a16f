"""Runs the BASELINE cfg2 forward a few times (for ncu captures): python tools/run_once.py [algo] [reps] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cspn_b200
from cspn_b200 import _lib
from cspn_b200.synth import make_inputs

algo = {'auto': 0, 'generic': 1, 'cluster': 2}[sys.argv[1] if len(sys.argv) > 1 else 'auto']
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
H, W = (352, 1216) if len(sys.argv) <= 4 else (int(sys.argv[4]), int(sys.argv[5]))
g, d, s = [t.cuda() for t in make_inputs(0, B, 1, H, W)]
for _ in range(reps):
    out = cspn_b200.propagate2d(g, d, s, 24, '8sum', algo)
torch.cuda.synchronize()
print('algo used:', _lib.ALGO_NAMES[_lib.lib().cspn_last_algo()], 'launches:', _lib.lib().cspn_last_launches(),
      'plan:', cspn_b200.describe_plan(B, 1, H, W, 24, algo), 'checksum', float(out.double().sum()))
