import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import cspn_b200
from cspn_b200 import _lib
from cspn_b200.synth import make_inputs
from oracle import cspn_numpy as onp
B, C, H, W, n = [int(a) for a in sys.argv[1:6]]
g, d, s = make_inputs(0, B, C, H, W)
print(cspn_b200.describe_plan(B, C, H, W, n, 2), flush=True)
out = cspn_b200.propagate2d(g.cuda(), d.cuda(), s.cuda(), n, '8sum', 2)
torch.cuda.synchronize()
ref = onp.cspn2d(g.numpy(), d.numpy(), s.numpy(), n, '8sum')
print('parity', onp.parity_ok(out.cpu().numpy(), ref), flush=True)
