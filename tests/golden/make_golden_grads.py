"""Generate tests/golden/grads/*.npz: gradients obtained by autograd through the UNMODIFIED reference module
(/root/reference/cspn_pytorch/models/cspn.py) on seeded CPU inputs, in fp32 as train.py runs it.

    python tests/golden/make_golden_grads.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from cspn_b200.synth import make_inputs  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'grads')

CASES = [
    # name, seed, B, C, H, W, gch, sparse, n_sample, prop_time, norm_type
    ('g_small_8sum',     21, 2, 1, 12, 16, 8, 'signed',    25, 5,  '8sum'),
    ('g_small_8sum_abs', 22, 2, 1, 12, 16, 8, 'signed',    25, 5,  '8sum_abs'),
    ('g_c2_shared',      23, 1, 2, 9,  20, 9, 'bernoulli', 25, 3,  '8sum'),
    ('g_nosparse',       24, 2, 1, 10, 15, 8, None,        0,  4,  '8sum_abs'),
    ('g_iter24',         25, 1, 1, 24, 32, 8, 'bernoulli', 25, 24, '8sum'),
    ('g_cluster_shape',  26, 1, 1, 48, 132, 8, 'bernoulli', 60, 6, '8sum_abs'),   # W % 4 == 0: the cluster forward's territory
]


def main():
    assert ref_loader.available(), 'reference tree not mounted'
    os.makedirs(OUT, exist_ok=True)
    for (name, seed, B, C, H, W, gch, sparse, ns, n, norm) in CASES:
        guidance, blur, sp = make_inputs(seed, B, C, H, W, gch, sparse, ns)
        go = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(seed + 1000))
        out, gg, gd = ref_loader.reference_gradients(guidance, blur, sp, go, n, norm)
        rec = dict(guidance=guidance.numpy(), blur=blur.numpy(), grad_out=go.numpy(), out=out.numpy(),
                   grad_guidance=gg.numpy(), grad_blur=gd.numpy(), prop_time=n, norm_type=norm)
        if sp is not None:
            rec['sparse_depth'] = sp.numpy()
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
        print(f'{name}: |grad_guidance| mean {float(gg.abs().mean()):.4g} max {float(gg.abs().max()):.4g}, '
              f'|grad_blur| mean {float(gd.abs().mean()):.4g}')


if __name__ == '__main__':
    main()
