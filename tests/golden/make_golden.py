"""Generate tests/golden/*.npz by running the UNMODIFIED reference module
(/root/reference/cspn_pytorch/models/cspn.py) on seeded CPU inputs.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The GPU box has no /root/reference; it checks parity against these committed files.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from cspn_b200.synth import make_inputs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


CASES = [
    # name, seed, B, C, H, W, gch, sparse, n_sample, prop_time, norm_type
    ('small_8sum',        1, 2, 1, 13, 17, 8,  'signed',    40, 5,  '8sum'),
    ('small_8sum_abs',    2, 2, 1, 13, 17, 8,  'signed',    40, 5,  '8sum_abs'),
    ('nosparse_8sum',     3, 1, 1, 9,  20, 8,  None,        0,  7,  '8sum'),
    ('c3_shared',         4, 2, 3, 11, 12, 8,  'bernoulli', 30, 4,  '8sum_abs'),
    ('gch12_extra',       5, 1, 1, 8,  8,  12, 'bernoulli', 10, 3,  '8sum'),
    ('iter1',             6, 1, 1, 6,  9,  8,  'bernoulli', 10, 1,  '8sum'),
    ('iter48_abs',        7, 1, 1, 24, 40, 8,  'bernoulli', 60, 48, '8sum_abs'),
    ('odd_w_8sum',        8, 1, 1, 10, 15, 8,  'bernoulli', 20, 6,  '8sum'),
    ('tiny_1x1',          9, 1, 1, 1,  1,  8,  None,        0,  3,  '8sum'),
    ('row_1xW',          10, 1, 1, 1,  12, 8,  'bernoulli', 4,  5,  '8sum_abs'),
    ('nyu_8sum',         11, 1, 1, 228, 304, 8, 'bernoulli', 500, 24, '8sum'),      # BASELINE cfg1
    ('nyu_8sum_abs',     12, 1, 1, 228, 304, 8, 'bernoulli', 500, 24, '8sum_abs'),
]


def main():
    assert ref_loader.available(), 'reference tree not mounted'
    index = []
    for (name, seed, B, C, H, W, gch, sparse, ns, n, norm) in CASES:
        guidance, blur, sp = make_inputs(seed, B, C, H, W, gch, sparse, ns)
        out = ref_loader.reference_forward(guidance, blur, sp, n, norm)
        big = H * W > 10000
        rec = dict(seed=seed, B=B, C=C, H=H, W=W, gch=gch, sparse=str(sparse), n_sample=ns,
                   prop_time=n, norm_type=norm, out=out.numpy())
        if not big:   # small cases carry their inputs too, so they do not depend on torch's RNG stream
            rec.update(guidance=guidance.numpy(), blur=blur.numpy())
            if sp is not None:
                rec['sparse_depth'] = sp.numpy()
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **rec)
        index.append(name)
        print(f'{name}: out shape {tuple(out.shape)} finite={bool(torch.isfinite(out).all())} '
              f'absmax={float(out.abs().max()):.4f}')
    # zero-affinity NaN semantics: all-zero guidance -> NaN everywhere
    guidance = torch.zeros(1, 8, 4, 5)
    blur = torch.ones(1, 1, 4, 5)
    out = ref_loader.reference_forward(guidance, blur, None, 2, '8sum')
    np.savez_compressed(os.path.join(HERE, 'zero_guidance_nan.npz'), guidance=guidance.numpy(),
                        blur=blur.numpy(), out=out.numpy(), prop_time=2, norm_type='8sum',
                        seed=0, B=1, C=1, H=4, W=5, gch=8, sparse='None', n_sample=0)
    print('zero_guidance_nan: all nan =', bool(torch.isnan(out).all()))


if __name__ == '__main__':
    main()
