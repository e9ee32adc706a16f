"""Seeded synthetic inputs of SURVEY.md section 8(d), shared by tests, bench.py and the golden generator.

guidance ~ randn (signed, like the un-activated conv that produces it in the reference model,
/root/reference/cspn_pytorch/models/torch_resnet_cspn_nyu.py:190,205); blur depth ~ U(0,10) m;
sparse depth = gt * Bernoulli(n_sample/HW) (the loader's sampling,
/root/reference/cspn_pytorch/nyu_dataset_loader.py:141-143, n_sample=500 in train_cspn_nyu.sh:6).
Always generated on the CPU generator so the CPU oracle and the GPU see identical bits.
"""
import torch


def make_inputs(seed, B, C, H, W, gch=8, sparse='bernoulli', n_sample=500):
    g = torch.Generator('cpu').manual_seed(seed)
    guidance = torch.randn(B, gch, H, W, generator=g)
    blur = torch.rand(B, C, H, W, generator=g) * 10
    if sparse is None or sparse == 'None':
        return guidance, blur, None
    gt = torch.rand(B, 1, H, W, generator=g) * 10
    p = min(1.0, n_sample / float(H * W))
    keep = torch.bernoulli(torch.full((B, 1, H, W), p), generator=g)
    sp = gt * keep
    if sparse == 'signed':        # exercise sign() == -1 (cspn.py:64)
        flip = torch.bernoulli(torch.full((B, 1, H, W), 0.3), generator=g)
        sp = sp * (1 - 2 * flip)
    return guidance, blur, sp


def make_inputs_3d(seed, B, C, D, H, W, signed=False):
    """3D: guide = rand(B,26,D,H,W), feat = rand(B,C,D,H,W) (cspn_paddle/demo.py:82-83)."""
    g = torch.Generator('cpu').manual_seed(seed)
    guide = torch.randn(B, 26, D, H, W, generator=g) if signed else torch.rand(B, 26, D, H, W, generator=g)
    feat = torch.rand(B, C, D, H, W, generator=g)
    return guide, feat
