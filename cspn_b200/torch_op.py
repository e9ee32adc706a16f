"""torch.ops.cspn_b200.* -- dispatcher registration of the propagation path (csrc/torch_op.cpp over the C ABI).

    from cspn_b200 import torch_op
    out = torch_op.propagate2d(guidance, blur_depth, sparse_depth, 24, '8sum')     # autograd-aware, torch.compile-able

What lives here (Python) and why: the FAKE kernels (shape/dtype propagation without touching data -- what
torch.compile, torch.export and FakeTensorMode need to trace the call the reference model makes at
/root/reference/cspn_pytorch/models/torch_resnet_cspn_nyu.py:375 as one opaque node) and the AUTOGRAD formulas
(train.py:196-199 back-propagates through the op; the backward is itself a registered op, so it is traceable too).
The CUDA kernels are registered in C++.  There is no CPU kernel: CPU tensors fail in the dispatcher.
"""
import os

import torch

from . import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
TORCH_LIB_PATH = os.path.join(HERE, '_build', 'libcspn_b200_torch.so')

_loaded = False


def available():
    return os.path.isfile(TORCH_LIB_PATH) and os.path.isfile(_lib.LIB_PATH)


def load():
    """Loads libcspn_b200_torch.so (once) and registers the fake kernels and autograd formulas."""
    global _loaded
    if _loaded:
        return
    if not available():
        raise _lib.CspnError(f'{TORCH_LIB_PATH} is missing: build it with `python -m cspn_b200.build --torch-op`')
    _lib.lib()                                   # libcspn_b200.so first: the shim resolves its symbols from it
    torch.ops.load_library(TORCH_LIB_PATH)
    _register()
    _loaded = True


def _register():
    lib = torch.library

    @lib.register_fake('cspn_b200::propagate2d')
    def _(guidance, blur_depth, sparse_depth, prop_time, norm_type, algo):
        torch._check(guidance.dim() == 4 and blur_depth.dim() == 4, lambda: 'guidance and blur_depth must be (B,C,H,W)')
        torch._check(guidance.shape[1] >= 8, lambda: 'guidance needs >= 8 channels')
        return torch.empty_like(blur_depth, memory_format=torch.contiguous_format)

    @lib.register_fake('cspn_b200::propagate2d_backward')
    def _(guidance, blur_depth, sparse_depth, grad_out, prop_time, norm_type, need_guidance, need_blur):
        gg = torch.empty_like(guidance, memory_format=torch.contiguous_format) if need_guidance else guidance.new_empty(0)
        gd = torch.empty_like(blur_depth, memory_format=torch.contiguous_format) if need_blur else blur_depth.new_empty(0)
        return gg, gd

    @lib.register_fake('cspn_b200::propagate3d')
    def _(guidance, feat, prop_time, norm_type):
        torch._check(guidance.dim() == 5 and feat.dim() == 5, lambda: 'guidance and feat must be (B,C,D,H,W)')
        return torch.empty_like(feat, memory_format=torch.contiguous_format)

    @lib.register_fake('cspn_b200::propagate3d_backward')
    def _(guidance, feat, grad_out, prop_time, norm_type, need_guidance, need_feat):
        gg = torch.empty_like(guidance, memory_format=torch.contiguous_format) if need_guidance else guidance.new_empty(0)
        gf = torch.empty_like(feat, memory_format=torch.contiguous_format) if need_feat else feat.new_empty(0)
        return gg, gf

    def setup2d(ctx, inputs, output):
        guidance, blur_depth, sparse_depth, prop_time, norm_type, algo = inputs
        ctx.save_for_backward(guidance, blur_depth, sparse_depth)
        ctx.cfg = (prop_time, norm_type)

    def backward2d(ctx, grad_out):
        guidance, blur_depth, sparse_depth = ctx.saved_tensors
        prop_time, norm_type = ctx.cfg
        need_g, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gg, gd = torch.ops.cspn_b200.propagate2d_backward(guidance, blur_depth, sparse_depth, grad_out.contiguous(),
                                                          prop_time, norm_type, need_g, need_d)
        return (gg if need_g else None), (gd if need_d else None), None, None, None, None

    lib.register_autograd('cspn_b200::propagate2d', backward2d, setup_context=setup2d)

    def setup3d(ctx, inputs, output):
        guidance, feat, prop_time, norm_type = inputs
        ctx.save_for_backward(guidance, feat)
        ctx.cfg = (prop_time, norm_type)

    def backward3d(ctx, grad_out):
        guidance, feat = ctx.saved_tensors
        prop_time, norm_type = ctx.cfg
        need_g, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gg, gf = torch.ops.cspn_b200.propagate3d_backward(guidance, feat, grad_out.contiguous(), prop_time, norm_type,
                                                          need_g, need_f)
        return (gg if need_g else None), (gf if need_f else None), None, None

    lib.register_autograd('cspn_b200::propagate3d', backward3d, setup_context=setup3d)


def propagate2d(guidance, blur_depth, sparse_depth=None, prop_time=24, norm_type='8sum', algo=_lib.ALGO_AUTO):
    """The dispatcher op (CUDA tensors only).  prop_time == 0 returns blur_depth itself, as cspn.py:61,83 does."""
    if prop_time == 0:
        return blur_depth
    load()
    return torch.ops.cspn_b200.propagate2d(guidance, blur_depth, sparse_depth, int(prop_time), _lib.NORM2D[norm_type], int(algo))


def propagate3d(guidance, feat, prop_time=12, norm_type='26sum_abs'):
    if prop_time == 0:
        return feat
    load()
    return torch.ops.cspn_b200.propagate3d(guidance, feat, int(prop_time), _lib.NORM3D[norm_type])
