"""The reference's evaluation driver (/root/reference/cspn_pytorch/eval.py, unmodified, module-level script incl. its val()
loop :130-168) executed end to end by tools/run_reference_eval.py: stub NYU loader, stand-ins for matplotlib / skimage / h5py,
numpy-2 names, a checkpoint in the reference's own format.  Here (CPU, build container) with the reference's OWN cspn.py: this
pins that the harness really drives eval.py.  With a GPU *and* the reference tree present the same harness runs with the
drop-in first on sys.path and must reproduce those outputs (skipped on the driver's boxes: the GPU box has no reference tree;
its model-level check is tests/test_reference_model_golden_gpu.py)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
REF = '/root/reference/cspn_pytorch'

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'eval.py')), reason='reference tree not mounted')


def test_eval_py_runs_end_to_end_with_stub_loader():
    from run_reference_eval import run_eval
    outs, which = run_eval(REF, cspn='reference', samples=2)
    assert which.startswith('/root/reference') and which.endswith('cspn.py')
    assert len(outs) == 2                                    # val() visited every sample of the stub loader
    for o in outs:
        assert o.shape == (1, 1, 228, 304) and o.dtype == torch.float32 and torch.isfinite(o).all()
    assert not torch.equal(outs[0], outs[1])


@pytest.mark.gpu
def test_eval_py_with_the_dropin_reproduces_the_reference_outputs():
    from run_reference_eval import run_eval
    from oracle import cspn_numpy as onp
    ref_outs, _ = run_eval(REF, cspn='reference', samples=2)   # on a GPU box eval.py moves the model to CUDA itself
    our_outs, which = run_eval(REF, cspn='dropin', samples=2)
    assert 'dropin' in which
    for a, b in zip(our_outs, ref_outs):
        ok, ratio, normwise = onp.parity_ok(a.numpy(), b.numpy(), 1e-4)
        assert ok, (ratio, normwise)
