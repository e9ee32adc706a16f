"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous batch sharding + final gather.
The per-rank compute is stood in for by the oracle (no GPU here); on a GPU box bench.py runs the same sharding with
the CUDA kernels and NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cspn_b200.sharding import gather_outputs, shard_range  # noqa: E402


def test_shard_range_covers_batch():
    for B in (1, 2, 5, 32, 33, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == B
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s1 == s0 + c0
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, B, tmpdir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from cspn_b200.synth import make_inputs
        from oracle import cspn_numpy as onp
        g, d, s = make_inputs(0, B, 1, 9, 12, 8, 'bernoulli', 10)      # every rank builds the same full batch
        start, count = shard_range(B, world, rank)
        local = onp.cspn2d(g[start:start + count].numpy(), d[start:start + count].numpy(),
                           s[start:start + count].numpy(), 4, '8sum')
        full = gather_outputs(torch.from_numpy(local), B)
        ref = onp.cspn2d(g.numpy(), d.numpy(), s.numpy(), 4, '8sum')
        np.save(os.path.join(tmpdir, f'ok{rank}.npy'), np.array([np.array_equal(full.numpy(), ref)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('B', [4, 5])
def test_two_rank_shard_and_gather_gloo(tmp_path, B):
    port = 29500 + (os.getpid() % 2000) + B
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(tmp_path / f'ok{r}.npy')[0]
