"""Multi-GPU check of the fused / chunked final gather (run under torchrun on 2..8 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gather_check.py

Every rank propagates its own images; all three gather routes must leave the same full batch on every rank, equal to what
each rank computes alone.  Prints one line per route with its time per step (CUDA events, max over ranks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import cspn_b200
from cspn_b200.gather import ChunkedGather, FusedGather
from cspn_b200.sharding import gather_outputs
from cspn_b200.synth import make_inputs

B, H, W, N = [int(a) for a in sys.argv[1:5]] if len(sys.argv) > 4 else (32, 352, 1216, 24)
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
g, d, s = [t.to(dev) for t in make_inputs(rank, min(B, 8), 1, H, W)]
if B > 8:
    g, d, s = [t.repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous() for t in (g, d, s)]
mine = cspn_b200.propagate2d(g, d, s, N, '8sum')
ref = gather_outputs(mine, world * B)                       # kernel, then NCCL all_gather_into_tensor
assert torch.equal(ref[rank * B:(rank + 1) * B], mine)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def report(name, ms, ok, extra=''):
    oks = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'{name}: {ms:.4f} ms/step on {world} GPUs, identical on every rank: {bool(oks.item())} {extra}', flush=True)
    return bool(oks.item())


full = torch.empty(world * B, 1, H, W, device=dev)
all_ok = report('kernel only', timed(lambda: cspn_b200.propagate2d(g, d, s, N, '8sum')), True)
all_ok &= report('serial NCCL (kernel, then all_gather_into_tensor)',
                 timed(lambda: dist.all_gather_into_tensor(full, cspn_b200.propagate2d(g, d, s, N, '8sum'))), torch.equal(full, ref))
if B % 4 == 0:
    cg = ChunkedGather(B, 1, H, W, dev, n_chunks=4)
    ms = timed(lambda: cg.propagate(g, d, s, N, '8sum'))
    all_ok &= report('chunked NCCL (4 chunks, kernel i+1 overlaps gather i)', ms, torch.equal(cg.as_rank_major(), ref))
for mc in (True, False):
    try:
        fg = FusedGather(B, 1, H, W, dev, multicast=mc)
        ms = timed(lambda: fg.propagate(g, d, s, N, '8sum'))
        res = fg.propagate(g, d, s, N, '8sum')
        torch.cuda.synchronize()
        all_ok &= report(f'fused epilogue ({fg.mode})', ms, torch.equal(res, ref))
        if fg.mode == 'peer_stores':
            break
    except Exception as e:                                     # noqa: BLE001 -- report and go on with the other routes
        if rank == 0:
            print(f'fused epilogue (multicast={mc}) unavailable: {type(e).__name__}: {str(e)[:300]}', flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if all_ok else 1)
