"""Data-parallel propagation with the final gather fused into the kernel (one process per GPU, torchrun).

The reference's multi-GPU story is nn.DataParallel (eval.py:117): replicas compute their slice of the batch, then the
outputs are gathered.  Images are independent, so there is no collective on the data path; the only exchange is that
gather.  Three ways to do it, all ending with the full (world * B_local, C, H, W) output on every rank:

  FusedGather (this file)   the cluster kernel's epilogue stores every finished tile into the gather buffers of ALL
                            GPUs while it goes on computing the next tiles: `cspn2d_fwd_gather_f32` with peer pointers
                            of a CUDA symmetric-memory allocation (plain NVLink stores) or, where the fabric offers it,
                            one NVLS multicast address (a single `multimem.st`, replicated by the switch).  One kernel,
                            one barrier across ranks; the transfer hides behind the arithmetic.
  ChunkedGather             no symmetric memory needed: the local batch is cut into chunks, the kernel of chunk i+1
                            runs while NCCL all-gathers chunk i.
  sharding.gather_outputs   kernel, then one NCCL all_gather_into_tensor (what round 1 measured: the baseline).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .cspn import propagate2d


class FusedGather:
    """Owns the symmetric gather buffer of one problem shape.  `propagate(...)` returns a view of it holding all ranks'
    results, rank-major: out[r * B_local + i] is image i of rank r."""

    def __init__(self, B_local, C, H, W, device, group=None, multicast=True):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise _lib.CspnError('FusedGather addresses at most 8 GPUs of one NVLink domain')
        self.shape = (B_local, C, H, W)
        self.block = B_local * C * H * W                  # floats per rank
        self.buf = symm_mem.empty((self.world * B_local, C, H, W), dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        mine = self.rank * self.block * 4                 # byte offset of this rank's block inside every buffer
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.local = ptrs[self.rank] + mine
        mc = int(self.hdl.multicast_ptr) if (multicast and self.hdl.has_multicast_support) else 0
        if mc:
            self.mode = 'nvls_multicast'
            self.mc = mc + mine
            self.peers = []
        else:
            self.mode = 'peer_stores'
            self.mc = 0
            self.peers = [ptrs[r] + mine for r in range(self.world) if r != self.rank]
        self._peer_arr = (ctypes.c_void_p * max(1, len(self.peers)))(*self.peers)
        self._ws = None

    def propagate(self, guidance, blur_depth, sparse_depth=None, prop_time=24, norm_type='8sum'):
        B, C, H, W = blur_depth.shape
        if (B, C, H, W) != self.shape:
            raise ValueError(f'this FusedGather was built for {self.shape}, got {(B, C, H, W)}')
        L = _lib.lib()
        g, d = guidance.contiguous(), blur_depth.contiguous()
        s = None if sparse_depth is None else sparse_depth.contiguous()
        ws_bytes = L.cspn2d_workspace_bytes(B, C, H, W, int(prop_time), _lib.ALGO_CLUSTER)
        if ws_bytes and (self._ws is None or self._ws.numel() < ws_bytes):
            self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=d.device)
        rc = L.cspn2d_fwd_gather_f32(g.data_ptr(), d.data_ptr(), None if s is None else s.data_ptr(), self.local,
                                     ctypes.cast(self._peer_arr, ctypes.c_void_p), len(self.peers), self.mc or None,
                                     B, C, H, W, g.shape[1], int(prop_time), _lib.NORM2D[norm_type],
                                     None if not ws_bytes else self._ws.data_ptr(), ws_bytes,
                                     torch.cuda.current_stream(d.device).cuda_stream)
        _lib.check(rc, 'cspn2d_fwd_gather_f32')
        self.hdl.barrier()            # every rank's kernel (and with it its remote stores) is complete
        return self.buf


class ChunkedGather:
    """Kernel on chunk i+1 while NCCL moves chunk i.  Result layout is CHUNK-major: out[c][r] is chunk c of rank r
    (shape (n_chunks, world, B_chunk, C, H, W)); `as_rank_major()` gives a copy in FusedGather's order."""

    def __init__(self, B_local, C, H, W, device, n_chunks=4, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        if B_local % n_chunks:
            raise ValueError('n_chunks must divide the local batch')
        self.n_chunks, self.bc = n_chunks, B_local // n_chunks
        self.full = torch.empty(n_chunks, self.world, self.bc, C, H, W, dtype=torch.float32, device=device)
        self.local = torch.empty(B_local, C, H, W, dtype=torch.float32, device=device)

    def propagate(self, guidance, blur_depth, sparse_depth=None, prop_time=24, norm_type='8sum', algo=_lib.ALGO_AUTO):
        works = []
        for c in range(self.n_chunks):
            sl = slice(c * self.bc, (c + 1) * self.bc)
            oc = self.local[sl]
            propagate2d(guidance[sl], blur_depth[sl], None if sparse_depth is None else sparse_depth[sl], prop_time,
                        norm_type, algo, out=oc)
            # NCCL's stream waits for the kernel just enqueued; the next chunk's kernel does not wait for NCCL
            works.append(dist.all_gather_into_tensor(self.full[c], oc, group=self.group, async_op=True))
        for w in works:
            w.wait()
        return self.full

    def as_rank_major(self):
        return self.full.permute(1, 0, 2, 3, 4, 5).reshape(self.world * self.n_chunks * self.bc, *self.full.shape[3:])
