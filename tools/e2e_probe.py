"""PCIe copy bandwidth + end-to-end (host buffers) timing of the C ABI host entry point."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cspn_b200
from cspn_b200.synth import make_inputs
x = torch.empty(512 << 20, dtype=torch.uint8).pin_memory(); y = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
for name, fn in (('H2D', lambda: y.copy_(x, non_blocking=True)), ('D2H', lambda: x.copy_(y, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f'{name} 512 MiB pinned: {0.5369 / dt:.1f} GB/s')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
x2 = torch.empty(512 << 20, dtype=torch.uint8).pin_memory(); y2 = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): y.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): x2.copy_(y2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print(f'H2D + D2H concurrently: {0.5369 / dt:.1f} GB/s each direction')
g, d, s = make_inputs(0, 32, 1, 352, 1216)
gp, dp, sp = g.pin_memory(), d.pin_memory(), s.pin_memory()
outp = torch.empty_like(dp).pin_memory()
for _ in range(2): cspn_b200.propagate2d(gp, dp, sp, 24, '8sum', out=outp)
t = time.perf_counter()
for _ in range(5): out = cspn_b200.propagate2d(gp, dp, sp, 24, '8sum', out=outp)
dt = (time.perf_counter() - t) / 5
print(f'CSPN_B200_HOST_CHUNK_MB={os.environ.get("CSPN_B200_HOST_CHUNK_MB", "default")}: e2e {dt*1e3:.2f} ms  {32*352*1216/dt/1e6:.0f} Mpx/s  ({0.6027/dt:.1f} GB/s moved)')
# raw torch copies of the same tensors (upper bound for any host pipeline)
gd, dd, sd = torch.empty_like(gp, device='cuda'), torch.empty_like(dp, device='cuda'), torch.empty_like(sp, device='cuda')
oh = torch.empty_like(dp).pin_memory()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3):
    gd.copy_(gp, non_blocking=True); dd.copy_(dp, non_blocking=True); sd.copy_(sp, non_blocking=True)
    torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 3
print(f'raw H2D of guidance+blur+sparse (548 MB, 3 copies, one stream): {dt*1e3:.2f} ms  {0.5479/dt:.1f} GB/s')
t = time.perf_counter()
for _ in range(3):
    for b0 in range(0, 32, 2):
        gd[b0:b0+2].copy_(gp[b0:b0+2], non_blocking=True); dd[b0:b0+2].copy_(dp[b0:b0+2], non_blocking=True); sd[b0:b0+2].copy_(sp[b0:b0+2], non_blocking=True)
    torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 3
print(f'same in 16 chunks of 2 images (48 copies, one stream): {dt*1e3:.2f} ms  {0.5479/dt:.1f} GB/s')
