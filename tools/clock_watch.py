"""Runs one shape back to back for ~2 s while sampling SM clock / power / throttle reasons: python tools/clock_watch.py B H W N"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
import cspn_b200
from cspn_b200.synth import make_inputs
B, H, W, N = [int(a) for a in sys.argv[1:5]]
g, d, s = [t.cuda() for t in make_inputs(0, min(B, 8), 1, H, W)]
if B > 8:
    rep = (B + 7) // 8
    g, d, s = [t.repeat(rep, 1, 1, 1)[:B].contiguous() for t in (g, d, s)]
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = threading.Event()
def loop():
    while not stop.is_set():
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.02)
for _ in range(3): cspn_b200.propagate2d(g, d, s, N, '8sum', 2)
torch.cuda.synchronize()
th = threading.Thread(target=loop); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3000
e0.record()
for _ in range(reps): cspn_b200.propagate2d(g, d, s, N, '8sum', 2)
e1.record(); torch.cuda.synchronize()
stop.set(); th.join()
clk = sorted(c for c, p, r in samples); pw = sorted(p for c, p, r in samples)
reasons = 0
for c, p, r in samples: reasons |= r
print(f'B={B} H={H} W={W} N={N}: {e0.elapsed_time(e1)/reps*1e3:.1f} us/launch over {reps} launches; sm clock min/median/max = '
      f'{clk[0]}/{clk[len(clk)//2]}/{clk[-1]} MHz, power median/max = {pw[len(pw)//2]:.0f}/{pw[-1]:.0f} W, reasons mask = {reasons:#x}, samples={len(samples)}')
