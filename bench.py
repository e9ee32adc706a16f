#!/usr/bin/env python
"""bench.py -- CSPN propagation throughput on B200 (contract: see the task brief / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--algo auto|generic|cluster]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (24-iteration 2D CSPN, '8sum', with sparse depth) over one batch of
synthetic inputs of BASELINE.json configs[1]: 32 x 1216x352 (W x H) per GPU, fp32.  Weak scaling: every
rank owns its own 32 images (N=8 is configs[4], 256 images); no data-path collective.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, B_PER_GPU, ITERS, NORM = 352, 1216, 32, 24, '8sum'
ALGO_BYTES_PER_PX = 44          # read 8 guidance + blur + sparse, write out, fp32
METRIC = 'CSPN Mpixels/s (24-iter 2D, 1216x352)'


def measured_hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json, burst copy)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def recorded_traffic(algo_name):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            t = json.load(fh)
        return t.get(algo_name, {}).get('dram_bytes_per_launch')
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (pynvml, ~2 ms period)."""
    REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
               0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != 'gpu_idle':
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': (s[len(s) // 2] if s else None), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        try:
            return int(vis.split(',')[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ---------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------

def time_torch_port(nb, steps, warmup, threads):
    """The reference's own op sequence (oracle/cspn_torch_port.py == cspn.py minus `.cuda()`), all host threads."""
    import torch
    from cspn_b200.synth import make_inputs
    from oracle import cspn_torch_port as tp
    torch.set_num_threads(threads)
    g, d, s = make_inputs(0, nb, 1, H, W)
    with torch.no_grad():
        for _ in range(warmup):
            tp.cspn2d_torch(g, d, s, ITERS, NORM)
        t0 = time.perf_counter()
        for _ in range(steps):
            tp.cspn2d_torch(g, d, s, ITERS, NORM)
        dt = time.perf_counter() - t0
    return nb * H * W * steps / dt / 1e6, dt / steps


def best_torch_threads():
    """The reference's op sequence is ~700 small ATen ops per forward: oversubscribing a 128-thread host makes it
    20x slower than a few threads.  'All the host threads it can use' = the count that maximises its throughput."""
    ncpu = os.cpu_count() or 1
    best, best_t = 1, float('inf')
    for n in sorted({min(ncpu, c) for c in (4, 8, 16, 32, 64, ncpu)}):
        t = time_torch_port(1, 1, 1 if n == 4 else 0, n)[1]
        if t < best_t:
            best, best_t = n, t
    return best, best_t


def cpu_baseline_leg():
    """Bounded sample (about 10-30 s): torch-op port of the reference, plus the C/OpenMP oracle for context."""
    import torch
    from cspn_b200.synth import make_inputs
    from oracle import c_oracle
    threads, t1 = best_torch_threads()
    nb = max(1, min(8, int(12.0 / max(t1, 1e-3) / 3)))            # ~12 s over 1 warm-up + 2 timed forwards
    mpx, per = time_torch_port(nb, 2, 1, threads)
    g, d, s = make_inputs(0, 8, 1, H, W)
    gn, dn, sn = g.numpy(), d.numpy(), s.numpy()
    ncpu = os.cpu_count() or 1
    c_mpx, c_thr = 0.0, 1
    for nt in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):      # the OpenMP port also has a sweet spot
        c_oracle.cspn2d(gn[:1], dn[:1], sn[:1], ITERS, NORM, nthreads=nt)
        t0 = time.perf_counter()
        c_oracle.cspn2d(gn, dn, sn, ITERS, NORM, nthreads=nt)
        r = 8 * H * W / (time.perf_counter() - t0) / 1e6
        if r > c_mpx:
            c_mpx, c_thr = r, nt
    # for context only: the same reference op sequence run eagerly on this GPU (the reference's native mode:
    # ~700 library-kernel launches per forward), 4 images
    gpu_eager = None
    try:
        from oracle import cspn_torch_port as tp
        if torch.cuda.is_available():
            gg, dd, ss = [t[:4].cuda() for t in (g, d, s)]
            with torch.no_grad():
                tp.cspn2d_torch(gg, dd, ss, ITERS, NORM)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(3):
                    tp.cspn2d_torch(gg, dd, ss, ITERS, NORM)
                torch.cuda.synchronize()
            gpu_eager = round(3 * 4 * H * W / (time.perf_counter() - t0) / 1e6, 1)
    except Exception:
        gpu_eager = None
    return {'value': round(mpx, 3), 'unit': 'Mpixels/s', 'cores': threads, 'kind': 'port',
            'reference_ops_eager_on_this_gpu_mpx_s': gpu_eager,
            'sample': f'{nb}x{W}x{H} images, {ITERS} iters, 1 warm-up + 2 timed forwards of oracle/cspn_torch_port.py '
                      f'(the reference op sequence of cspn.py:42-83 on CPU; /root/reference is absent on this box); '
                      f'{threads} torch threads = the fastest of 4..{os.cpu_count()} on this host',
            'host_cpus': os.cpu_count(),
            'c_openmp_port_mpx_s': round(c_mpx, 3), 'c_openmp_threads': c_thr}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (torch-op port), rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads, _ = best_torch_threads()
    budget = 150.0                                               # seconds for the whole run
    t_img = time_torch_port(4, 1, 1, threads)[1] / 4             # per-image time at a batch that no longer fits the caches
    nb = max(1, min(B_PER_GPU, int(budget / ((args.steps + args.warmup) * max(t_img, 1e-3)))))
    mpx, per = time_torch_port(nb, args.steps, args.warmup, threads)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': round(mpx, 3), 'unit': 'Mpixels/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(per * 1e3, 3), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'2D CSPN 3x3, {ITERS} iters, {NORM}, batch {B_PER_GPU}x{W}x{H} per GPU (BASELINE configs[1])',
                   'sample_batch': nb},
        'cpu_baseline': {'value': round(mpx, 3), 'unit': 'Mpixels/s', 'cores': threads, 'kind': 'port',
                         'sample': f'each step = {nb}x{W}x{H} images through oracle/cspn_torch_port.py (reference op '
                                   f'sequence, cspn.py:42-83) with {threads} torch threads (fastest of 4..{os.cpu_count()})', 'host_cpus': os.cpu_count()},
        'e2e': {'value': round(mpx, 3), 'unit': 'Mpixels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner comes from C code)
    is diverted to stderr for the duration of the run."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + '\n').encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--algo', default='auto', choices=['auto', 'generic', 'cluster'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    quiet_stdout()

    if args.impl == 'reference':
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import cspn_b200
    from cspn_b200 import _lib
    from cspn_b200.synth import make_inputs

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run --nproc-per-node N for --gpus N')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback exists)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # keep stdout to the single JSON line: NCCL's version banner / debug log goes to stderr
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)

    algo = {'auto': _lib.ALGO_AUTO, 'generic': _lib.ALGO_GENERIC, 'cluster': _lib.ALGO_CLUSTER}[args.algo]
    L = _lib.lib()

    # per-rank shard: its own 32 images (in the real pipeline the UNet produces them on this GPU)
    g_h, d_h, s_h = make_inputs(rank, B_PER_GPU, 1, H, W)
    g, d, s = g_h.to(dev), d_h.to(dev), s_h.to(dev)
    px = B_PER_GPU * H * W

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return cspn_b200.propagate2d(g, d, s, ITERS, NORM, algo)

    for _ in range(args.warmup):
        out = step()
    launches_per_step = L.cspn_last_launches()
    algo_used = _lib.ALGO_NAMES[L.cspn_last_algo()]
    plan = cspn_b200.describe_plan(B_PER_GPU, 1, H, W, ITERS, algo)

    # ---- timed region: K steps, inputs resident in HBM (602.7 MB per step > 126 MB L2) -------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(physical_gpu_index(local_rank)) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            out = step()
        ev1.record()
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    value = world * px / (ms_per_step * 1e-3) / 1e6

    # ---- dominant-kernel roofline: events around each launch sequence of one step, averaged --------
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for a, b in evs:
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    kern_ms = sorted(a.elapsed_time(b) for a, b in evs)
    kern_ms_avg = sum(kern_ms) / len(kern_ms)
    peak, peak_src = measured_hbm_peak()
    achieved = ALGO_BYTES_PER_PX * px / (kern_ms_avg * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'achieved': round(achieved, 1), 'peak': peak, 'unit': 'GB/s',
                'frac': round(achieved / peak, 4), 'traffic': recorded_traffic(algo_used),
                'peak_source': peak_src, 'kernel': f'{algo_used} ({launches_per_step} launch(es) per step)',
                'algorithmic_bytes_per_launch_sequence': ALGO_BYTES_PER_PX * px,
                'launch_ms_avg': round(kern_ms_avg, 4), 'launch_ms_min': round(kern_ms[0], 4)}

    # ---- e2e: the reference-facing call with HOST buffers (pinned), H2D + kernels + D2H inside ------
    gp, dp, sp = g_h.pin_memory(), d_h.pin_memory(), s_h.pin_memory()
    out_h = torch.empty_like(dp).pin_memory()                                  # the serving loop's result buffer
    for _ in range(2):                                                         # warm-up (allocates pipeline slots)
        cspn_b200.propagate2d(gp, dp, sp, ITERS, NORM, algo, out=out_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        cspn_b200.propagate2d(gp, dp, sp, ITERS, NORM, algo, out=out_h)         # H2D + kernels + D2H, blocking
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_launches = L.cspn_last_launches()
    e2e = {'value': round(world * px / float(e2e_s.item()) / 1e6, 1), 'unit': 'Mpixels/s',
           'h2d_bytes_per_step': world * 10 * px * 4, 'd2h_bytes_per_step': world * px * 4,
           'ms_per_step': round(float(e2e_s.item()) * 1e3, 3), 'launches_per_step': e2e_launches,
           'api': 'cspn_b200.propagate2d(cpu pinned tensors) -> C ABI cspn2d_fwd_f32_host'}
    same = bool(torch.equal(out_h, out.cpu()))

    # ---- optional: what the north_star calls "NCCL only for the final gather" (outside the timed value) ----
    gather = None
    if world > 1:
        full = torch.empty(world * B_PER_GPU, 1, H, W, device=dev)
        for _ in range(5):
            dist.all_gather_into_tensor(full, out)
        barrier()
        ev0.record()
        for _ in range(5):
            o = step()
            dist.all_gather_into_tensor(full, o)
        ev1.record()
        barrier()
        gms = torch.tensor([ev0.elapsed_time(ev1) / 5], device=dev)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        gather = {'ms_per_step_with_all_gather': round(float(gms.item()), 4),
                  'mpx_s_with_all_gather': round(world * px / (float(gms.item()) * 1e-3) / 1e6, 1),
                  'bytes_received_per_rank': (world - 1) * px * 4}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg()

    if rank == 0:
        line = {
            'metric': METRIC, 'value': round(value, 1), 'unit': 'Mpixels/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': round(ms_per_step, 4), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'2D CSPN 3x3, {ITERS} iters, {NORM}, with sparse depth, batch {B_PER_GPU}x{W}x{H} '
                                   f'per GPU (BASELINE configs[1]; N=8 is configs[4])',
                       'global_batch': world * B_PER_GPU, 'parallelism': f'dp{world} (independent images per rank)',
                       'l2': 'inputs per step (602.7 MB/GPU) exceed the 126 MB L2; no flush needed',
                       'algo': algo_used, 'plan': plan},
            'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e,
            'gpu_launches': world * args.steps * launches_per_step,
            'clocks': clocks.summary(), 'e2e_matches_device_path': same,
        }
        if gather:
            line['gather'] = gather
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
