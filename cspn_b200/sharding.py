"""Batch sharding for one-process-per-GPU runs (torchrun): images are independent (nothing in
/root/reference/cspn_pytorch/models/cspn.py couples the batch dimension), so a rank propagates its contiguous
slice of the batch with no collective on the data path; the only exchange is the optional final gather of outputs
(what nn.DataParallel's gather does in the reference, eval.py:117)."""
import torch
import torch.distributed as dist


def shard_range(batch, world_size, rank):
    """Contiguous split of `batch` items over `world_size` ranks; earlier ranks take the remainder. -> (start, count)"""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, rem = divmod(batch, world_size)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def gather_outputs(local_out, batch, group=None):
    """All ranks receive the full (batch, C, H, W) output.  Equal shards use all_gather_into_tensor (NCCL's fast
    path); ragged shards fall back to padded all_gather.  Works on gloo (CPU tensors) and nccl (CUDA tensors)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [shard_range(batch, world, r)[1] for r in range(world)]
    assert local_out.shape[0] == counts[rank], (local_out.shape, counts, rank)
    rest = tuple(local_out.shape[1:])
    if len(set(counts)) == 1 and dist.get_backend(group) == 'nccl':
        full = torch.empty((batch,) + rest, dtype=local_out.dtype, device=local_out.device)
        dist.all_gather_into_tensor(full, local_out.contiguous(), group=group)
        return full
    mx = max(counts)
    padded = torch.zeros((mx,) + rest, dtype=local_out.dtype, device=local_out.device)
    padded[:counts[rank]] = local_out
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)


def bind_to_gpu_numa(device_index):
    """Pins the calling process to the CPUs of the NUMA node its GPU hangs off (and, by first touch, the pinned host
    buffers it allocates afterwards).  Eight ranks each pushing 50 GB/s of pinned H2D from whatever socket the scheduler
    put them on lose a third of that bandwidth (round 1: e2e weak-scaling efficiency 0.65 at 8 GPUs).  Returns a short
    description, or None when the topology cannot be read (containers without sysfs PCI entries): then nothing changes."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f'{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{bdf}/numa_node') as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f'/sys/devices/system/node/node{node}/cpulist') as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f'numa node {node} ({len(cpus)} cpus) for GPU {bdf}'
    except Exception:
        return None
