"""Host-side helpers for the multi-GPU layout of the path: frames are sharded over ranks with no data-path collective
in inference; training adds ONE gradient all-reduce per step over a single flat buffer (NCCL over NVLink on the GPU boxes, gloo in
the CPU tests) and the naiveSyncBN statistics exchange (sst_b200/norm.py).  Pure torch.distributed plumbing - no kernels here."""
import os

import torch
import torch.distributed as dist


def pin_to_gpu_numa(device_index):
    """Bind this process to the CPUs that are local to GPU `device_index` (the PCI device's `local_cpulist` in sysfs), so the
    host thread that feeds the engine and the pinned staging buffers it allocates afterwards live on the GPU's NUMA node.
    Returns the CPU set, or None when the topology is not exposed (containers without sysfs PCI info): never fatal."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        path = f"/sys/bus/pci/devices/{bdf.lower()}/local_cpulist"
        if not os.path.exists(path):
            return None
        cpus = set()
        for part in open(path).read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_range(num_frames, rank, world):
    """Contiguous frame range [lo, hi) of `rank` (B/n frames each, the remainder spread over the first ranks) - the partition
    the reference's DistributedSampler-style loaders produce for inference (apis/test.py:133-139 gathers host-side)."""
    base, rem = divmod(num_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device=None):
    """A timing is the MAX over ranks (bench contract): all-reduce of one float64."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_grads_flat(params, dtype=None):
    """DDP-style gradient averaging with ONE collective: gradients are packed into a single flat buffer (optionally cast, e.g.
    to bf16, for the wire), summed over the ranks, divided by the world size and unpacked in place.  Parameters without a
    gradient contribute zeros (every rank must pack the same layout)."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    dev = params[0].device
    wire = dtype or torch.float32
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(wire) for p in params])
    if world > 1:
        dist.all_reduce(flat)
        flat /= world
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    assert flat.device == dev
    return flat.numel()
