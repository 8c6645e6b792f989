"""Multi-GPU sharding of the hot path: one process per GPU, no data-path collective.

* single-channel input (cfg1, cfg2, cfg4): the stream is cut into super-blocks of overlap-save
  parts, rank g takes parts [first, first+n) and re-reads nsamp_overlap samples at its left edge --
  what the reference's threads do through Input overlap (MultiThread.C:120-148);
  the private PhaseSeries are summed at sub-integration boundaries (PhaseSeries::combine,
  PhaseSeries.C:442-480): data +=, hits +=, integration_length +=, ndat_total +=.
* multi-channel input (cfg3, cfg5): contiguous channel ranges per rank, every rank folds with the
  same bin plan; the disjoint [chan][pol][bin][dim] blocks are gathered (no arithmetic).

torch.distributed provides the plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_parts(total_parts, world, rank):
    """Contiguous, balanced split of `total_parts` overlap-save parts: -> (first_part, nparts)."""
    base, extra = divmod(total_parts, world)
    n = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, n


def shard_channels(nchan, world, rank):
    """Contiguous, balanced split of `nchan` input channels: -> (first_chan, nchan_local)."""
    return shard_parts(nchan, world, rank)


def part_byte_range(first_part, nparts, nsamp_step, nsamp_overlap, bytes_per_sample):
    """Byte range [lo, hi) of the raw stream a rank must read for its parts (overlap re-read included)."""
    lo = first_part * nsamp_step * bytes_per_sample
    hi = ((first_part + nparts) * nsamp_step + nsamp_overlap) * bytes_per_sample
    return lo, hi


def combine_time_sharded(profile, hits, integration_length, ndat_total, dst=0, group=None):
    """PhaseSeries::combine across ranks (time sharding): profile/hits tensors are summed onto `dst`
    in rank order by the collective; the scalar attributes are summed on the host side."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return integration_length, ndat_total
    dist.reduce(profile, dst, op=dist.ReduceOp.SUM, group=group)
    dist.reduce(hits, dst, op=dist.ReduceOp.SUM, group=group)
    s = torch.tensor([float(integration_length), float(ndat_total)], dtype=torch.float64, device=profile.device)
    dist.reduce(s, dst, op=dist.ReduceOp.SUM, group=group)
    return float(s[0].item()), int(round(s[1].item()))


def gather_channel_sharded(profile_local, nchan_total, dst=0, group=None):
    """Concatenate disjoint channel shards [nchan_local, ...] on `dst` (no arithmetic).
    Shards may be ragged (nchan not divisible by the world size)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return profile_local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    tail = tuple(profile_local.shape[1:])
    nmax = -(-nchan_total // world)
    padded = torch.zeros((nmax,) + tail, dtype=profile_local.dtype, device=profile_local.device)
    padded[: profile_local.shape[0]] = profile_local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        _, n = shard_channels(nchan_total, world, r)
        out.append(bufs[r][:n])
    return torch.cat(out, dim=0)


def bind_cpu_affinity(device_index):
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index` (the role of dspsr's
    `-cpu` option next to `-cuda`, dspsr.C / SingleThread.C set_affinity), so that the pinned staging buffers
    allocated afterwards are first touched on the GPU's NUMA node.  With one process per GPU on a multi-socket
    host this keeps every host-to-device stream off the inter-socket link.  Returns the number of cores bound
    (0: NVML unavailable or no usable mask -- the affinity is then left alone)."""
    import os
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        h = nv.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (dom, bus, dev)).encode())
        ncpu = os.cpu_count() or 1
        mask = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1} & os.sched_getaffinity(0)
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0
