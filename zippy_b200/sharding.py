"""Multi-GPU sharding of a batch of independent members.

One process per GPU.  Members are partitioned by contiguous index range; every rank
compresses (or inflates) its own range with no data-path collective.  The path's single
exchange is one all_gather of the per-member output sizes, after which every rank knows
where each member lands in the concatenated stream.  Works with any torch.distributed
backend (NCCL on the GPUs; gloo in the CPU tests)."""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of member indices for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_range_by_bytes(lens, rank, world):
    """Contiguous ranges balanced by input bytes (for skewed member sizes)."""
    lens = np.asarray(lens, dtype=np.uint64)
    csum = np.concatenate([[0], np.cumsum(lens)])
    total = int(csum[-1])
    cuts = [int(np.searchsorted(csum, total * r // world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, len(lens)
    for i in range(1, world + 1):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts[rank], cuts[rank + 1]


def gather_sizes(local_sizes, n_total, device=None, group=None, on_device=False):
    """all_gather the per-member output sizes of every rank: the path's one exchange.

    local_sizes: int64 array (or tensor on `device`) for this rank's shard (shard_range order).  Returns
    (all_sizes int64[n_total], global_offsets int64[n_total + 1]) as numpy arrays, or -- on_device=True -- as
    tensors on `device` without a host synchronisation (a rank that only needs to know where its shard lands
    reads two elements later).  Shards may differ in length by one, so sizes are padded to the longest shard
    for the collective; equal shards (n_total divisible by the world size) skip the padding."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        if on_device:
            ls = torch.as_tensor(np.asarray(local_sizes, dtype=np.int64) if not torch.is_tensor(local_sizes) else local_sizes,
                                 device=device)
            return ls, torch.cat([ls.new_zeros(1), torch.cumsum(ls, 0)])
        sizes = np.asarray(local_sizes, dtype=np.int64)
        return sizes, np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    longest = (n_total + world - 1) // world
    ls = local_sizes if torch.is_tensor(local_sizes) else torch.as_tensor(np.asarray(local_sizes, dtype=np.int64), device=device)
    ls = ls.to(device=device, dtype=torch.int64)
    even = n_total % world == 0
    if even:
        pad = ls
    else:
        pad = torch.full((longest,), -1, dtype=torch.int64, device=device)
        pad[:ls.numel()] = ls
    out = torch.empty(world * longest, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if on_device:
        if not even:
            keep = torch.cat([torch.arange(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0], device=device)
                              + r * longest for r in range(world)])
            out = out[keep]
        return out, torch.cat([out.new_zeros(1), torch.cumsum(out, 0)])
    out = out.cpu().numpy().reshape(world, longest)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(out[r, :hi - lo])
    sizes = np.concatenate(parts)
    assert (sizes >= 0).all()
    return sizes, np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_node(pci_bus_id):
    """NUMA node of a GPU from sysfs (pci_bus_id like '0000:1b:00.0'); None when unknown."""
    import os
    for cand in (pci_bus_id.lower(), pci_bus_id.lower()[-12:], "0000:" + pci_bus_id.lower()[-7:]):
        p = "/sys/bus/pci/devices/%s/numa_node" % cand
        if os.path.exists(p):
            try:
                node = int(open(p).read().strip())
                return node if node >= 0 else None
            except (OSError, ValueError):
                return None
    return None


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is
    allocated: page-locked memory is then first-touched on that node and the H2D / D2H copies of the
    host-buffer calls do not cross the inter-socket link (with 8 ranks on a two-socket host the far
    half otherwise loses a third of its copy bandwidth).  Returns a description dict; never raises."""
    import os
    info = {"device": device_index, "node": None, "cpus": None, "bound": False}
    try:
        import torch
        prop = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        info["pci"] = bus
        node = gpu_numa_node(bus)
        info["node"] = node
        if node is None:
            return info
        cpus = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return info
        os.sched_setaffinity(0, use)
        info["cpus"] = len(use)
        info["bound"] = True
    except Exception as e:  # sysfs layout differs, no permission, ...: run unbound
        info["error"] = repr(e)
    return info
