"""The one exchange step of the multi-GPU path: merging the per-shard direct-repeat sets.

Reads are sharded contiguously, one process per GPU.  After phase 1 every rank holds the ordered list of the
distinct low-lexi DR strings of its shard (first-appearance order).  One all-gather of the sizes and one of the
padded byte records (NCCL over NVLink on GPUs, gloo in the CPU tests) give every rank all lists; concatenating them in
rank order and keeping first occurrences reproduces exactly the token order a single sequential run would have
produced (StringCheck numbers tokens by first appearance, StringCheck.cpp:46-55), so clustering, the non-redundant
pattern set and the automaton come out identical on every rank without any further communication.
"""
import torch
import torch.distributed as dist

from . import api


def allgather_dr_lists(local, device=None, group=None):
    """local: list of bytes (this rank's distinct DRs in first-appearance order) -> merged global list."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(local)
    world = dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    blob = b"".join(d + b"\n" for d in local)
    size = torch.tensor([len(blob)], dtype=torch.int64, device=dev)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, size, group=group)                 # counts ...
    sizes = sizes.cpu().tolist()
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if blob:
        buf[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    allb = torch.empty(world * mx, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allb, buf, group=group)                   # ... then the padded DR records
    allb = allb.cpu().numpy()
    drs_all = []
    for r in range(world):
        drs_all += [x for x in allb[r * mx: r * mx + sizes[r]].tobytes().split(b"\n") if x]
    return api.merge_dr_lists(drs_all)                                    # rank-ordered first appearance
