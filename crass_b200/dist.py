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


def allgather_unique_tokens(records, first_read, n_unique, stride=64, group=None):
    """The same exchange, fed straight from the device de-duplication (K4b) without a host round trip per rank.

    records: uint8 tensor [>= n_unique*stride], first_read: int32 tensor [>= n_unique] (both on the rank's GPU, or on
    the CPU under gloo).  One all-gather of the counts and one of (records | first-read indices) padded to the largest
    count; every rank then orders each shard's records by first read, concatenates the shards in rank order and keeps
    first occurrences.  Returns the merged DR list as '\\n'-terminated text.
    """
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        rec = records[: max(n_unique, 1) * stride].cpu().numpy()
        fr = first_read[: max(n_unique, 1)].cpu().numpy().view("uint32")[:n_unique]
        return api.dr_list_from_unique(rec, stride, fr, raw=True)
    dev = records.device
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, torch.tensor([n_unique], dtype=torch.int64, device=dev), group=group)
    sizes = sizes.cpu().tolist()
    mx = max(max(sizes), 1)
    per = mx * (stride + 4)
    buf = torch.zeros(per, dtype=torch.uint8, device=dev)
    if n_unique:
        buf[: n_unique * stride] = records[: n_unique * stride]
        buf[mx * stride: mx * stride + 4 * n_unique] = first_read[:n_unique].contiguous().view(torch.uint8)
    allb = torch.empty(world * per, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allb, buf, group=group)
    host = allb.cpu().numpy()
    parts = []
    for r in range(world):
        if sizes[r]:
            rec = host[r * per: r * per + sizes[r] * stride]
            fr = host[r * per + mx * stride: r * per + mx * stride + 4 * sizes[r]].view("uint32")
            parts.append(api.dr_list_from_unique(rec, stride, fr, raw=True))
    return api.merge_dr_lists(b"".join(parts))


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
