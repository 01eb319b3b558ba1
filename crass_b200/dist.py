"""The one exchange step of the multi-GPU path: merging the per-shard direct-repeat sets.

Reads are sharded contiguously, one process per GPU.  After phase 1 every rank holds the distinct low-lexi DR tokens
of its shard with the read each first appeared in.  Concatenating the shards' tokens in rank order and keeping first
occurrences reproduces exactly the token order a single sequential run would have produced (StringCheck numbers
tokens by first appearance, StringCheck.cpp:46-55), so clustering, the non-redundant pattern set and the matcher come
out identical wherever they are computed.

  TokenExchange     token blocks (K4b) -> one NCCL all-gather -> merged on every GPU (K4c) -> every rank clusters
  PatternExchange   the same gather, but only the root merges and clusters; one NCCL broadcast returns the pattern set
  allgather_unique_tokens / allgather_dr_lists   the host-merged forms (any backend; gloo in the CPU tests)
"""
import torch
import torch.distributed as dist

import os

from . import api

_HOST_PASSES = os.environ.get("CRASS_B200_CLUSTER", "") == "host"   # comparison knob: clustering passes A/B on the host


class TokenExchange:
    """The exchange on the GPUs: K4b writes this shard's distinct DR tokens into a fixed-size token block, one NCCL
    all-gather delivers every rank's block, K4c de-duplicates them with global first-appearance keys, and one
    device-to-host copy brings the merged block back (one host synchronisation in total; the host only sorts the
    merged records).  At world size 1 the all-gather and K4c drop out.  Blocks grow (x2) and the step is repeated if
    a shard ever has more distinct tokens than fit; every rank sees the same merged header, so all ranks decide alike.
    """

    def __init__(self, ctx, device, shard_reads, stride=64, cap=16384, group=None):
        self.ctx, self.dev, self.stride, self.group = ctx, torch.device(device), stride, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.shard_reads = int(shard_reads)
        self._alloc(cap)
        self.own_comm = False
        if self.world > 1 and group is None and self.dev.type == "cuda" and os.environ.get("CRASS_B200_EXCHANGE", "") != "torch":
            self._init_own_comm()

    def _init_own_comm(self):
        """The library's own NCCL communicator (libnccl is dlopen'ed by it): K4b, the all-gather and K4c then go onto the caller's
        stream as three enqueues of one C call -- through torch.distributed the collective alone costs 0.1 ms of host time per
        step and hops to NCCL's stream and back.  The unique id travels over the process group that is there anyway."""
        rank = dist.get_rank()
        if self.ctx.comm_world == self.world:                           # an earlier exchange on this context made it (on every rank alike)
            self.own_comm = True
            return
        try:
            uid = api.Context.comm_unique_id() if rank == 0 else bytes(128)
        except api.CrassB200Error:
            uid = None
        box = [uid]
        dist.broadcast_object_list(box, src=0)
        if box[0] is None:
            return
        try:
            self.ctx.comm_init(box[0], rank, self.world)
            ok = 1
        except api.CrassB200Error:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)                      # all ranks or none
        self.own_comm = bool(int(flag.item()))

    def _alloc(self, cap):
        self.cap = cap
        self.out_cap = cap * min(self.world, 4) if self.world > 1 else cap
        nb = api.token_block_bytes(cap, self.stride)
        self.send = torch.empty(nb, dtype=torch.uint8, device=self.dev)
        if self.world > 1:
            self.recv = torch.empty(self.world * nb, dtype=torch.uint8, device=self.dev)
            self.merged = torch.empty(api.token_block_bytes(self.out_cap, self.stride), dtype=torch.uint8, device=self.dev)
        else:
            self.merged = self.send
        self.host = torch.empty(self.merged.numel(), dtype=torch.uint8, pin_memory=True)

    def _stream(self, stream):
        """The exchange's kernels, the NCCL calls and the host copies must be ordered on ONE stream: torch's current
        stream of the device unless the caller names another (and has made it torch's current stream)."""
        return torch.cuda.current_stream(self.dev).cuda_stream if stream is None else stream

    def run(self, d_hits, n_hits, d_tokens, stream=None):
        """-> (merged DR list as '\\n'-terminated text, number of distinct tokens of this shard or None when N > 1)"""
        stream = self._stream(stream)
        while True:
            self.ctx.unique_tokens_block_dev(d_hits, n_hits, d_tokens, self.stride, self.send, self.cap, stream)
            if self.world > 1:
                dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
                self.ctx.merge_token_blocks_dev(self.recv, self.world, self.cap, self.stride, self.shard_reads, self.merged, self.out_cap, stream)
            self.host.copy_(self.merged, non_blocking=True)
            torch.cuda.current_stream(self.dev).synchronize()
            text, count, flags = api.dr_list_from_block(self.host, self.out_cap, self.stride)
            if flags & 2:
                raise api.CrassB200Error(api.EINVAL, "token stride too small for the DR lengths in use")
            if (flags & 1) or count > self.out_cap:
                self._alloc(self.cap * 2)
                continue
            return text, count

    def run_matcher(self, d_hits, n_hits, d_tokens, kmer_clust=6, stream=None):
        """The same exchange, ending in createNonRedundantSet + matcher build straight from the merged block
        -> (api.Automaton or None when there is no DR, number of distinct DR variants)."""
        import time
        stream = self._stream(stream)
        while True:
            t0 = time.perf_counter()
            if self.world > 1 and self.own_comm:
                self.ctx.exchange_tokens_dev(d_hits, n_hits, d_tokens, self.stride, self.send, self.cap, self.recv, self.shard_reads,
                                             self.merged, self.out_cap, stream)
            else:
                self.ctx.unique_tokens_block_dev(d_hits, n_hits, d_tokens, self.stride, self.send, self.cap, stream)
                if self.world > 1:
                    dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
                    self.ctx.merge_token_blocks_dev(self.recv, self.world, self.cap, self.stride, self.shard_reads, self.merged, self.out_cap, stream)
            if _HOST_PASSES:                                      # CRASS_B200_CLUSTER=host: all clustering passes on the host
                self.host.copy_(self.merged, non_blocking=True)
                torch.cuda.current_stream(self.dev).synchronize()
                t1 = time.perf_counter()
                ac, count, flags = api.Automaton.from_block(self.host, self.out_cap, self.stride, kmer_clust)
            else:                                                 # K5: token order, k-mer keys and first holders on the GPU
                t1 = time.perf_counter()
                ac, count, flags = self.ctx.cluster_block_dev(self.merged, self.out_cap, self.stride, kmer_clust, stream)
            self.last_ms = {"tokens_to_host": (t1 - t0) * 1e3, "cluster_build": (time.perf_counter() - t1) * 1e3}
            if flags & 2:
                raise api.CrassB200Error(api.EINVAL, "token stride too small for the DR lengths in use")
            if (flags & 1) or count > self.out_cap:
                self._alloc(self.cap * 2)
                continue
            return ac, count


class PatternExchange(TokenExchange):
    """The step between the phases for N > 1 with the serial part done once.

    createNonRedundantSet is a serial section of the job: running the identical clustering on every rank does not make
    it faster, it only divides the box's cores between N copies.  Here the token blocks are gathered as in
    TokenExchange, but only the root merges them (K4c), brings the DR list to its host and clusters it with all the
    helper threads it is given; the resulting pattern set (a few hundred KB of text) goes back to every GPU with one
    NCCL broadcast of a fixed-size message [u32 status, u32 n_variants, u64 text_len | text], and every rank builds its
    matcher from it.  Status 1 / 2 tell all ranks alike to repeat the round with larger token blocks / a larger message.
    """

    HEADER = 16

    def __init__(self, ctx, device, shard_reads, kmer_clust=6, stride=64, cap=16384, text_cap=1 << 18, root=0, group=None):
        super().__init__(ctx, device, shard_reads, stride, cap, group)
        self.kmer_clust, self.root = kmer_clust, root
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._alloc_msg(text_cap)

    def _alloc_msg(self, text_cap):
        self.text_cap = text_cap
        self.msg = torch.zeros(self.HEADER + text_cap, dtype=torch.uint8, device=self.dev)
        self.msg_host = torch.zeros(self.HEADER + text_cap, dtype=torch.uint8, pin_memory=True)

    def run(self, d_hits, n_hits, d_tokens, stream=None):
        """-> (pattern set as '\\n'-terminated text, number of distinct DR variants over all ranks)"""
        stream = self._stream(stream)
        if self.world == 1:
            text, count = super().run(d_hits, n_hits, d_tokens, stream)
            return (api.non_redundant_patterns(text, self.kmer_clust) if text else b""), count
        import time
        import numpy as np
        while True:
            t0 = time.perf_counter()
            self.ctx.unique_tokens_block_dev(d_hits, n_hits, d_tokens, self.stride, self.send, self.cap, stream)
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
            hdr = self.msg_host[: self.HEADER].numpy()
            if self.rank == self.root:
                # whatever goes wrong on the root travels as status 3 + message: the other ranks are already waiting in
                # the broadcast below and must hear about it instead of hanging there
                status, count, text, t1 = 0, 0, b"", t0
                try:
                    self.ctx.merge_token_blocks_dev(self.recv, self.world, self.cap, self.stride, self.shard_reads, self.merged, self.out_cap, stream)
                    if _HOST_PASSES:
                        self.host.copy_(self.merged, non_blocking=True)
                        torch.cuda.current_stream(self.dev).synchronize()
                        t1 = time.perf_counter()
                        text, count, flags = api.non_redundant_patterns_from_block(self.host, self.out_cap, self.stride, self.kmer_clust)
                    else:
                        t1 = time.perf_counter()
                        text, count, flags = self.ctx.cluster_block_patterns_dev(self.merged, self.out_cap, self.stride, self.kmer_clust, stream)
                    if flags & 2:
                        raise api.CrassB200Error(api.EINVAL, "token stride too small for the DR lengths in use")
                    if (flags & 1) or count > self.out_cap:
                        status, text = 1, b""
                    elif len(text) > self.text_cap:
                        status = 2
                except Exception as e:                            # noqa: BLE001 -- reported on every rank below
                    status, text = 3, ("%s: %s" % (type(e).__name__, e)).encode()[: self.text_cap]
                t3 = time.perf_counter()
                self.last_ms = {"gather_merge_d2h": (t1 - t0) * 1e3, "cluster": (t3 - t1) * 1e3}
                hdr.view(np.uint32)[0:2] = (status, min(count, 0xFFFFFFFF))
                hdr.view(np.uint64)[1] = len(text)
                n_send = self.HEADER
                if status in (0, 3) and text:
                    self.msg_host[self.HEADER: self.HEADER + len(text)] = torch.frombuffer(bytearray(text), dtype=torch.uint8)
                    n_send += len(text)
                self.msg[:n_send].copy_(self.msg_host[:n_send], non_blocking=True)
            dist.broadcast(self.msg, src=self.root, group=self.group)
            if self.rank != self.root:
                self.msg_host.copy_(self.msg, non_blocking=True)
                torch.cuda.current_stream(self.dev).synchronize()
            status, count = (int(x) for x in hdr.view(np.uint32)[0:2])
            text_len = int(hdr.view(np.uint64)[1])
            if self.rank == self.root:
                self.last_ms["broadcast"] = (time.perf_counter() - t3) * 1e3
            if status == 3:
                raise api.CrassB200Error(api.EINVAL, "pattern exchange failed on rank %d: %s"
                                         % (self.root, self.msg_host[self.HEADER: self.HEADER + text_len].numpy().tobytes().decode("latin-1")))
            if status == 1:
                self._alloc(self.cap * 2)
                continue
            if status == 2:
                cap = self.text_cap
                while cap < text_len:
                    cap *= 2
                self._alloc_msg(cap)
                continue
            return self.msg_host[self.HEADER: self.HEADER + text_len].numpy().tobytes(), count


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
