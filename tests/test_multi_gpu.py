"""The N>1 exchange on real GPUs (needs two of them; skipped otherwise): two NCCL ranks with contiguous shards must end up
with exactly the pattern set of one sequential run, through both forms of the exchange (every rank merges and clusters /
the root clusters and broadcasts)."""
import os
import random
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _reads():
    import fuzzgen
    rng = random.Random(211)
    pool = [fuzzgen.rand_seq(rng, rng.randint(24, 40)) for _ in range(10)]
    return [fuzzgen.planted_read(rng, rng.choice([100, 150, 150, 250]), dr=rng.choice(pool), sub_rate=rng.choice([0, 0.01])) for _ in range(8000)]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    import crass_b200 as cb
    from crass_b200 import api, dist as cbdist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
    reads = _reads()
    cuts = [0, 3100, len(reads)]                                        # unequal shards
    lo, hi = cuts[rank], cuts[rank + 1]
    bases, offs = cb.pack_reads(reads[lo:hi])
    n = hi - lo
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    ctx = cb.Context(rank)
    TOK = 64
    d_b = torch.from_numpy(bases).to(dev)
    d_o = torch.from_numpy(offs.astype(np.int64)).to(dev)
    d_found = torch.empty(n, dtype=torch.uint8, device=dev)
    d_hits = torch.empty((n + 16) * 4, dtype=torch.int32, device=dev)
    d_pool = torch.empty(64 * n, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    d_tok = torch.empty((n + 16) * TOK, dtype=torch.uint8, device=dev)
    ctx.set_token_output(d_tok, TOK)
    ctx.dr_search_dev(d_b, d_o, n, 256, cb.Params(), d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
    ctx.set_token_output(None)
    nh = int(d_cnt.cpu()[0])
    shard_reads = max(b - a for a, b in zip(cuts, cuts[1:]))
    # small blocks / message on purpose: both growth paths are taken
    every = cbdist.TokenExchange(ctx, dev, shard_reads, stride=TOK, cap=16)
    drs, count = every.run(d_hits, nh, d_tok, s.cuda_stream)
    root = cbdist.PatternExchange(ctx, dev, shard_reads, kmer_clust=6, stride=TOK, cap=16, text_cap=64)
    pats, count2 = root.run(d_hits, nh, d_tok, s.cuda_stream)
    assert count == count2 == drs.count(b"\n")
    with open(os.path.join(out_dir, "rank%d.bin" % rank), "wb") as fh:
        fh.write(drs + b"--\n" + pats + b"--\n" + api.non_redundant_patterns(drs, 6))
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpus_reproduce_the_sequential_pattern_set(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import crass_b200 as cb
    from crass_b200 import api
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    bases, offs = cb.pack_reads(_reads())
    ctx = cb.Context(0)
    ctx.upload(bases, offs)
    hits, pool, _ = ctx.dr_search_resident(cb.Params())
    want_list = b"".join(d + b"\n" for d in api.dr_list_from_hits(bases, offs, hits, pool))
    want_pats = api.non_redundant_patterns(want_list, 6)
    assert want_list.count(b"\n") > 100
    for r in range(2):
        drs, pats, pats_local = open(os.path.join(str(tmp_path), "rank%d.bin" % r), "rb").read().split(b"--\n")
        assert drs == want_list                                            # token order of one sequential run, on every rank
        assert pats == want_pats and pats_local == want_pats               # the broadcast set == clustering the list locally
