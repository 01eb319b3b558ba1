"""The N>1 path on CPU: two gloo ranks, each with a contiguous shard, run the DR-set exchange of crass_b200.dist and
must end up with exactly the token order, pattern set and (after replay) dump of a single sequential run.
Device results are stood in for by the oracle (no GPU here); the merge / renumbering / clustering code is the product's."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    import crass_b200 as cb
    from crass_b200 import api, dist as cbdist
    from test_host_logic import oracle_hits_phase1, oracle_hits_phase2

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    whole = cb.Batch.from_file(path)
    n = len(whole)
    lo, hi = rank * n // world, (rank + 1) * n // world
    offs, bases = whole.offsets, whole.bases
    names = [whole.name(i) for i in range(lo, hi)]
    shard = cb.Batch.from_arrays(bases[int(offs[lo]):int(offs[hi])], offs[lo:hi + 1] - offs[lo], names)
    hits, pool = oracle_hits_phase1(shard)
    local = api.dr_list_from_hits(shard.bases, shard.offsets, hits, pool)
    merged = cbdist.allgather_dr_lists(local)                        # the collective under test
    pats = api.non_redundant_list(merged, 6)
    # the same exchange fed with K4b-style records (64-byte token records + first-read index, in device hash order)
    import torch
    order = np.random.default_rng(rank).permutation(len(local))
    rec = np.zeros((len(local) + 3, 64), dtype=np.uint8)
    fr = np.zeros(len(local) + 3, dtype=np.int32)
    for slot, i in enumerate(order):
        d = local[i]
        rec[slot, 0] = len(d)
        rec[slot, 2:2 + len(d)] = np.frombuffer(d, dtype=np.uint8)
        fr[slot] = 10 * i + 7
    blob = cbdist.allgather_unique_tokens(torch.from_numpy(rec.reshape(-1)), torch.from_numpy(fr), len(local), 64)
    assert blob == b"".join(d + b"\n" for d in merged)
    assert cb.Automaton.from_dr_list(blob, 6).num_patterns == len(pats)
    res = cb.Results()
    res.add_phase1(shard, hits, pool)
    res.adopt_tokens(merged)
    skip = np.zeros(len(shard), dtype=np.uint8)
    skip[hits["read_index"]] = 1
    h2, p2 = oracle_hits_phase2(shard, pats, skip)
    with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as fh:
        fh.write("\n".join(d.decode() for d in merged) + "\n--\n" + "\n".join(sorted(p.decode() for p in pats)) + "\n--\n")
        fh.write("%d %d\n" % (len(hits), len(h2)))
        fh.write("\n".join(d.decode() for d in res.dr_list()))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_reproduce_the_sequential_token_order(tmp_path):
    sys.path.insert(0, HERE)
    import checkers
    import crass_b200 as cb
    from crass_b200 import api
    from test_host_logic import oracle_hits_phase1, oracle_hits_phase2

    path = os.path.join(checkers.REF_DATA, "CN_gDC.fa.gz")
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, path, str(tmp_path)), nprocs=2, join=True)
    # sequential run
    whole = cb.Batch.from_file(path)
    hits, pool = oracle_hits_phase1(whole)
    seq_list = api.dr_list_from_hits(whole.bases, whole.offsets, hits, pool)
    seq_pats = sorted(p.decode() for p in api.non_redundant_list(seq_list, 6))
    skip = np.zeros(len(whole), dtype=np.uint8)
    skip[hits["read_index"]] = 1
    h2, _ = oracle_hits_phase2(whole, [p.encode() for p in seq_pats], skip)
    outs = [open(os.path.join(str(tmp_path), "rank%d.txt" % r)).read().split("\n--\n") for r in range(2)]
    n1 = n2 = 0
    for merged, pats, tail in outs:
        assert merged.split("\n") == [d.decode() for d in seq_list]          # identical global token order on every rank
        assert pats.split("\n") == seq_pats                                   # identical pattern set
        counts, *adopted = tail.split("\n")
        assert adopted == [d.decode() for d in seq_list]                      # adopt_tokens == sequential StringCheck
        a, b = map(int, counts.split())
        n1 += a
        n2 += b
    assert n1 == len(hits) and n2 == len(h2)                                  # shards partition the hits of both phases
