#!/usr/bin/env python
"""The merged DR list an N-rank bench run clusters, produced on ONE GPU: the N shards of bench.py's workload are searched
one after the other, their token blocks are merged by K4c exactly as after the all-gather.

  python tools/merged_dr_list.py --ranks 8 --out gpurun_out/dr_list_8x10M.txt

Input for host-side tuning of the clustering step (crass_b200_ac_build_from_dr_list) at multi-GPU list sizes."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    import torch
    import crass_b200 as cb
    from crass_b200 import api, synth
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    n, TOK, cap = args.reads, 64, 16384
    genome, _, _ = synth.make_genome(20242)
    ctx = cb.Context(0)
    params = cb.Params()
    hits_cap = n // 4 + 1024
    d_found = torch.empty(n, dtype=torch.uint8, device=dev)
    d_hits = torch.empty(hits_cap * 4, dtype=torch.int32, device=dev)
    d_pool = torch.empty(n + 4096, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    d_tok = torch.empty(hits_cap * TOK, dtype=torch.uint8, device=dev)
    nb = api.token_block_bytes(cap, TOK)
    blocks = torch.empty(args.ranks * nb, dtype=torch.uint8, device=dev)
    for rank in range(args.ranks):
        d_bases, d_offsets = synth.sample_fixed_torch(genome, n, 150, 20242 + 1000 + rank, dev)
        d_offsets = d_offsets.to(torch.int64)
        ctx.set_token_output(d_tok, TOK)
        ctx.dr_search_dev(d_bases, d_offsets, n, 150, params, d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
        ctx.set_token_output(None)
        nh = int(d_cnt.cpu()[0])
        ctx.unique_tokens_block_dev(d_hits, nh, d_tok, TOK, blocks[rank * nb:(rank + 1) * nb], cap, s.cuda_stream)
        s.synchronize()
        del d_bases, d_offsets
    out_cap = cap * args.ranks
    merged = torch.empty(api.token_block_bytes(out_cap, TOK), dtype=torch.uint8, device=dev)
    ctx.merge_token_blocks_dev(blocks, args.ranks, cap, TOK, n, merged, out_cap, s.cuda_stream)
    s.synchronize()
    text, count, flags = api.dr_list_from_block(merged.cpu().numpy(), out_cap, TOK)
    assert flags == 0 and count <= out_cap
    with open(args.out, "wb") as fh:
        fh.write(text)
    print("%d ranks -> %d distinct DRs, %d bytes" % (args.ranks, count, len(text)))
    # what the root of an N-rank run does next: K5 + host passes on the merged block (CRASS_B200_TRACE=1 shows the stages)
    import time
    for _ in range(4):
        t0 = time.perf_counter()
        pats, cnt, fl = ctx.cluster_block_patterns_dev(merged, out_cap, TOK, 6, s.cuda_stream)
        dt = time.perf_counter() - t0
    assert pats == api.non_redundant_patterns(text, 6)
    print("clustering of the merged block: %.3f ms, %d patterns" % (dt * 1e3, pats.count(b"\n")))
    ctx.close()


if __name__ == "__main__":
    main()
