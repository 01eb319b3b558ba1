#!/usr/bin/env python
"""K2 (singleton scan) timing on the bench workload under the kernel-selection knobs of capi.cu.

  python tools/k2_variants.py [--reads N]

CRASS_B200_K2V=list picks the thread-per-candidate verify kernel (default: one warp per candidate).
Prints one JSON object per variant: CUDA-event milliseconds, candidates after the filter, hits."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import crass_b200 as cb
    from crass_b200 import api, synth, dist as cbdist
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    n = args.reads
    genome, _, _ = synth.make_genome(20242)
    d_bases, d_offsets = synth.sample_fixed_torch(genome, n, 150, 20242 + 1000, dev)
    d_offsets = d_offsets.to(torch.int64)
    ctx = cb.Context(0)
    params = cb.Params()
    hits_cap, pool_cap = n // 4 + 1024, n + 4096
    d_found = torch.empty(n, dtype=torch.uint8, device=dev)
    d_found2 = torch.empty(n, dtype=torch.uint8, device=dev)
    d_hits = torch.empty(hits_cap * 4, dtype=torch.int32, device=dev)
    d_pool = torch.empty(pool_cap, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    d_tok = torch.empty(hits_cap * 64, dtype=torch.uint8, device=dev)
    ex = cbdist.TokenExchange(ctx, dev, shard_reads=n, stride=64)
    ctx.set_token_output(d_tok, 64)
    ctx.dr_search_dev(d_bases, d_offsets, n, 150, params, d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
    ctx.set_token_output(None)
    nh = int(d_cnt.cpu()[0])
    merged, nu = ex.run(d_hits, nh, d_tok, s.cuda_stream)
    ac = cb.Automaton.from_dr_list(merged, params.kmer_clust)
    ctx.ac_upload(ac)
    for name, env in (("thread-per-candidate verify", {"CRASS_B200_K2V": "list"}), ("warp-per-candidate verify", {})):
        for k in ("CRASS_B200_K2V", "CRASS_B200_K2F"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ts = []
        for it in range(args.steps + 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.ac_scan_dev(ac, d_bases, d_offsets, n, 150, d_found, d_found2, d_hits, d_pool, d_cnt, s.cuda_stream)
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b))
        cnt = d_cnt.cpu().numpy()
        print(json.dumps({"variant": name, "k2_ms": float(np.mean(ts)), "candidates": int(cnt[3]), "hits": int(cnt[0]), "patterns": ac.num_patterns}))
    ctx.close()


if __name__ == "__main__":
    main()
