#!/usr/bin/env python
"""K1 (direct-repeat search) timing on the bench workload under the kernel-selection knobs of capi.cu.

  python tools/k1_variants.py [--reads N] [--steps K] [--set NAME[,NAME...]]

Knobs (environment, read by crass_b200_dr_search_dev):
  CRASS_B200_K1F=tma        filter on CTA tiles staged by bulk copies (round 1); default: warp tiles, no staged bytes
  CRASS_B200_K1E=lockstep   exact kernel with 32 candidates per warp task (round 1); =refill: finished lanes take new
                            candidates; default: staged (lanes wait for a quorum per stage)
  CRASS_B200_K1_CHUNKS=n    chunks of the batch; with n > 1 the exact kernel of chunk i runs beside the filter of chunk i+1
  CRASS_B200_K1F_CTAS / CRASS_B200_K1E_CTAS   CTAs per SM of the two kernels
Prints one JSON object per variant: CUDA-event milliseconds of the whole K1 call, candidates, hits, and whether the found
flags equal those of the first variant."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

KEYS = ("CRASS_B200_K1F", "CRASS_B200_K1E", "CRASS_B200_K1_CHUNKS", "CRASS_B200_K1F_CTAS", "CRASS_B200_K1E_CTAS", "CRASS_B200_K1_QUORUM", "CRASS_B200_K1_REFILL")

VARIANTS = [
    ("r1: tma filter + lockstep exact", {"CRASS_B200_K1F": "tma", "CRASS_B200_K1E": "lockstep"}),
    ("warp filter + lockstep exact", {"CRASS_B200_K1E": "lockstep"}),
    ("default: warp filter + staged exact 4/SM, quorum 16, refill 8", {}),
    ("staged quorum 8", {"CRASS_B200_K1_QUORUM": "8"}),
    ("staged quorum 12", {"CRASS_B200_K1_QUORUM": "12"}),
    ("staged quorum 20", {"CRASS_B200_K1_QUORUM": "20"}),
    ("staged quorum 24", {"CRASS_B200_K1_QUORUM": "24"}),
    ("staged quorum 32", {"CRASS_B200_K1_QUORUM": "32"}),
    ("staged refill 4", {"CRASS_B200_K1_REFILL": "4"}),
    ("staged refill 16", {"CRASS_B200_K1_REFILL": "16"}),
    ("staged quorum 24 refill 4", {"CRASS_B200_K1_QUORUM": "24", "CRASS_B200_K1_REFILL": "4"}),
    ("staged 3/SM", {"CRASS_B200_K1E_CTAS": "3"}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--keep-packed", type=int, default=1)
    ap.add_argument("--only", default=None, help="substring filter on the variant names")
    args = ap.parse_args()
    import torch
    import crass_b200 as cb
    from crass_b200 import synth
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    n = args.reads
    genome, _, _ = synth.make_genome(20242)
    d_bases, d_offsets = synth.sample_fixed_torch(genome, n, 150, 20242 + 1000, dev)
    d_offsets = d_offsets.to(torch.int64)
    ctx = cb.Context(0)
    ctx.keep_packed(bool(args.keep_packed))
    params = cb.Params()
    hits_cap, pool_cap = n // 4 + 1024, n + 4096
    d_found = torch.empty(n, dtype=torch.uint8, device=dev)
    d_hits = torch.empty(hits_cap * 4, dtype=torch.int32, device=dev)
    d_pool = torch.empty(pool_cap, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    d_tok = torch.empty(hits_cap * 64, dtype=torch.uint8, device=dev)
    first = None
    for name, env in VARIANTS:
        if args.only and args.only not in name:
            continue
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        ts = []
        for it in range(args.steps + 3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.set_token_output(d_tok, 64)
            ctx.dr_search_dev(d_bases, d_offsets, n, 150, params, d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
            ctx.set_token_output(None)
            b.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(a.elapsed_time(b))
        cnt = d_cnt.cpu().numpy()
        found = d_found.clone()
        if first is None:
            first = found
        algo = n * 159 + int(cnt[0]) * 24
        ms = float(np.mean(ts))
        print(json.dumps({"variant": name, "k1_ms": ms, "k1_ms_min": float(np.min(ts)), "frac_of_6544.7": algo / (ms / 1e3) / 1e9 / 6544.7,
                          "candidates": int(cnt[3]), "hits": int(cnt[0]), "same_flags_as_first": bool(torch.equal(found, first))}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
