#!/usr/bin/env python
"""K1 on a batch of mostly short reads with a few long ones (the length cliff): 10 M x 150 bp + N reads of 305..2000 bp.

  python tools/bench_mixed.py [--reads 10000000] [--long 1000]

Prints one JSON line per mode: CRASS_B200_K1_MIXED=1 (default: short path + listed long reads) and =0 (round 1: the whole
batch on the warp-per-read kernel), with the hit counts (which must agree)."""
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
    ap.add_argument("--long", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import crass_b200 as cb
    from crass_b200 import synth
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    genome, _, _ = synth.make_genome(20242)
    n = args.reads
    bases, offs = synth.sample_fixed(genome, n, 150, 20243)
    rng = np.random.default_rng(5)
    lens = np.full(n + args.long, 150, dtype=np.int64)
    where = np.sort(rng.choice(n + args.long, args.long, replace=False))
    lens[where] = rng.integers(305, 2000, args.long)
    offsets = np.zeros(n + args.long + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    g = np.frombuffer(genome, dtype=np.uint8) if isinstance(genome, (bytes, bytearray)) else np.asarray(genome)
    out = np.empty(total, dtype=np.uint8)
    short_mask = np.ones(n + args.long, dtype=bool)
    short_mask[where] = False
    short_starts = offsets[:-1][short_mask]
    idx = (short_starts[:, None] + np.arange(150)[None, :]).reshape(-1)
    out[idx] = bases
    for w in where:
        L = int(lens[w]); p = int(rng.integers(0, len(g) - L))
        out[offsets[w]:offsets[w] + L] = g[p:p + L]
    d_bases = torch.from_numpy(out).to(dev)
    d_offsets = torch.from_numpy(offsets).to(dev)
    m = n + args.long
    ctx = cb.Context(0)
    ctx.keep_packed(True)
    d_found = torch.empty(m, dtype=torch.uint8, device=dev)
    d_hits = torch.empty((m // 4 + 1024) * 4, dtype=torch.int32, device=dev)
    d_pool = torch.empty(m + 4096, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    max_len = int(lens.max())
    for mode in ("1", "0"):
        os.environ["CRASS_B200_K1_MIXED"] = mode
        ts = []
        for it in range(args.steps + 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.dr_search_dev(d_bases, d_offsets, m, max_len, cb.Params(), d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b))
        c = d_cnt.cpu().numpy()
        print(json.dumps({"mode": "mixed" if mode == "1" else "all reads on the warp-per-read kernel", "reads": m, "long_reads": args.long, "max_read_len": max_len,
                          "k1_ms": float(np.mean(ts)), "hits": int(c[0]), "found_flags": int(d_found.sum().item())}), flush=True)
    os.environ.pop("CRASS_B200_K1_MIXED", None)
    ctx.close()


if __name__ == "__main__":
    main()
