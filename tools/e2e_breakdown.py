#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its time (host-buffer C-ABI path), call by call.

  python tools/e2e_breakdown.py [--reads N]

Prints one JSON line with the mean milliseconds of each call of the e2e step over a few repetitions."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    import torch
    import crass_b200 as cb
    from crass_b200 import synth
    dev = torch.device("cuda", 0)
    genome, _, _ = synth.make_genome(20242)
    d_bases, d_offsets = synth.sample_fixed_torch(genome, args.reads, 150, 20242 + 1000, dev)
    h_bases = torch.empty(d_bases.shape, dtype=torch.uint8, pin_memory=True)
    h_offsets = torch.empty(d_offsets.shape, dtype=torch.int64, pin_memory=True)
    h_bases.copy_(d_bases)
    h_offsets.copy_(d_offsets.to(torch.int64))
    torch.cuda.synchronize()
    del d_bases, d_offsets
    ctx = cb.Context(0)
    params = cb.Params()
    acc = {}

    def lap(name, t0):
        t = time.perf_counter()
        acc.setdefault(name, []).append((t - t0) * 1e3)
        return t

    for rep in range(args.reps + 1):
        if rep == 1:
            acc.clear()
        t = time.perf_counter()
        ctx.upload(h_bases, h_offsets)
        t = lap("upload", t)
        hits, pool, _ = ctx.dr_search_resident(params)
        t = lap("dr_search_resident", t)
        merged = ctx.last_dr_list()
        t = lap("last_dr_list", t)
        ac = cb.Automaton.from_dr_list(merged, params.kmer_clust)
        t = lap("cluster+build", t)
        hits2, pool2, _ = ctx.ac_scan_resident(ac, skip_found=True)
        t = lap("ac_scan_resident", t)
    print(json.dumps({k: float(np.mean(v)) for k, v in acc.items()}))
    ctx.close()


if __name__ == "__main__":
    main()
