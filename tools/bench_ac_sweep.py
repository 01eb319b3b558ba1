#!/usr/bin/env python
"""Kernel timing on BASELINE.json configs[4]: the singleton scan (K2) against 100 ... 20 000 patterns, device resident.

  python tools/bench_ac_sweep.py [--reads N] [--patterns 100,300,1000,3000,10000,20000] [--steps K] [--check M]

Workload: N x 150 bp uniform A/C/G/T reads; a pattern set of size P = P/2 random DR-like strings
(length U[23,47]) + their reverse complements; 1 % of the reads carry one planted occurrence; the phase-1 flag is all
zero, so every read is scanned.  For each P prints one JSON line: matcher build time (host), CUDA-event time of K2
(filter + verify + hit records), reads/s, Gbp/s, the fraction of the HBM roofline (algorithmic bytes: L + 8 B offset +
1 B skip flag + 1 B found flag per read + 16 B per hit) and a parity check of the found flags of a prefix of the reads
against the oracle's acism restatement.  Not the headline bench (bench.py is).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

READ_LEN = 150


def make_reads(n, seed, dev):
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    acgt = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    out = torch.empty(n * READ_LEN, dtype=torch.uint8, device=dev)
    chunk = 1 << 28
    for lo in range(0, out.numel(), chunk):
        m = min(chunk, out.numel() - lo)
        out[lo:lo + m] = acgt[torch.randint(0, 4, (m,), generator=g, device=dev)]
    return out


def plant(d_bases, n, patterns, fraction, seed, dev):
    """One occurrence of a random pattern at a random position in `fraction` of the reads (vectorised on the device)."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    P = len(patterns)
    maxlen = max(len(p) for p in patterns)
    mat = np.zeros((P, maxlen), dtype=np.uint8)
    lens = np.zeros(P, dtype=np.int64)
    for i, p in enumerate(patterns):
        mat[i, :len(p)] = np.frombuffer(p, dtype=np.uint8)
        lens[i] = len(p)
    d_mat = torch.from_numpy(mat).to(dev)
    d_len = torch.from_numpy(lens).to(dev)
    pick = torch.nonzero(torch.rand(n, generator=g, device=dev) < fraction).flatten()
    m = pick.numel()
    which = torch.randint(0, P, (m,), generator=g, device=dev)
    plen = d_len[which]
    at = (torch.rand(m, generator=g, device=dev) * (READ_LEN - plen + 1).float()).long().clamp_(min=0)
    at = torch.minimum(at, READ_LEN - plen)
    j = torch.arange(maxlen, device=dev)
    mask = j[None, :] < plen[:, None]
    dst = (pick * READ_LEN + at)[:, None] + j[None, :]
    d_bases[dst[mask]] = d_mat[which][mask]
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=50_000_000)
    ap.add_argument("--patterns", default="100,300,1000,3000,10000,20000")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", type=int, default=200_000)
    args = ap.parse_args()
    import torch
    import crass_b200 as cb
    from crass_b200 import synth
    import checkers

    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    n = args.reads
    peak_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peak_file))["hbm_gbs"] if os.path.exists(peak_file) else 6650.0
    ctx = cb.Context(0)
    d_offsets = torch.arange(n + 1, dtype=torch.int64, device=dev) * READ_LEN
    d_skip = torch.zeros(n, dtype=torch.uint8, device=dev)
    d_found = torch.empty(n, dtype=torch.uint8, device=dev)
    hits_cap, pool_cap = n // 16 + 4096, n // 4 + 4096
    d_hits = torch.empty(hits_cap * 4, dtype=torch.int32, device=dev)
    d_pool = torch.empty(pool_cap, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    O = checkers.port()
    m = min(args.check, n)
    h_offsets = (np.arange(m + 1, dtype=np.uint64) * READ_LEN)
    d_bases = make_reads(n, 20245, dev)
    params = cb.Params()
    d_found1 = torch.empty(n, dtype=torch.uint8, device=dev)
    for P in [int(x) for x in args.patterns.split(",")]:
        patterns = synth.pattern_set(P, seed=20245 + P)
        planted = plant(d_bases, n, patterns, 0.01, 20245 + 7 * P, dev)      # earlier sets stay in the reads as background
        torch.cuda.synchronize()                                           # the planting is asynchronous
        t0 = time.perf_counter()
        ac = cb.Automaton(patterns)
        ctx.ac_upload(ac)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        def timed():
            ts = []
            for it in range(args.steps + 3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ctx.ac_scan_dev(ac, d_bases, d_offsets, n, READ_LEN, d_skip, d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
                b.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(a.elapsed_time(b))
            c = d_cnt.cpu().numpy()
            assert not c[2], "hit buffers overflowed"
            return float(np.mean(ts)), c
        # (1) K2 straight from the bytes (k_ac_filter): phase 2 of a batch whose phase 1 ran elsewhere
        os.environ["CRASS_B200_K2F"] = "bytes"
        ctx.keep_packed(False)
        ms_bytes, cnt = timed()
        nh = int(cnt[0])
        got_bytes = d_found[:m].cpu().numpy()
        # (2) K2 on the 2-bit stream phase 1 of the same batch leaves behind (k_ac_filter_packed): the pipeline's form
        os.environ.pop("CRASS_B200_K2F", None)
        ctx.keep_packed(True)
        ctx.dr_search_dev(d_bases, d_offsets, n, READ_LEN, params, d_found1, d_hits, d_pool, d_cnt, s.cuda_stream)
        ms, cnt = timed()
        assert int(cnt[0]) == nh, "byte and 2-bit forms disagree on the hit count"
        # parity of a prefix against the oracle (acism restatement: first match per read)
        h_bases = d_bases[: m * READ_LEN].cpu().numpy()
        want = np.zeros(m, dtype=np.uint8)
        oh = O.ac_create(patterns)
        O.lib.orc_phase2_batch(oh, h_bases.ctypes.data, h_offsets.ctypes.data, m, want.ctypes.data)
        O.ac_destroy(oh)
        got = d_found[:m].cpu().numpy()
        alg = n * (READ_LEN + 8 + 1 + 1) + 16 * nh
        print(json.dumps({"workload": "config5: %d x %d bp uniform reads, %d patterns (P/2 + revcomps), 1%% planted, every read scanned" % (n, READ_LEN, P),
                          "patterns": P, "reads": n, "planted_reads": int(planted), "hits": nh, "candidates_after_filter": int(cnt[3]),
                          "k2_ms": ms, "reads_per_s": n / ms * 1e3, "gbp_per_s": n * READ_LEN / ms / 1e6,
                          "k2_frac_of_hbm": alg / (ms / 1e3) / 1e9 / peak,
                          "k2_bytes_form_ms": ms_bytes, "k2_bytes_form_frac_of_hbm": alg / (ms_bytes / 1e3) / 1e9 / peak, "matcher_build_upload_ms": build_ms,
                          "parity_prefix_reads": m, "parity_ok": bool(np.array_equal(got, want) and np.array_equal(got_bytes, want)), "oracle_hits_in_prefix": int(want.sum())}), flush=True)
        del ac
    ctx.close()


if __name__ == "__main__":
    main()
