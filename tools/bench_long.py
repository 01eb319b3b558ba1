#!/usr/bin/env python
"""Kernel timing on BASELINE.json configs[2]: long reads (1-10 kb) through K1 (+ clustering + K2), device resident.

  python tools/bench_long.py [--reads N] [--steps K]

Not the headline bench (bench.py is); prints one JSON line with the per-kernel CUDA-event times, the HBM-roofline
fractions (algorithmic bytes: sum(L) + 8 B/read + flags + hit records) and a parity check of the found
flags of a prefix against the oracle.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", type=int, default=2000)
    args = ap.parse_args()
    import torch
    import crass_b200 as cb
    from crass_b200 import api, synth

    t0 = time.time()
    bases, offsets, drs = synth.config3(args.reads)
    gen_s = time.time() - t0
    n = args.reads
    max_len = int(np.diff(offsets.astype(np.int64)).max())
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    d_bases = torch.from_numpy(bases).to(dev)
    d_offsets = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    ctx = cb.Context(0)
    ctx.keep_packed(os.environ.get("CRASS_B200_K2F", "") != "bytes")    # K2 reads the 2-bit stream K1 leaves behind
    params = cb.Params()
    hits_cap, pool_cap = n + 1024, 64 * n + 4096
    d_found = torch.empty(n, dtype=torch.uint8, device=dev)
    d_found2 = torch.empty(n, dtype=torch.uint8, device=dev)
    d_hits = torch.empty(hits_cap * 4, dtype=torch.int32, device=dev)
    d_pool = torch.empty(pool_cap, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    d_tok = torch.empty(hits_cap * 64, dtype=torch.uint8, device=dev)
    k1, k2 = [], []
    stats = {}
    for it in range(args.steps + 2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        ctx.set_token_output(d_tok, 64)
        ctx.dr_search_dev(d_bases, d_offsets, n, max_len, params, d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
        ctx.set_token_output(None)
        e[1].record()
        cnt = d_cnt.cpu().numpy()
        nh, npool = int(cnt[0]), int(cnt[1])
        assert not cnt[2], "hit buffers overflowed"
        hits = d_hits[: nh * 4].cpu().numpy().view(api.HIT_DTYPE)
        toks = d_tok[: max(nh, 1) * 64].cpu().numpy()
        local = api.dr_list_from_tokens(toks, 64, hits)
        pats = api.non_redundant_list(local, params.kmer_clust)
        ac = cb.Automaton(pats)
        ctx.ac_upload(ac)
        e[2].record()
        ctx.ac_scan_dev(ac, d_bases, d_offsets, n, max_len, d_found, d_found2, d_hits, d_pool, d_cnt, s.cuda_stream)
        e[3].record()
        torch.cuda.synchronize()
        n2 = int(d_cnt.cpu().numpy()[0])
        if it >= 2:
            k1.append(e[0].elapsed_time(e[1]))
            k2.append(e[2].elapsed_time(e[3]))
        stats = dict(hits_phase1=nh, ss_entries=npool, dr_variants=len(local), patterns=len(pats), hits_phase2=n2)
    # parity of a prefix against the oracle
    import checkers
    P = checkers.port()
    m = min(args.check, n)
    want = np.zeros(m, dtype=np.uint8)
    P.lib.orc_phase1_batch(bases.ctypes.data, offsets.ctypes.data, m, checkers.params_array(), want.ctypes.data)
    got = d_found[:m].cpu().numpy()
    n_bases = int(offsets[-1])
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    b1 = n_bases + 9 * n + 8 * stats["hits_phase1"] + 4 * stats["ss_entries"]
    b2 = n_bases + 10 * n + 16 * stats["hits_phase2"]
    k1m, k2m = float(np.mean(k1)), float(np.mean(k2))
    print(json.dumps({"workload": "config3: %d long reads U[1000,10000] bp (%.2f Gbp), 3%% array bases" % (n, n_bases / 1e9),
                      "max_read_len": max_len, "k1_ms": k1m, "k1_gbp_per_s": n_bases / k1m / 1e6, "k1_frac_of_hbm": b1 / (k1m / 1e3) / 1e9 / peak,
                      "k2_ms": k2m, "k2_gbp_per_s": n_bases / k2m / 1e6, "k2_frac_of_hbm": b2 / (k2m / 1e3) / 1e9 / peak,
                      "stats": stats, "parity_prefix_reads": m, "parity_ok": bool(np.array_equal(got, want)),
                      "oracle_hits_in_prefix": int(want.sum()), "generation_s": gen_s}))
    ctx.close()


if __name__ == "__main__":
    main()
