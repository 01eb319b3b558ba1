#!/usr/bin/env python
"""Kernel timing of K6 (partial-DR recovery = ReadHolder::updateStartStops + smithWaterman) on the
bench workload: the phase-1 hits of config 2 (synthetic 150 bp reads with planted arrays), every hit against the DR it
carries (front offset 0), device resident.

  python tools/bench_uss.py [--reads N] [--steps K] [--check M]

Prints one JSON line: CUDA-event time of k_update_start_stops, found reads/s, alignment cells/s, the share of jobs that
gained a partial repeat, a parity check of the first M jobs against the oracle and the oracle's single-thread time on
them (the reference's cost for the same work).  Not the headline bench (bench.py is).
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
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", type=int, default=3000)
    args = ap.parse_args()
    import torch
    import crass_b200 as cb
    from crass_b200 import api, synth
    import checkers

    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    n, L = args.reads, 150
    genome, _, _ = synth.make_genome(20242)
    d_bases, d_offsets = synth.sample_fixed_torch(genome, n, L, 20242 + 1000, dev)
    ctx = cb.Context(0)
    params = cb.Params()
    hits_cap, pool_cap = n // 4 + 1024, n + 4096
    d_found = torch.empty(n, dtype=torch.uint8, device=dev)
    d_hits = torch.empty(hits_cap * 4, dtype=torch.int32, device=dev)
    d_pool = torch.empty(pool_cap, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    ctx.dr_search_dev(d_bases, d_offsets, n, L, params, d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
    cnt = d_cnt.cpu().numpy()
    nh, npool = int(cnt[0]), int(cnt[1])
    hits = d_hits[: nh * 4].cpu().numpy().view(api.HIT_DTYPE)
    hits = hits[np.argsort(hits["read_index"], kind="stable")]
    pool = d_pool[:npool].cpu().numpy().astype(np.uint32)
    # one job per hit; its DR = the repeat DRLowLexi would pick (the second one when there are two or more)
    k = np.where(hits["n_ss"] >= 4, 2, 0).astype(np.int64)
    st = pool[hits["ss_offset"].astype(np.int64) + k].astype(np.int64)
    en = pool[hits["ss_offset"].astype(np.int64) + k + 1].astype(np.int64)
    dlen = np.minimum(en - st + 1, 127)
    dr_offs = np.zeros(nh + 1, dtype=np.uint32)
    dr_offs[1:] = np.cumsum(dlen)
    src = np.repeat(hits["read_index"].astype(np.int64) * L + st, dlen) + (np.arange(int(dr_offs[-1])) - np.repeat(dr_offs[:-1].astype(np.int64), dlen))
    d_dr_bytes = d_bases[torch.from_numpy(src).to(dev)].contiguous()
    jobs = np.zeros(nh, dtype=api.USS_JOB_DTYPE)
    jobs["read"] = hits["read_index"]
    jobs["ss_offset"] = hits["ss_offset"]
    jobs["n_ss"] = hits["n_ss"]
    jobs["front_offset"] = 0
    jobs["dr"] = np.arange(nh)
    out_off = np.zeros(nh + 1, dtype=np.int64)
    out_off[1:] = np.cumsum(hits["n_ss"].astype(np.int64) + 4)
    jobs["out_offset"] = out_off[:-1]
    d_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).to(dev)
    d_dr_offs = torch.from_numpy(dr_offs.view(np.int32).copy()).to(dev)
    d_out = torch.zeros(int(out_off[-1]), dtype=torch.int32, device=dev)
    d_n = torch.zeros(nh, dtype=torch.int32, device=dev)
    d_st = torch.zeros(nh, dtype=torch.uint8, device=dev)
    ts = []
    for it in range(args.steps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.update_start_stops_dev(d_bases, d_offsets, d_dr_bytes, d_dr_offs, d_jobs, nh, d_pool, params.low_spacer, d_out, d_n, d_st, s.cuda_stream)
        b.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts))
    n_out = d_n.cpu().numpy()
    status = d_st.cpu().numpy()
    out = d_out.cpu().numpy().view(np.uint32)
    # alignment cells: (first start - lowSpacer) x DR in front, (L - last end - lowSpacer) x DR behind, on the shifted lists
    first = pool[hits["ss_offset"].astype(np.int64)].astype(np.int64)
    last = np.minimum(pool[hits["ss_offset"].astype(np.int64) + hits["n_ss"].astype(np.int64) - 2].astype(np.int64) + dlen - 1, L - 1)
    cells = (np.maximum(first - params.low_spacer, 0) + np.maximum(L - last - params.low_spacer, 0)) * dlen
    # parity + the oracle's time on a prefix of the jobs
    P = checkers.port()
    m = min(args.check, nh)
    h_bases = d_bases.cpu().numpy()
    h_dr = d_dr_bytes.cpu().numpy()
    ok = True
    t_cpu = 0.0
    for i in range(m):
        r = int(hits["read_index"][i])
        seq = h_bases[r * L:(r + 1) * L].tobytes()
        ss = pool[int(hits["ss_offset"][i]): int(hits["ss_offset"][i]) + int(hits["n_ss"][i])].tolist()
        dr = h_dr[int(dr_offs[i]): int(dr_offs[i + 1])].tobytes()
        t0 = time.perf_counter()
        want = P.update_start_stops(seq, ss, 0, dr, params.low_spacer)
        t_cpu += time.perf_counter() - t0
        got = (int(status[i]), out[int(out_off[i]): int(out_off[i]) + int(n_out[i])].tolist())
        ok &= got == ((3, []) if want[0] == -3 else (0, want[1]))
    print(json.dumps({"workload": "K6 on the %d phase-1 hits of config 2 (%d x %d bp), DR = the hit's own repeat, front offset 0" % (nh, n, L),
                      "jobs": nh, "k6_ms": ms, "found_reads_per_s": nh / ms * 1e3, "sw_cells": int(cells.sum()), "sw_gcells_per_s": float(cells.sum()) / ms / 1e6,
                      "jobs_with_partial_repeat": int((n_out > hits["n_ss"]).sum()), "status_nonzero": int((status != 0).sum()),
                      "parity_jobs": m, "parity_ok": bool(ok), "oracle_1_thread_jobs_per_s": m / t_cpu if t_cpu else None,
                      "speedup_vs_oracle_1_thread": (nh / ms * 1e3) / (m / t_cpu) if t_cpu else None}))
    ctx.close()


if __name__ == "__main__":
    main()
