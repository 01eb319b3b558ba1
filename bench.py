#!/usr/bin/env python
"""bench.py -- headline benchmark of the read-scanning hot path (BASELINE.json: Mreads/s DR search + singleton scan).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--reads R]

Workload (config.workload = "config2"): BASELINE.json configs[1], synthetic 10M x 150 bp Illumina reads with 50 planted
CRISPR DR types per GPU (weak scaling: every rank scans its own 10M-read shard; configs[3] is the same recipe sharded).
A "step" is one pass of the hot path over the shard:
    K1 direct-repeat search (+K4 tokens, hit ordering) -> K4b distinct-token block [-> NCCL all-gather -> K4c merge on rank 0
    when N>1] -> K5 clustering passes on the GPU + order-dependent passes on the host = createNonRedundantSet [-> NCCL
    broadcast of the pattern set when N>1] -> matcher build + upload -> K2 singleton scan (on the 2-bit stream K1 left in
    HBM) -> both hit lists to pinned host memory in read order.
`value`  : reads/s with the batch already resident in HBM (device timed with CUDA events, max over ranks).
`e2e`    : the same pass through the host-buffer C-ABI (crass_b200_batch_upload / _dr_search_resident / _ac_scan_resident
           + replay into the ReadMap mirror), pinned host input copied H2D and hit records copied D2H inside the timed region.
`roofline`: dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json), algorithmic bytes per SURVEY.md 8(d).
`cpu_baseline`: the reference's own searchFile/findSingletons (oracle/_ref, compiled from the unmodified sources) on a
           bounded prefix of the same reads, 1 thread.
--impl reference: the reference's CPU implementation on all host cores (one process per core over disjoint shards, the only
           parallelism the reference supports) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

READ_LEN = 150
SEED = 20242
METRIC = "reads_per_s_dr_search_plus_singleton_scan"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run_nvml(self):
        """NVML in-process: a sample costs microseconds, so even a 20 ms timed region gets several."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self._halt.is_set():
            self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for b, n in bits.items():
                if r & b:
                    self.reasons.add(n)
            self._halt.wait(0.01)

    def run(self):
        try:
            self.run_nvml()
            return
        except Exception:
            pass                                                       # no NVML binding: poll nvidia-smi instead
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def write_fasta(path, bases, offsets, lo, hi):
    L = int(offsets[1] - offsets[0])
    blk = bases[int(offsets[lo]):int(offsets[hi])].reshape(hi - lo, L)
    hdr = np.frombuffer(b"".join(b">r%010d\n" % i for i in range(lo, hi)), dtype=np.uint8).reshape(hi - lo, 13)
    rec = np.concatenate([hdr, blk, np.full((hi - lo, 1), 10, dtype=np.uint8)], axis=1)
    rec.tofile(path)


def ref_worker(args):
    """One reference process over one FASTA shard; returns (reads, seconds phase1, cluster, phase2)."""
    path, n = args[:2]
    import checkers
    R = checkers.ref()
    t = time.time()
    dump, tm = R.run_files([path])
    return n, tm[0] / 1e3, tm[1] / 1e3, tm[2] / 1e3, time.time() - t, (dump if len(args) > 2 and args[2] else None)


def cpu_reference_rate(bases, offsets, n_sample, n_procs, tmpdir, keep_dump=False):
    """reads/s of the reference's own searchFile + findSingletons on n_sample reads split over n_procs processes.
    keep_dump (one process only): also return the reference's result dump and the FASTA path for the parity check."""
    import multiprocessing as mp
    import checkers
    kind = "reference" if checkers.have_ref() else "port"
    per = n_sample // n_procs
    jobs = []
    for p in range(n_procs):
        path = os.path.join(tmpdir, "shard%d.fa" % p)
        write_fasta(path, bases, offsets, p * per, (p + 1) * per)
        jobs.append((path, per))
    t0 = time.time()
    if kind == "reference":
        if n_procs == 1:
            res = [ref_worker(jobs[0] + (keep_dump,))]
        else:
            with mp.get_context("spawn").Pool(n_procs) as pool:
                res = pool.map(ref_worker, jobs)
    else:
        P = checkers.port()
        res = []
        for path, n in jobs:
            _, tm = P.run_files([path])
            res.append((n, tm[0] / 1e3, tm[1] / 1e3, tm[2] / 1e3, 0.0, None))
    wall = time.time() - t0
    worst = max(r[1] + r[2] + r[3] for r in res)
    return dict(kind=kind, reads=per * n_procs, seconds=worst, wall=wall, rate=per * n_procs / worst,
                phase1_s=max(r[1] for r in res), phase2_s=max(r[3] for r in res),
                dump=res[0][5] if keep_dump else None, path=jobs[0][0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU (weak scaling)")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-dr-list", default=None, help="write the merged DR list of the last step here (host-side tuning input)")
    args = ap.parse_args()

    # stdout carries exactly ONE JSON line: everything libraries print while the run is going on (e.g. NCCL's version
    # banner) is sent to stderr, the line is written to the saved descriptor at the very end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "config2: synthetic %d x %dbp Illumina reads per GPU, 50 planted CRISPR DR types, 1%% array bases, 0.1%% subs, 0.05%% N, default crass options"
                          % (args.reads, READ_LEN),
              "reads_per_gpu": args.reads, "read_len": READ_LEN, "seed": SEED, "sharding": "contiguous read ranges, one rank per GPU",
              "l2": "inputs (%.2f GB per GPU) are larger than L2 (126 MB); no flush needed" % (args.reads * (READ_LEN + 8) / 1e9)}

    from crass_b200 import synth
    genome, drs, _ = synth.make_genome(SEED)

    # ------------------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        n_procs = os.cpu_count() or 1
        n_sample = max(n_procs, min(args.reads, 4_000_000, 250_000 * n_procs))
        n_sample -= n_sample % n_procs
        bases, offsets = synth.sample_fixed(genome, n_sample, READ_LEN, SEED + 1000)
        vals = []
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
            for it in range(args.warmup + args.steps):
                r = cpu_reference_rate(bases, offsets, n_sample, n_procs, d)
                if it >= args.warmup:
                    vals.append(r)
        rate = float(np.mean([v["rate"] for v in vals]))
        ms = float(np.mean([v["seconds"] for v in vals])) * 1e3
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": dict(config, sample_reads=n_sample),
                "cpu_baseline": {"value": rate, "unit": "reads/s", "cores": n_procs, "kind": vals[0]["kind"],
                                 "sample": "%d-read prefix of the config2 recipe split over %d processes (searchFile + createNonRedundantSet + findSingletons each, FASTA in tmpfs)" % (n_sample, n_procs)},
                "e2e": {"value": rate, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------------------------------------ own arm
    import torch
    import torch.distributed as dist
    import crass_b200 as cb
    from crass_b200 import api
    from crass_b200 import dist as cbdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.reads
    d_bases, d_offsets = synth.sample_fixed_torch(genome, n, READ_LEN, SEED + 1000 + rank, dev)
    d_offsets = d_offsets.to(torch.int64)
    h_bases = torch.empty(d_bases.shape, dtype=torch.uint8, pin_memory=True)
    h_offsets = torch.empty(d_offsets.shape, dtype=torch.int64, pin_memory=True)
    h_bases.copy_(d_bases)
    h_offsets.copy_(d_offsets)
    torch.cuda.synchronize()
    np_bases, np_offsets = h_bases.numpy(), h_offsets.numpy().view(np.uint64)

    # The host passes between the kernels use helper threads.  At N > 1 the clustering runs once, on rank 0
    # (crass_b200/dist.py::PatternExchange), so rank 0 gets the cores the other ranks do not need.
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    cores = os.cpu_count() or 8
    if world == 1:
        host_threads = min(8, cores)
    else:
        host_threads = max(2, min(16, cores - 2 * local_world)) if rank == 0 else 2
    threads_given = "CRASS_B200_HOST_THREADS" in os.environ
    os.environ.setdefault("CRASS_B200_HOST_THREADS", str(host_threads))
    ctx = cb.Context(local_rank)
    ctx.keep_packed(True)                                # K2 reads the 2-bit stream K1's filter leaves in HBM (same, unchanged batch)
    params = cb.Params()
    work_stream = torch.cuda.Stream(device=dev)          # a real (non-NULL) stream: kernels, copies and events all go here
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    assert stream != 0
    hits_cap, pool_cap = n // 4 + 1024, n + 4096
    d_found = torch.empty(n, dtype=torch.uint8, device=dev)
    d_found2 = torch.empty(n, dtype=torch.uint8, device=dev)
    d_hits = torch.empty(hits_cap * 4, dtype=torch.int32, device=dev)
    d_sorted = torch.empty(hits_cap * 4, dtype=torch.int32, device=dev)      # the hit records in read order
    d_pool = torch.empty(pool_cap, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    h_cnt = torch.zeros(8, dtype=torch.int32, pin_memory=True)
    kt = {"k1": [], "k2": []}
    stats = {}

    TOK = 64                                                           # bytes per K4 token record (>= high_dr + 2)
    d_tokens = torch.empty(hits_cap * TOK, dtype=torch.uint8, device=dev)
    # host time line of a step: wait for K1 | token exchange + clustering (N == 1: + matcher build) | matcher build from the
    # broadcast pattern set (N > 1) | matcher upload | wait for K2 + ordering | last copy
    HOST_KEYS = ("k1_wait", "exchange_cluster", "matcher_build", "ac_upload", "k2_wait", "fetch_hits2")
    host_ms = {k: [] for k in HOST_KEYS}

    if world == 1:
        exchange = cbdist.TokenExchange(ctx, dev, shard_reads=n, stride=TOK)                     # K4b -> DR list
    else:
        exchange = cbdist.PatternExchange(ctx, dev, shard_reads=n, kmer_clust=params.kmer_clust, stride=TOK)   # + all-gather, K4c, broadcast
    h_hits = [torch.empty(hits_cap * 4, dtype=torch.int32, pin_memory=True) for _ in range(2)]
    h_pool = [torch.empty(pool_cap, dtype=torch.int32, pin_memory=True) for _ in range(2)]

    def read_counters():
        h_cnt.copy_(d_cnt, non_blocking=False)                         # 32 bytes, synchronises the stream
        nh, npool, ovf = int(h_cnt[0]), int(h_cnt[1]), int(h_cnt[2])
        assert not ovf, "bench hit buffers overflowed"
        return nh, npool

    def fetch_hits_async(which, nh, npool):
        """hit records -> pinned host memory, asynchronously on the work stream (device order; consumers sort by read index)"""
        h_hits[which][: nh * 4].copy_(d_sorted[: nh * 4], non_blocking=True)
        h_pool[which][: max(npool, 1)].copy_(d_pool[: max(npool, 1)], non_blocking=True)

    def host_hits(which, nh, npool):
        return h_hits[which][: nh * 4].numpy().view(api.HIT_DTYPE), h_pool[which][: max(npool, 1)].numpy().view(np.uint32)

    def merge_dr_lists(local):
        # one NCCL all-gather of the per-shard DR sets + deterministic merge (crass_b200/dist.py); identity at N=1
        return cbdist.allgather_dr_lists(local, device=dev)

    def step_resident(record):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        ctx.set_token_output(d_tokens, TOK)                            # K4: DR tokens are extracted where the hits are found
        ctx.dr_search_dev(d_bases, d_offsets, n, READ_LEN, params, d_found, d_hits, d_pool, d_cnt, stream)
        ctx.set_token_output(None)
        e[1].record()
        ctx.sort_hits_dev(d_found, n, d_hits, d_cnt, hits_cap, d_sorted, stream)       # read order (what replay consumes)
        t0 = time.perf_counter()
        nh, npool = read_counters()
        t1 = time.perf_counter()
        # distinct low-lexi DRs of all shards in first-appearance order (crass_b200/dist.py): K4b de-duplicates this
        # shard's tokens on the device, one NCCL all-gather + K4c merge the shards, one copy brings the list back
        fetch_hits_async(0, nh, npool)                                 # the phase-1 hit records travel while the host clusters
        if world == 1:
            # K4b block -> host -> createNonRedundantSet + matcher, straight from the block
            if args.dump_dr_list and not record:
                open(args.dump_dr_list, "wb").write(exchange.run(d_hits, nh, d_tokens, stream)[0])
            ac, nu = exchange.run_matcher(d_hits, nh, d_tokens, params.kmer_clust, stream)
            t3 = t4 = time.perf_counter()
        else:
            pat_text, nu = exchange.run(d_hits, nh, d_tokens, stream)  # the pattern set, clustered once on rank 0
            t3 = time.perf_counter()
            ac = cb.Automaton.from_pattern_text(pat_text) if pat_text else None
            t4 = time.perf_counter()
        pats = ac.num_patterns if ac else 0
        n2 = 0
        if pats:
            ctx.ac_upload(ac)
            t5 = time.perf_counter()
            e[2].record()
            ctx.ac_scan_dev(ac, d_bases, d_offsets, n, READ_LEN, d_found, d_found2, d_hits, d_pool, d_cnt, stream)
            e[3].record()
            ctx.sort_hits_dev(d_found2, n, d_hits, d_cnt, hits_cap, d_sorted, stream)
            n2, npool2 = read_counters()
            t6 = time.perf_counter()
            fetch_hits_async(1, n2, npool2)
        torch.cuda.synchronize()                                       # both hit lists are on the host now, in read order
        hits, pool = host_hits(0, nh, npool)
        t7 = time.perf_counter()
        if record:
            kt["k1"].append(e[0].elapsed_time(e[1]))
            if pats:
                kt["k2"].append(e[2].elapsed_time(e[3]))
                for k, v in zip(HOST_KEYS,
                                (t1 - t0, t3 - t1, t4 - t3, t5 - t4, t6 - t5, t7 - t6)):
                    host_ms[k].append(v * 1e3)
                for k, v in getattr(exchange, "last_ms", {}).items():                   # inside exchange_cluster (rank 0's view)
                    host_ms.setdefault("exchange:" + k, []).append(v)
        stats.update(hits_phase1=len(hits), dr_variants_merged=nu, patterns=pats, hits_phase2=n2)

    def step_e2e():
        ctx.upload(h_bases, h_offsets)                                 # H2D from pinned host memory
        hits, pool, _ = ctx.dr_search_resident(params)
        merged = merge_dr_lists(ctx.last_dr_list())
        nb = hits.nbytes + pool.nbytes
        if merged:
            ac = cb.Automaton.from_dr_list(merged, params.kmer_clust)
            hits2, pool2, _ = ctx.ac_scan_resident(ac, skip_found=True)
            nb += hits2.nbytes + pool2.nbytes
        return nb

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident(False)
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None         # one nvidia-smi poller per job, not per rank
    if sampler:
        sampler.start()
    ms_total = timed(lambda: step_resident(True), args.steps)
    launches = ctx.launch_count - launches0
    # in the host-buffer path every rank clusters the merged list itself: share the cores evenly again
    if not threads_given:
        os.environ["CRASS_B200_HOST_THREADS"] = str(max(1, min(8, cores // max(local_world, 1))))
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    d2h = [0]
    ms_e2e = timed(lambda: d2h.__setitem__(0, step_e2e()), args.steps)
    clocks = sampler.stop() if sampler else None                       # sampled across both timed regions

    if rank == 0:
        total_reads = n * world
        ms_step = ms_total / args.steps
        value = total_reads / (ms_step / 1e3)
        e2e_value = total_reads / (ms_e2e / args.steps / 1e3)
        peak, peak_src = peaks()
        k1 = float(np.mean(kt["k1"]))
        k2 = float(np.mean(kt["k2"])) if kt["k2"] else 0.0
        n_bases = n * READ_LEN
        bytes_k1 = n_bases + 8 * n + n + stats["hits_phase1"] * 24            # L + 8 B/read in, 1 B flag, 8+8n B per hit
        bytes_k2 = n_bases + 8 * n + n + n + stats["hits_phase2"] * 16
        dom, dom_ms, dom_bytes = ("K1 dr_search", k1, bytes_k1) if k1 >= k2 else ("K2 singleton_scan", k2, bytes_k2)
        achieved = dom_bytes / (dom_ms / 1e3) / 1e9
        traffic = None                                                         # measured DRAM bytes per launch, from the committed ncu capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["config2_%dx%d" % (n, READ_LEN)][dom]["bytes"]
        except (OSError, KeyError, ValueError):
            pass
        line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config,
                "gbp_per_s": value * READ_LEN / 1e9,
                "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h_bases.numel() + h_offsets.numel() * 8),
                        "d2h_bytes_per_step": int(d2h[0]), "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                             "launches": (["k_dr_filter", "k_dr_exact_packed"] if dom.startswith("K1") else ["k_ac_filter_packed", "k_ac_verify_mask"]),
                             "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": int(dom_bytes),
                             "kernel_ms": dom_ms},
                "kernels": {"k1_dr_search_ms": k1, "k1_frac_of_hbm": bytes_k1 / (k1 / 1e3) / 1e9 / peak,
                            "k2_singleton_scan_ms": k2, "k2_frac_of_hbm": (bytes_k2 / (k2 / 1e3) / 1e9 / peak) if k2 else None,
                            "k1_plus_k2_frac_of_hbm": ((bytes_k1 + bytes_k2) / ((k1 + k2) / 1e3) / 1e9 / peak) if k2 else None,
                            "host_between_kernels_ms": ms_step - k1 - k2,
                            "host_breakdown_ms": {k: float(np.mean(v)) for k, v in host_ms.items() if v}},
                "stats": stats, "clocks": clocks}
        if not args.no_cpu_baseline:
            ns = min(args.cpu_sample, n)
            with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
                r = cpu_reference_rate(np_bases, np_offsets, ns, 1, d, keep_dump=True)
                if r["dump"] is not None:
                    # SURVEY 8d "parity check accompanying every timing": the product's whole path (parser -> K1 -> clustering
                    # -> K2 -> replay) on the very FASTA the reference just processed; the two result dumps (tokens in
                    # numbering order, DRs, reads, orientation, start/stops, patterns) must be the same bytes.
                    import hashlib
                    t0 = time.time()
                    res, max_len = ctx.run_files([r["path"]])
                    mine = res.dump(max_len)
                    line["parity"] = {"sample": "the cpu_baseline sample (%d reads), whole path through the C-ABI vs the reference" % ns,
                                      "dump_identical": bool(mine == r["dump"]), "dump_bytes": len(r["dump"]),
                                      "dump_md5": hashlib.md5(r["dump"].encode("latin-1")).hexdigest(),
                                      "found_reads": int(res.num_reads), "tokens": int(res.num_tokens), "b200_seconds": time.time() - t0}
            line["cpu_baseline"] = {"value": r["rate"], "unit": "reads/s", "cores": 1, "kind": r["kind"],
                                    "sample": "first %d reads of rank 0's shard as FASTA in tmpfs: searchFile %.2fs + findSingletons %.2fs, 1 thread" % (ns, r["phase1_s"], r["phase2_s"])}
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
