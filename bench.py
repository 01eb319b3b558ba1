#!/usr/bin/env python
"""bench.py -- benchmark of the read-scanning hot path (BASELINE.json: Mreads/s DR search + singleton scan).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|3|4|5] [--no-extra]

Headline workload (config.workload, default "config2"): BASELINE.json configs[1], synthetic 10M x 150 bp Illumina reads
with 50 planted CRISPR DR types per GPU (weak scaling: every rank scans its own 10M-read shard).  A "step" is one pass of
the hot path over the shard:
    K1 direct-repeat search (+K4 tokens, hit ordering) -> K4b distinct-token block [-> NCCL all-gather -> K4c merge on rank 0
    when N>1] -> K5 clustering passes on the GPU + order-dependent passes on the host = createNonRedundantSet [-> NCCL
    broadcast of the pattern set when N>1] -> matcher build + upload -> K2 singleton scan (on the 2-bit stream K1 left in
    HBM) -> both hit lists to pinned host memory in read order.
`value`   : reads/s with the batch already resident in HBM (device timed with CUDA events, max over ranks).
`e2e`     : FILE PATH IN -> CONTAINERS OUT through the C-ABI (crass_b200_engine_run_files on the shard's FASTA in tmpfs):
            kseq-compatible parse on worker threads -> page-locked memory -> H2D -> K1 -> token exchange + clustering ->
            K2 -> D2H -> replay into the ReadMap / StringCheck mirror.  The same work as the reference arm.
`e2e_hostbuf`: the pass through the host-buffer C-ABI alone (pinned arrays in, hit records out; no parsing, no replay).
`roofline`: dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json), algorithmic bytes per SURVEY.md 8(d).
`cpu_baseline`: the reference's own searchFile/findSingletons (oracle/_ref, compiled from the unmodified sources) on a
            bounded prefix of the same reads, 1 thread.
`parity`  : N=1: the e2e path's result dump on the cpu_baseline sample must equal the reference's byte for byte.
            N>1: `merged_identical` -- a >= 1M-read prefix cut into N contiguous shards goes through the N-rank path (K1 per
            rank, token blocks, NCCL all-gather, K4c merge, pattern broadcast, K2 per rank); every rank's pattern set must be
            the one a sequential run computes, and the hit records of all ranks, gathered to rank 0 and replayed in global
            read order into ONE set of containers, must give the reference's dump.
`configs` : the other BASELINE.json configs as bounded legs of the same run (skipped by --no-extra):
            config3 2M long reads (1-10 kb), config4 100M x 150 bp cut over the N GPUs (strong scaling), config5 the
            singleton scan against 100..20k patterns over 50M reads (sharded over the N GPUs).
--config C makes config C the headline instead (its metric, value and roofline).
--impl reference: the reference's CPU implementation on all host cores (one process per core over disjoint shards, the only
            parallelism the reference supports) on a bounded sample of the same workload.
"""
import argparse
import hashlib
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
TOK = 64                                                            # bytes per K4 token record (>= high_dr + 2)


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


def write_fasta(path, bases, offsets, lo, hi, first_name=None):
    """reads [lo, hi) as FASTA, one line per read; names r%010d from `first_name` (default lo)"""
    first_name = lo if first_name is None else first_name
    L = int(offsets[lo + 1] - offsets[lo]) if hi > lo else 0
    fixed = hi > lo and bool(np.all(np.diff(offsets[lo:hi + 1].astype(np.int64)) == L))
    if fixed:
        blk = bases[int(offsets[lo]):int(offsets[hi])].reshape(hi - lo, L)
        hdr = np.frombuffer(b"".join(b">r%010d\n" % (first_name + i) for i in range(hi - lo)), dtype=np.uint8).reshape(hi - lo, 13)
        np.concatenate([hdr, blk, np.full((hi - lo, 1), 10, dtype=np.uint8)], axis=1).tofile(path)
        return
    with open(path, "wb") as fh:
        for i in range(lo, hi):
            fh.write(b">r%010d\n" % (first_name + i - lo))
            fh.write(bases[int(offsets[i]):int(offsets[i + 1])].tobytes())
            fh.write(b"\n")


def ref_worker(args):
    """One reference process over one FASTA shard; returns (reads, seconds phase1, cluster, phase2, wall, dump or None)."""
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
            dump, tm = P.run_files([path])
            res.append((n, tm[0] / 1e3, tm[1] / 1e3, tm[2] / 1e3, 0.0, dump if keep_dump else None))
    wall = time.time() - t0
    worst = max(r[1] + r[2] + r[3] for r in res)
    n_bases = int(offsets[per * n_procs] - offsets[0])
    return dict(kind=kind, reads=per * n_procs, bases=n_bases, seconds=worst, wall=wall, rate=per * n_procs / worst,
                phase1_s=max(r[1] for r in res), phase2_s=max(r[3] for r in res),
                dump=res[0][5] if keep_dump else None, path=jobs[0][0])


# ------------------------------------------------------------------------------------------------- workloads
def sample_variable_torch(genome, n_reads, len_lo, len_hi, seed, device, sub_rate=0.001, n_rate=0.0005, chunk=1 << 14):
    """config 3 reads (length U[len_lo, len_hi], both strands, substitutions, N) gathered on the device; the same recipe as
    crass_b200.synth.sample_variable.  Returns (bases uint8 tensor, offsets int64 tensor, max_len)."""
    import torch
    from crass_b200 import synth
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    dev = torch.device(device)
    G = torch.from_numpy(genome).to(dev)
    comp = torch.from_numpy(synth._COMP).to(dev)
    acgt = torch.from_numpy(synth._ACGT.copy()).to(dev)
    lens = torch.randint(len_lo, len_hi + 1, (n_reads,), generator=g, dtype=torch.int64)
    offsets = torch.zeros(n_reads + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(lens, 0)
    out = torch.empty(int(offsets[-1]), dtype=torch.uint8, device=dev)
    for lo in range(0, n_reads, chunk):
        m = min(chunk, n_reads - lo)
        l = lens[lo:lo + m].to(dev)
        starts = torch.randint(0, len(genome) - len_hi, (m,), generator=g, dtype=torch.int64).to(dev)
        rev = torch.randint(0, 2, (m,), generator=g, dtype=torch.int64).to(dev).bool()
        rid = torch.repeat_interleave(torch.arange(m, device=dev), l)
        first = torch.cumsum(l, 0) - l
        within = torch.arange(int(l.sum()), device=dev) - first[rid]
        pos = torch.where(rev[rid], starts[rid] + l[rid] - 1 - within, starts[rid] + within)
        blk = G[pos]
        blk = torch.where(rev[rid], comp[blk.long()], blk)
        nn = blk.numel()
        k = int(torch.binomial(torch.tensor(float(nn)), torch.tensor(sub_rate), generator=g).item()) if sub_rate > 0 else 0
        if k:
            blk[torch.randint(0, nn, (k,), generator=g, dtype=torch.int64).to(dev)] = acgt[torch.randint(0, 4, (k,), generator=g, dtype=torch.int64).to(dev)]
        k = int(torch.binomial(torch.tensor(float(nn)), torch.tensor(n_rate), generator=g).item()) if n_rate > 0 else 0
        if k:
            blk[torch.randint(0, nn, (k,), generator=g, dtype=torch.int64).to(dev)] = ord("N")
        b0 = int(offsets[lo])
        out[b0:b0 + nn] = blk
    return out, offsets.to(dev), int(lens.max())


class Resident:
    """One rank's device-resident pass over a batch: K1 -> hit ordering -> token exchange + createNonRedundantSet ->
    matcher -> K2 -> hit ordering -> both hit lists in pinned host memory.  step(record) runs it once."""

    HOST_KEYS = ("k1_wait", "exchange_cluster", "matcher_build", "ac_upload", "k2_wait", "fetch_hits2")

    def __init__(self, ctx, dev, world, d_bases, d_offsets, n, max_len, params, stream, hits_frac=4, pool_per_read=1.0, cap=8192):
        import torch
        from crass_b200 import dist as cbdist
        self.torch, self.ctx, self.dev, self.world, self.n, self.max_len, self.params, self.stream = torch, ctx, dev, world, n, max_len, params, stream
        self.d_bases, self.d_offsets = d_bases, d_offsets
        self.hits_cap, self.pool_cap = n // hits_frac + 1024, int(n * pool_per_read) + 4096
        self.d_found = torch.empty(n, dtype=torch.uint8, device=dev)
        self.d_found2 = torch.empty(n, dtype=torch.uint8, device=dev)
        self.d_hits = torch.empty(self.hits_cap * 4, dtype=torch.int32, device=dev)
        self.d_sorted = torch.empty(self.hits_cap * 4, dtype=torch.int32, device=dev)
        self.d_pool = torch.empty(self.pool_cap, dtype=torch.int32, device=dev)
        self.d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
        self.h_cnt = torch.zeros(8, dtype=torch.int32, pin_memory=True)
        self.d_tokens = torch.empty(self.hits_cap * TOK, dtype=torch.uint8, device=dev)
        self.h_hits = [torch.empty(self.hits_cap * 4, dtype=torch.int32, pin_memory=True) for _ in range(2)]
        self.h_pool = [torch.empty(self.pool_cap, dtype=torch.int32, pin_memory=True) for _ in range(2)]
        # token blocks of `cap` records per rank (they double by themselves if a shard ever holds more distinct DRs);
        # every rank merges the gathered token blocks and clusters them on its own GPU (K5: the whole of
        # createNonRedundantSet runs as kernels), so nothing is broadcast; CRASS_B200_EXCHANGE=root selects the
        # round-1 form (the root clusters on its host, one broadcast returns the pattern set) for comparison
        self.root_form = world > 1 and os.environ.get("CRASS_B200_EXCHANGE", "") == "root"
        if self.root_form:
            self.exchange = cbdist.PatternExchange(ctx, dev, shard_reads=n, kmer_clust=params.kmer_clust, stride=TOK, cap=cap)
        else:
            self.exchange = cbdist.TokenExchange(ctx, dev, shard_reads=n, stride=TOK, cap=cap)
        self.kt = {"k1": [], "k2": []}
        self.host_ms = {k: [] for k in self.HOST_KEYS}
        self.stats = {}
        self.last = None                                                # (hits1, pool1, hits2, pool2, pattern text) of the last step

    def read_counters(self):
        self.h_cnt.copy_(self.d_cnt, non_blocking=False)               # 32 bytes, synchronises the stream
        nh, npool, ovf = int(self.h_cnt[0]), int(self.h_cnt[1]), int(self.h_cnt[2])
        assert not ovf, "bench hit buffers overflowed"
        return nh, npool

    def fetch_async(self, which, nh, npool):
        self.h_hits[which][: nh * 4].copy_(self.d_sorted[: nh * 4], non_blocking=True)
        self.h_pool[which][: max(npool, 1)].copy_(self.d_pool[: max(npool, 1)], non_blocking=True)

    def host_hits(self, which, nh, npool):
        from crass_b200 import api
        return self.h_hits[which][: nh * 4].numpy().view(api.HIT_DTYPE), self.h_pool[which][: max(npool, 1)].numpy().view(np.uint32)

    def step(self, record, keep=False):
        import crass_b200 as cb
        torch, ctx, n = self.torch, self.ctx, self.n
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        ctx.set_token_output(self.d_tokens, TOK)                       # K4: DR tokens are extracted where the hits are found
        ctx.dr_search_dev(self.d_bases, self.d_offsets, n, self.max_len, self.params, self.d_found, self.d_hits, self.d_pool, self.d_cnt, self.stream)
        ctx.set_token_output(None)
        e[1].record()
        ctx.sort_hits_dev(self.d_found, n, self.d_hits, self.d_cnt, self.hits_cap, self.d_sorted, self.stream)   # read order (what replay consumes)
        t0 = time.perf_counter()
        nh, npool = self.read_counters()
        t1 = time.perf_counter()
        self.fetch_async(0, nh, npool)                                 # the phase-1 hit records travel while the host clusters
        pat_text = None
        if not self.root_form:
            ac, nu = self.exchange.run_matcher(self.d_hits, nh, self.d_tokens, self.params.kmer_clust, self.stream)
            t3 = t4 = time.perf_counter()
        else:
            pat_text, nu = self.exchange.run(self.d_hits, nh, self.d_tokens, self.stream)   # the pattern set, clustered once on rank 0
            t3 = time.perf_counter()
            ac = cb.Automaton.from_pattern_text(pat_text) if pat_text else None
            t4 = time.perf_counter()
        pats = ac.num_patterns if ac else 0
        n2 = npool2 = 0
        t5 = t6 = t4
        if pats:
            ctx.ac_upload(ac)
            t5 = time.perf_counter()
            e[2].record()
            ctx.ac_scan_dev(ac, self.d_bases, self.d_offsets, n, self.max_len, self.d_found, self.d_found2, self.d_hits, self.d_pool, self.d_cnt, self.stream)
            e[3].record()
            ctx.sort_hits_dev(self.d_found2, n, self.d_hits, self.d_cnt, self.hits_cap, self.d_sorted, self.stream)
            n2, npool2 = self.read_counters()
            t6 = time.perf_counter()
            self.fetch_async(1, n2, npool2)
        torch.cuda.synchronize()                                       # both hit lists are on the host now, in read order
        t7 = time.perf_counter()
        if record:
            self.kt["k1"].append(e[0].elapsed_time(e[1]))
            if pats:
                self.kt["k2"].append(e[2].elapsed_time(e[3]))
                for k, v in zip(self.HOST_KEYS, (t1 - t0, t3 - t1, t4 - t3, t5 - t4, t6 - t5, t7 - t6)):
                    self.host_ms[k].append(v * 1e3)
                for k, v in getattr(self.exchange, "last_ms", {}).items():           # inside exchange_cluster (rank 0's view)
                    self.host_ms.setdefault("exchange:" + k, []).append(v)
        self.stats.update(hits_phase1=nh, ss_entries_phase1=npool, dr_variants_merged=nu, patterns=pats, hits_phase2=n2)
        if keep:
            h1, p1 = self.host_hits(0, nh, npool)
            h2, p2 = self.host_hits(1, n2, npool2)
            if pat_text is None and ac is not None:
                pat_text = ac.pattern_text()
            self.last = (h1.copy(), p1.copy(), h2.copy(), p2.copy(), pat_text)

    def summary(self, n_bases, peak):
        k1 = float(np.mean(self.kt["k1"]))
        k2 = float(np.mean(self.kt["k2"])) if self.kt["k2"] else 0.0
        s = self.stats
        b1 = n_bases + 9 * self.n + 8 * s["hits_phase1"] + 4 * s["ss_entries_phase1"]   # L + 8 B/read in, 1 B flag, 8+8n B per hit
        b2 = n_bases + 10 * self.n + 16 * s["hits_phase2"]
        return dict(k1_ms=k1, k2_ms=k2, bytes_k1=b1, bytes_k2=b2,
                    k1_frac=b1 / (k1 / 1e3) / 1e9 / peak, k2_frac=(b2 / (k2 / 1e3) / 1e9 / peak) if k2 else None,
                    k12_frac=((b1 + b2) / ((k1 + k2) / 1e3) / 1e9 / peak) if k2 else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json config that is the headline")
    ap.add_argument("--reads", type=int, default=None, help="reads per GPU of the headline config (default: the config's own size)")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the legs of the other configs")
    ap.add_argument("--no-e2e", action="store_true")
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
    C = args.config
    default_reads = {2: 10_000_000, 3: 2_000_000 // world, 4: 100_000_000 // world, 5: 50_000_000 // world}[C]
    n_reads = args.reads or default_reads
    scaling = "weak" if C == 2 else "strong"
    workloads = {
        2: "config2: synthetic %d x %dbp Illumina reads per GPU, 50 planted CRISPR DR types, 1%% array bases, 0.1%% subs, 0.05%% N, default crass options" % (n_reads, READ_LEN),
        3: "config3: synthetic %d long reads (1-10 kb) in all, 3%% array bases, cut into contiguous shards over the GPUs, default crass options" % (n_reads * world),
        4: "config4: synthetic %d x %dbp metagenome reads in all (the config2 recipe), cut into contiguous shards over the GPUs, NCCL DR-set all-gather" % (n_reads * world, READ_LEN),
        5: "config5: singleton scan of %d x %dbp reads in all (sharded) against 100..20000 patterns, every read scanned, 1%% planted; headline = 3000 patterns" % (n_reads * world, READ_LEN),
    }
    config = {"workload": workloads[C], "reads_per_gpu": n_reads, "read_len": READ_LEN if C != 3 else "U[1000,10000]", "seed": SEED + C - 2,
              "sharding": "contiguous read ranges, one rank per GPU",
              "l2": "inputs (>= %.2f GB per GPU) are larger than L2 (126 MB); no flush needed" % (n_reads * (READ_LEN + 8) / 1e9)}

    from crass_b200 import synth
    genome, drs, _ = synth.make_genome(SEED)

    # ------------------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        n_procs = os.cpu_count() or 1
        if C == 3:
            g3, _, _ = synth.make_genome(SEED + 1, array_fraction=0.03, min_spacers=60, max_spacers=200)
            n_sample = max(n_procs, min(n_reads, 2_000 * n_procs))
            n_sample -= n_sample % n_procs
            bases, offsets = synth.sample_variable(g3, n_sample, 1000, 10000, SEED + 1001)
        else:
            n_sample = max(n_procs, min(n_reads, 4_000_000, 250_000 * n_procs))
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
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": dict(config, sample_reads=n_sample),
                "gbp_per_s": rate * vals[0]["bases"] / vals[0]["reads"] / 1e9,
                "cpu_baseline": {"value": rate, "unit": "reads/s", "cores": n_procs, "kind": vals[0]["kind"],
                                 "sample": "%d-read prefix of the config%d recipe split over %d processes (searchFile + createNonRedundantSet + findSingletons each, FASTA in tmpfs)" % (n_sample, 2 if C in (4, 5) else C, n_procs)},
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
    peak, peak_src = peaks()

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
    os.environ.setdefault("CRASS_B200_PARSE_THREADS", str(max(1, min(16, cores // max(local_world, 1)))))
    ctx = cb.Context(local_rank)
    ctx.keep_packed(True)                                # K2 reads the 2-bit stream K1's filter leaves in HBM (same, unchanged batch)
    params = cb.Params()
    work_stream = torch.cuda.Stream(device=dev)          # a real (non-NULL) stream: kernels, copies and events all go here
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    assert stream != 0

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

    def allsum(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------------------------------------------------------------------------- legs
    def leg_fixed(n, seed, steps, warmup, label):
        """config 2 / 4: n x 150 bp reads of the config-2 recipe on this rank, device resident -> dict"""
        d_bases, d_offsets = synth.sample_fixed_torch(genome, n, READ_LEN, seed, dev)
        d_offsets = d_offsets.to(torch.int64)
        R = Resident(ctx, dev, world, d_bases, d_offsets, n, READ_LEN, params, stream)
        for _ in range(warmup):
            R.step(False)
        launches0 = ctx.launch_count
        ms = timed(lambda: R.step(True), steps)
        launches = ctx.launch_count - launches0
        return R, d_bases, d_offsets, ms, launches

    def leg_config3(n, steps, warmup):
        """2M long reads in all, n on this rank (contiguous shard = own seed), device resident"""
        g3, _, _ = synth.make_genome(SEED + 1, array_fraction=0.03, min_spacers=60, max_spacers=200)
        d_bases, d_offsets, max_len = sample_variable_torch(g3, n, 1000, 10000, SEED + 1001 + rank, dev)
        mx = torch.tensor([max_len], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        R = Resident(ctx, dev, world, d_bases, d_offsets, n, int(mx.item()), params, stream, hits_frac=1, pool_per_read=64.0, cap=65536)
        for _ in range(warmup):
            R.step(False)
        ms = timed(lambda: R.step(True), steps)
        n_bases = int(d_offsets[-1].item())
        s = R.summary(n_bases, peak)
        tot_reads, tot_bases = allsum(n), allsum(n_bases)
        out = {"workload": "config3: %d long reads U[1000,10000] bp in all (%.2f Gbp), 3%% array bases, %d per GPU" % (int(tot_reads), tot_bases / 1e9, n),
               "ms_per_step": ms / steps, "reads_per_s": tot_reads / (ms / steps / 1e3), "gbp_per_s": tot_bases / (ms / steps / 1e3) / 1e9,
               "k1_ms": s["k1_ms"], "k1_gbp_per_s": n_bases / s["k1_ms"] / 1e6, "k1_frac_of_hbm": s["k1_frac"],
               "k2_ms": s["k2_ms"], "k2_gbp_per_s": (n_bases / s["k2_ms"] / 1e6) if s["k2_ms"] else None, "k2_frac_of_hbm": s["k2_frac"],
               "k1_plus_k2_frac_of_hbm": s["k12_frac"], "launches": ["k_dr_long", "k_ac_filter_long", "k_ac_verify_warp"], "stats": dict(R.stats)}
        if rank == 0 and not args.no_cpu_baseline:                      # parity of a prefix against the oracle (found flags of phase 1)
            import checkers
            P = checkers.port()
            m = min(2000, n)
            hb = d_bases[: int(d_offsets[m].item())].cpu().numpy()
            ho = d_offsets[: m + 1].cpu().numpy().astype(np.uint64)
            want = np.zeros(m, dtype=np.uint8)
            P.lib.orc_phase1_batch(hb.ctypes.data, ho.ctypes.data, m, checkers.params_array(), want.ctypes.data)
            out["parity"] = {"prefix_reads": m, "oracle_hits": int(want.sum()), "found_flags_identical": bool(np.array_equal(R.d_found[:m].cpu().numpy(), want))}
        del R, d_bases, d_offsets
        torch.cuda.empty_cache()
        return out, ms, s

    def leg_config5(n, steps, pattern_counts=(100, 300, 1000, 3000, 10000, 20000)):
        """the singleton scan alone: n uniform reads on this rank, every read scanned, pattern sets of growing size"""
        gen = torch.Generator(device=dev)
        gen.manual_seed(SEED + 3 + rank)
        acgt = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
        d_bases = torch.empty(n * READ_LEN, dtype=torch.uint8, device=dev)
        for lo in range(0, d_bases.numel(), 1 << 28):
            m = min(1 << 28, d_bases.numel() - lo)
            d_bases[lo:lo + m] = acgt[torch.randint(0, 4, (m,), generator=gen, device=dev)]
        d_offsets = torch.arange(n + 1, dtype=torch.int64, device=dev) * READ_LEN
        d_skip = torch.zeros(n, dtype=torch.uint8, device=dev)
        d_found = torch.empty(n, dtype=torch.uint8, device=dev)
        d_found1 = torch.empty(n, dtype=torch.uint8, device=dev)
        hits_cap, pool_cap = n // 16 + 4096, n // 4 + 4096
        d_hits = torch.empty(hits_cap * 4, dtype=torch.int32, device=dev)
        d_pool = torch.empty(pool_cap, dtype=torch.int32, device=dev)
        d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
        rows = []
        for P_ in pattern_counts:
            patterns = synth.pattern_set(P_, seed=SEED + 3 + P_)
            # one occurrence of a random pattern in 1 % of the reads (earlier sets stay in the reads as background)
            mat = np.zeros((len(patterns), 47), dtype=np.uint8)
            lens = np.zeros(len(patterns), dtype=np.int64)
            for i, p in enumerate(patterns):
                mat[i, :len(p)] = np.frombuffer(p, dtype=np.uint8)
                lens[i] = len(p)
            d_mat, d_len = torch.from_numpy(mat).to(dev), torch.from_numpy(lens).to(dev)
            pick = torch.nonzero(torch.rand(n, generator=gen, device=dev) < 0.01).flatten()
            which = torch.randint(0, len(patterns), (pick.numel(),), generator=gen, device=dev)
            plen = d_len[which]
            at = torch.minimum((torch.rand(pick.numel(), generator=gen, device=dev) * (READ_LEN - plen + 1).float()).long().clamp_(min=0), READ_LEN - plen)
            j = torch.arange(47, device=dev)
            mask = j[None, :] < plen[:, None]
            d_bases[((pick * READ_LEN + at)[:, None] + j[None, :])[mask]] = d_mat[which][mask]
            torch.cuda.synchronize()                                   # (the planting above is asynchronous: it must not land in the build time)
            t0 = time.perf_counter()
            ac = cb.Automaton(patterns)
            ctx.ac_upload(ac)
            torch.cuda.synchronize()
            build_ms = (time.perf_counter() - t0) * 1e3
            # the pipeline's form: K2 on the 2-bit stream phase 1 of the same batch leaves behind
            ctx.dr_search_dev(d_bases, d_offsets, n, READ_LEN, params, d_found1, d_hits, d_pool, d_cnt, stream)
            ks = []

            def one():
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ctx.ac_scan_dev(ac, d_bases, d_offsets, n, READ_LEN, d_skip, d_found, d_hits, d_pool, d_cnt, stream)
                b.record()
                ks.append((a, b))
            for _ in range(3):
                one()
            ks.clear()
            ms = timed(one, steps)
            k2 = float(np.mean([a.elapsed_time(b) for a, b in ks]))
            c = d_cnt.cpu().numpy()
            assert not c[2], "hit buffers overflowed"
            nh = int(c[0])
            alg = n * (READ_LEN + 8 + 1 + 1) + 16 * nh
            row = {"patterns": P_, "reads": int(allsum(n)), "hits": int(allsum(nh)), "k2_ms": k2, "ms_per_step": ms / steps,
                   "reads_per_s": allsum(n) / (ms / steps / 1e3), "gbp_per_s": allsum(n) * READ_LEN / (ms / steps / 1e3) / 1e9,
                   "k2_frac_of_hbm": alg / (k2 / 1e3) / 1e9 / peak, "matcher_build_upload_ms": build_ms}
            if rank == 0 and not args.no_cpu_baseline:                  # found flags of a prefix against the oracle's acism restatement
                import checkers
                O = checkers.port()
                m = min(100_000, n)
                hb = d_bases[: m * READ_LEN].cpu().numpy()
                ho = np.arange(m + 1, dtype=np.uint64) * READ_LEN
                want = np.zeros(m, dtype=np.uint8)
                oh = O.ac_create(patterns)
                O.lib.orc_phase2_batch(oh, hb.ctypes.data, ho.ctypes.data, m, want.ctypes.data)
                O.ac_destroy(oh)
                row["parity"] = {"prefix_reads": m, "oracle_hits": int(want.sum()), "found_flags_identical": bool(np.array_equal(d_found[:m].cpu().numpy(), want))}
            rows.append(row)
            del ac
        del d_bases, d_offsets, d_skip, d_found, d_found1, d_hits, d_pool
        torch.cuda.empty_cache()
        return rows

    def merged_parity(total_reads):
        """N > 1: a total_reads prefix cut into `world` contiguous shards through the N-rank path; see the module docstring."""
        per = total_reads // world
        total = per * world
        bases, offsets = synth.sample_fixed(genome, total, READ_LEN, SEED + 5000)       # every rank draws the same sample
        lo, hi = rank * per, (rank + 1) * per
        d_b = torch.from_numpy(bases[lo * READ_LEN: hi * READ_LEN]).to(dev)
        d_o = torch.arange(per + 1, dtype=torch.int64, device=dev) * READ_LEN
        R = Resident(ctx, dev, world, d_b, d_o, per, READ_LEN, params, stream)
        R.step(False, keep=True)
        h1, p1, h2, p2, pat_text = R.last
        my_md5 = hashlib.md5(pat_text or b"").hexdigest()
        # hit records of every rank -> rank 0 (sizes, then padded payloads: tiny next to the reads)
        def gather(arr_u32):
            t = torch.from_numpy(arr_u32.astype(np.uint32).view(np.int32)).to(dev)
            sz = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
            sizes = torch.zeros(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(sizes, sz)
            sizes = sizes.cpu().tolist()
            mx = max(max(sizes), 1)
            buf = torch.zeros(mx, dtype=torch.int32, device=dev)
            buf[: t.numel()] = t
            allb = torch.empty(world * mx, dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(allb, buf)
            allb = allb.cpu().numpy().view(np.uint32)
            return [allb[r * mx: r * mx + sizes[r]] for r in range(world)]
        parts = [gather(x.view(np.uint32).reshape(-1)) for x in (h1, p1, h2, p2)]
        md5s = [None] * world
        dist.all_gather_object(md5s, my_md5)
        out = None
        if rank == 0:
            import checkers
            def concat(hp, pp):
                hits, pools, pool0 = [], [], 0
                for r in range(world):
                    h = hp[r].view(api.HIT_DTYPE).copy()
                    h["read_index"] += r * per
                    h["ss_offset"] += pool0
                    pool0 += len(pp[r])
                    hits.append(h)
                    pools.append(pp[r])
                return np.concatenate(hits), np.concatenate(pools) if pools else np.zeros(0, np.uint32)
            H1, P1 = concat(parts[0], parts[1])
            H2, P2 = concat(parts[2], parts[3])
            batch = cb.Batch.from_arrays(bases, offsets)
            res = cb.Results()
            res.add_phase1(batch, H1, P1)
            seq_pats = res.non_redundant(params.kmer_clust)              # what ONE process computes from its single token list
            seq_md5 = hashlib.md5(b"".join(p + b"\n" for p in seq_pats)).hexdigest()
            res.add_phase2(batch, H2, P2)
            mine = res.dump(READ_LEN)
            ref_dump, kind = None, None
            with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
                path = os.path.join(d, "merged.fa")
                write_fasta(path, bases, offsets, 0, total)
                if checkers.have_ref():
                    ref_dump, kind = checkers.ref().run_files([path])[0], "reference"
                else:
                    ref_dump, kind = checkers.port().run_files([path])[0], "port"
                # ... and the engine's single-GPU whole path on the same file
                eng = cb.Engine((local_rank,))
                r1, ml = eng.run_files([path])
                single = r1.dump(ml)
                eng.close()
            out = {"sample": "%d-read prefix of the config2 recipe in %d contiguous shards, one per rank" % (total, world),
                   "pattern_set_md5_per_rank": md5s, "pattern_set_md5_sequential": seq_md5,
                   "pattern_sets_identical": bool(all(m == seq_md5 for m in md5s)),
                   "dump_identical_to_" + kind: bool(mine == ref_dump), "dump_identical_to_single_gpu": bool(mine == single),
                   "dump_bytes": len(mine), "found_reads": int(res.num_reads), "tokens": int(res.num_tokens),
                   "merged_identical": bool(all(m == seq_md5 for m in md5s) and mine == ref_dump and mine == single)}
        del R, d_b, d_o
        torch.cuda.empty_cache()
        return out

    # ---------------------------------------------------------------------------------- the run
    sampler = ClockSampler(local_rank) if rank == 0 else None         # one clock poller per job, not per rank
    if sampler:
        sampler.start()
    extra = {}
    line = None
    if C in (2, 4):
        R, d_bases, d_offsets, ms_total, launches = leg_fixed(n_reads, SEED + 1000 + rank, args.steps, args.warmup, "config%d" % C)
        n_bases = n_reads * READ_LEN
        s = R.summary(n_bases, peak)
        host_breakdown = {k: float(np.mean(v)) for k, v in R.host_ms.items() if v}
        stats = dict(R.stats)
        total_reads = n_reads * world
        ms_step = ms_total / args.steps
        value = total_reads / (ms_step / 1e3)
        dom, dom_ms, dom_bytes = ("K1 dr_search", s["k1_ms"], s["bytes_k1"]) if s["k1_ms"] >= s["k2_ms"] else ("K2 singleton_scan", s["k2_ms"], s["bytes_k2"])
        traffic = None                                                         # measured DRAM bytes per launch, from the committed ncu capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["config2_%dx%d" % (n_reads, READ_LEN)][dom]["bytes"]
        except (OSError, KeyError, ValueError):
            pass
        # ---- e2e: file in -> containers out, the shard's FASTA in tmpfs through the engine
        e2e = e2e_hostbuf = None
        if not args.no_e2e:
            h_bases = torch.empty(d_bases.shape, dtype=torch.uint8, pin_memory=True)
            h_offsets = torch.empty(d_offsets.shape, dtype=torch.int64, pin_memory=True)
            h_bases.copy_(d_bases)
            h_offsets.copy_(d_offsets)
            torch.cuda.synchronize()
            np_bases, np_offsets = h_bases.numpy(), h_offsets.numpy().view(np.uint64)
            if not threads_given:                                      # every rank clusters its own shard here: share the cores evenly
                os.environ["CRASS_B200_HOST_THREADS"] = str(max(1, min(16, cores // max(local_world, 1))))

            def step_hostbuf():
                ctx.upload(h_bases, h_offsets)                         # H2D from pinned host memory
                hits, pool, _ = ctx.dr_search_resident(params)
                merged = cbdist.allgather_dr_lists(ctx.last_dr_list(), device=dev)
                nb = hits.nbytes + pool.nbytes
                if merged:
                    ac = cb.Automaton.from_dr_list(merged, params.kmer_clust)
                    hits2, pool2, _ = ctx.ac_scan_resident(ac, skip_found=True)
                    nb += hits2.nbytes + pool2.nbytes
                return nb
            for _ in range(min(args.warmup, 2)):
                step_hostbuf()
            d2h = [0]
            ms_hb = timed(lambda: d2h.__setitem__(0, step_hostbuf()), args.steps)
            e2e_hostbuf = {"value": total_reads / (ms_hb / args.steps / 1e3), "unit": "reads/s", "ms_per_step": ms_hb / args.steps,
                           "h2d_bytes_per_step": int(h_bases.numel() + h_offsets.numel() * 8), "d2h_bytes_per_step": int(d2h[0]),
                           "what": "pinned host arrays in, hit records out (no parsing, no replay)"}
            R = d_bases = d_offsets = None                          # the resident leg's buffers go before the file leg starts
            torch.cuda.empty_cache()
            tmp = tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            fasta = os.path.join(tmp.name, "shard%d.fa" % rank)
            write_fasta(fasta, np_bases, np_offsets, 0, n_reads, first_name=rank * n_reads)
            eng = cb.Engine((local_rank,))
            info = {}

            kept = []                                                  # the containers of the timed runs are torn down after the clock stops
                                                                       # (the reference arm times searchFile + findSingletons, not the destructors)
            def step_file():
                res, ml = eng.run_files([fasta])
                info.update(found_reads=int(res.num_reads), tokens=int(res.num_tokens), stage_ms=eng.stage_ms())
                kept.append(res)
            cold_ms = []                                               # the engine's first runs: ordinary memory + staged copies, then the
            for _ in range(min(args.warmup, 2)):                       # run that page-locks the pooled buffers (CRASS_B200_PIN=auto)
                t0 = time.time()
                step_file()
                cold_ms.append((time.time() - t0) * 1e3)
            del kept[:]
            b0 = eng.transfer_bytes()
            l0 = eng.launch_count
            ms_file = timed(step_file, args.steps)
            b1 = eng.transfer_bytes()
            del kept[:]
            e2e = {"value": total_reads / (ms_file / args.steps / 1e3), "unit": "reads/s", "ms_per_step": ms_file / args.steps,
                   "h2d_bytes_per_step": int((b1[0] - b0[0]) // args.steps), "d2h_bytes_per_step": int((b1[1] - b0[1]) // args.steps),
                   "what": "crass_b200_engine_run_files on the shard's FASTA (%d bytes, tmpfs), streamed in ranges of 128 MB: parse of range i+1 || pinned -> H2D -> K1 -> D2H of range i || replay of range i-1 into the containers; then clustering -> K2 over the resident ranges -> D2H -> replay (stage_ms overlap)" % os.path.getsize(fasta),
                   "file_bytes": os.path.getsize(fasta), "gpu_launches_per_step": int((eng.launch_count - l0) // args.steps),
                   "stage_ms": info.get("stage_ms"), "found_reads": info.get("found_reads"), "tokens": info.get("tokens"),
                   "warmup_runs_ms": cold_ms,                          # [0] = a cold engine (what a one-shot caller sees), [1] = the run that page-locks
                   "timed_runs": "steady state: an engine that is used again (pooled page-locked buffers)"}
            eng.close()
        line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config, "gbp_per_s": value * READ_LEN / 1e9,
                "e2e": e2e, "e2e_hostbuf": e2e_hostbuf, "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_bytes / (dom_ms / 1e3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                             "launches": (["k_dr_filter_warp", "k_dr_exact_staged"] if dom.startswith("K1") else ["k_ac_filter_packed", "k_ac_verify_mask"]),
                             "frac": dom_bytes / (dom_ms / 1e3) / 1e9 / peak, "traffic": traffic, "algorithmic_bytes_per_launch": int(dom_bytes),
                             "kernel_ms": dom_ms},
                "kernels": {"k1_dr_search_ms": s["k1_ms"], "k1_frac_of_hbm": s["k1_frac"], "k2_singleton_scan_ms": s["k2_ms"], "k2_frac_of_hbm": s["k2_frac"],
                            "k1_plus_k2_frac_of_hbm": s["k12_frac"], "host_between_kernels_ms": ms_step - s["k1_ms"] - s["k2_ms"],
                            "host_breakdown_ms": host_breakdown},
                "stats": stats}
    elif C == 3:
        out3, ms_total, s = leg_config3(n_reads, args.steps, args.warmup)
        dom, dom_ms, dom_bytes = ("K1 dr_search (k_dr_long)", s["k1_ms"], s["bytes_k1"]) if s["k1_ms"] >= s["k2_ms"] else ("K2 singleton_scan (long)", s["k2_ms"], s["bytes_k2"])
        line = {"metric": METRIC, "value": out3["reads_per_s"], "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": out3["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config, "gbp_per_s": out3["gbp_per_s"], "e2e": None, "gpu_launches": None,
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_bytes / (dom_ms / 1e3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                             "frac": dom_bytes / (dom_ms / 1e3) / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_launch": int(dom_bytes), "kernel_ms": dom_ms},
                "kernels": out3, "stats": out3["stats"]}
    else:
        rows = leg_config5(n_reads, args.steps)
        head = [r for r in rows if r["patterns"] == 3000][0]
        line = {"metric": "reads_per_s_singleton_scan", "value": head["reads_per_s"], "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config, "gbp_per_s": head["gbp_per_s"], "e2e": None, "gpu_launches": None,
                "roofline": {"bound": "hbm", "kernel": "K2 singleton_scan", "achieved": head["k2_frac_of_hbm"] * peak, "peak": peak, "peak_source": peak_src,
                             "unit": "GB/s", "frac": head["k2_frac_of_hbm"], "traffic": None, "kernel_ms": head["k2_ms"]},
                "sweep": rows}
    # ---------------------------------------------------------------------------------- the other configs
    if not args.no_extra:
        if C != 3:
            try:
                extra["config3"] = leg_config3(max(1, 2_000_000 // world), min(args.steps, 3), 2)[0]
            except Exception as ex:                                    # noqa: BLE001 -- a leg must not take the headline down
                extra["config3"] = {"error": repr(ex)}
        if C != 5:
            try:
                extra["config5"] = {"workload": "config5: %d x %dbp uniform reads in all (sharded), every read scanned, 1%% planted" % (50_000_000 // world * world, READ_LEN),
                                    "sweep": leg_config5(50_000_000 // world, min(args.steps, 3))}
            except Exception as ex:                                    # noqa: BLE001
                extra["config5"] = {"error": repr(ex)}
        if C != 4:
            try:
                n4 = 100_000_000 // world
                R4, b4, o4, ms4, _ = leg_fixed(n4, SEED + 2000 + rank, min(args.steps, 3), 2, "config4")
                s4 = R4.summary(n4 * READ_LEN, peak)
                st = ms4 / min(args.steps, 3)
                extra["config4"] = {"workload": "config4: %d x %dbp reads in all, %d per GPU (strong scaling of a fixed 100M-read set)" % (n4 * world, READ_LEN, n4),
                                    "ms_per_step": st, "reads_per_s": n4 * world / (st / 1e3), "gbp_per_s": n4 * world * READ_LEN / (st / 1e3) / 1e9,
                                    "k1_ms": s4["k1_ms"], "k1_frac_of_hbm": s4["k1_frac"], "k2_ms": s4["k2_ms"], "k2_frac_of_hbm": s4["k2_frac"],
                                    "k1_plus_k2_frac_of_hbm": s4["k12_frac"], "host_breakdown_ms": {k: float(np.mean(v)) for k, v in R4.host_ms.items() if v},
                                    "stats": dict(R4.stats)}
                del R4, b4, o4
                torch.cuda.empty_cache()
            except Exception as ex:                                    # noqa: BLE001
                extra["config4"] = {"error": repr(ex)}
    parity = None
    if world > 1 and not args.no_cpu_baseline:
        parity = merged_parity(max(args.cpu_sample, 1_000_000))
    clocks = sampler.stop() if sampler else None                       # sampled across all timed regions

    if rank == 0:
        line["clocks"] = clocks
        if extra:
            line["configs"] = extra
        if parity:
            line["parity"] = parity
        if not args.no_cpu_baseline and C in (2, 4):
            ns = min(args.cpu_sample, n_reads)
            sb, so = synth.sample_fixed(genome, ns, READ_LEN, SEED + 7000)
            with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
                r = cpu_reference_rate(sb, so, ns, 1, d, keep_dump=True)
                if r["dump"] is not None:
                    # SURVEY 8d "parity check accompanying every timing": the product's whole path (parser -> K1 -> clustering
                    # -> K2 -> replay) on the very FASTA the reference just processed; the two result dumps (tokens in
                    # numbering order, DRs, reads, orientation, start/stops, patterns) must be the same bytes.
                    t0 = time.time()
                    eng = cb.Engine((local_rank,))
                    res, max_len = eng.run_files([r["path"]])
                    t_run = time.time() - t0
                    mine = res.dump(max_len)
                    eng.close()
                    p1 = {"sample": "the cpu_baseline sample (%d reads), whole path through the C-ABI engine vs the %s" % (ns, r["kind"]),
                          "dump_identical": bool(mine == r["dump"]), "dump_bytes": len(r["dump"]),
                          "dump_md5": hashlib.md5(r["dump"].encode("latin-1")).hexdigest(),
                          "found_reads": int(res.num_reads), "tokens": int(res.num_tokens), "b200_seconds": t_run}
                    if parity:
                        line["parity"]["single_gpu_whole_path"] = p1
                    else:
                        line["parity"] = p1
            line["cpu_baseline"] = {"value": r["rate"], "unit": "reads/s", "cores": 1, "kind": r["kind"],
                                    "sample": "%d reads of the config2 recipe as FASTA in tmpfs: searchFile %.2fs + findSingletons %.2fs, 1 thread" % (ns, r["phase1_s"], r["phase2_s"])}
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
