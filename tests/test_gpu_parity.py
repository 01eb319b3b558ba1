"""Parity of the CUDA path against the oracle, through the C-ABI (libcrass_b200.so) -- needs a B200.

Bit-exact for everything: hit/no-hit per read, start/stop lists, repeat length, match offsets, edit
distances, float similarities (compared as bit patterns), token numbering and the whole ReadMap dump.
"""
import gzip
import json
import os
import random
import struct

import numpy as np
import pytest

import checkers
import fuzzgen
import crass_b200 as cb
from crass_b200 import api, synth

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BUNDLED = ["Ill100.fx.gz", "CN_gDC.fa.gz", "Ill.nr.miss.fa.gz", "front_offset_bug.fa.gz", "poor_dr_ext.fa.gz"]


@pytest.fixture(scope="module")
def ctx():
    c = cb.Context(0)
    yield c
    c.close()


K1_ENV = {"fast": {"CRASS_B200_K1": "fast"},
          "fast-r1": {"CRASS_B200_K1": "fast", "CRASS_B200_K1F": "tma", "CRASS_B200_K1E": "lockstep"},
          "fast-refill": {"CRASS_B200_K1": "fast", "CRASS_B200_K1E": "refill"},
          "fast-piped": {"CRASS_B200_K1": "fast", "CRASS_B200_K1_CHUNKS": "3"},
          "fast-r1-piped": {"CRASS_B200_K1": "fast", "CRASS_B200_K1F": "tma", "CRASS_B200_K1E": "lockstep", "CRASS_B200_K1_CHUNKS": "5"},
          "generic": {"CRASS_B200_K1": "generic"}}


@pytest.fixture(params=list(K1_ENV))
def k1path(request):
    """K1's device paths, all with identical results: the 2-bit seed filter + exact candidate kernel (default options,
    reads <= 304 bp) in its warp-tile / staged form ("fast"; "fast-refill": exact kernel with refilled lanes), in the round-1 form (CTA tiles staged by bulk copies,
    32 candidates per warp in lock step: "fast-r1"), either of them cut into chunks with the exact kernel of one chunk on
    a second stream beside the filter of the next ("-piped"), and the generic one-thread-per-read kernel."""
    keys = ("CRASS_B200_K1", "CRASS_B200_K1F", "CRASS_B200_K1E", "CRASS_B200_K1_CHUNKS")
    old = {k: os.environ.get(k) for k in keys}
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(K1_ENV[request.param])
    yield request.param.split("-")[0]
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.fixture(params=["fast", "fast-warp", "fast-list", "generic"])
def k2path(request):
    """K2 likewise: 16-mer q-gram filter + verification of the candidates (one warp per candidate over the starts the
    filter's hits allow; "fast-warp": over all starts; "fast-list": one thread per candidate), or the plain
    one-thread-per-read automaton walk."""
    old = os.environ.get("CRASS_B200_K2"), os.environ.get("CRASS_B200_K2V")
    os.environ["CRASS_B200_K2"] = request.param.split("-")[0]
    os.environ.pop("CRASS_B200_K2V", None)
    if "-" in request.param:
        os.environ["CRASS_B200_K2V"] = request.param.split("-")[1]
    yield request.param
    for k, v in zip(("CRASS_B200_K2", "CRASS_B200_K2V"), old):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.fixture(scope="module")
def P():
    return checkers.port()


def load(name):
    return json.load(open(os.path.join(G, name)))


def hits_by_read(hits, pool):
    return {int(h["read_index"]): (list(map(int, pool[h["ss_offset"]:h["ss_offset"] + h["n_ss"]])), int(h["repeat_len"])) for h in hits}


def check_batch_against_oracle(ctx, P, reads, params=None):
    """dr_search over `reads` must equal the oracle's searchCore read by read."""
    bases, offs = cb.pack_reads(reads)
    prm = cb.Params(**params) if params else cb.Params()
    hits, pool, found = ctx.dr_search(bases, offs, prm)
    got = hits_by_read(hits, pool)
    assert list(hits["read_index"]) == sorted(hits["read_index"])
    n_hit = 0
    for i, s in enumerate(reads):
        f, ss, rl = P.search_core(s, params)
        assert int(found[i]) == (1 if f == 1 else 0), (i, s, params)
        if f == 1:
            n_hit += 1
            assert got[i] == (ss, rl), (i, s, params)
        else:
            assert i not in got
    assert len(got) == n_hit
    return n_hit


# ---- the reference's own known-answer tests (src/test/test_libcrispr.cpp) through the device code ----
def test_catch_scan_right(ctx):
    for v in load("catch_vectors.json")["scan_right"]:
        assert ctx.scan_right(v["seq"].encode(), v["ss"], v["pattern"].encode(), v["min_spacer"], v["scan_range"]) == v["expect"], v["cite"]


def test_catch_extend_pre_repeat(ctx):
    for v in load("catch_vectors.json")["extend_pre_repeat"]:
        assert ctx.extend_pre_repeat(v["seq"].encode(), v["ss"], v["window"], v["min_spacer"]) == (v["expect_len"], v["expect"]), v["cite"]


# ---- golden vectors produced by the reference ---------------------------------------------------------
def test_search_core_golden_vectors(ctx):
    vec = load("search_core_vectors.json")
    groups = {}
    for v in vec:
        groups.setdefault(json.dumps(v["params"], sort_keys=True), []).append(v)
    for key, vs in groups.items():
        params = json.loads(key)
        reads = [v["seq"].encode("latin-1") for v in vs]
        bases, offs = cb.pack_reads(reads)
        hits, pool, found = ctx.dr_search(bases, offs, cb.Params(**params) if params else cb.Params())
        got = hits_by_read(hits, pool)
        for i, v in enumerate(vs):
            assert int(found[i]) == (1 if v["found"] == 1 else 0)
            if v["found"] == 1:
                assert got[i] == (v["ss"], v["replen"])


def test_edit_distance_golden_vectors(ctx):
    vec = load("edit_distance_vectors.json")
    dist, sim = ctx.edit_distance_batch([(a.encode(), b.encode()) for a, b, _, _ in vec])
    for i, (_, _, d, simhex) in enumerate(vec):
        assert int(dist[i]) == d
        assert struct.pack(">f", float(sim[i])).hex() == simhex


def test_ac_golden_vectors(ctx, k2path):
    for case in load("ac_vectors.json"):
        ac = cb.Automaton([p.encode() for p in case["patterns"]])
        texts = [t.encode() for t, _ in case["texts"]]
        bases, offs = cb.pack_reads(texts)
        hits, pool, found = ctx.ac_scan(ac, bases, offs)
        got = hits_by_read(hits, pool)
        for i, (t, expect) in enumerate(case["texts"]):
            if expect is None:
                assert i not in got and found[i] == 0
            else:
                end, plen = expect
                dr_end = min(end - 1, len(t) - 1)
                assert got[i] == ([dr_end - (plen - 1), dr_end], 0)
                assert found[i] == 1


@pytest.mark.parametrize("name", BUNDLED)
def test_bundled_files_whole_path(ctx, name):
    """BASELINE.json configs[0]: searchFile -> createNonRedundantSet -> findSingletons on the reference's bundled
    read sets; the dump (tokens, DRs, read order, orientation, start/stops, patterns) must match the reference's."""
    path = os.path.join(checkers.REF_DATA, name)
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged")
    want = gzip.open(os.path.join(G, "bundled", name + ".dump.gz")).read().decode("latin-1")
    res, max_len = ctx.run_files([path])
    assert res.dump(max_len) == want


def test_bundled_files_other_options(ctx, P):
    path = os.path.join(checkers.REF_DATA, "Ill100.fx.gz")
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged")
    for prm in (dict(window=6), dict(min_repeats=3), dict(window=9, low_dr=20), dict(window=7, low_spacer=20, high_spacer=60)):
        want, _ = P.run_files([path], prm)
        res, max_len = ctx.run_files([path], cb.Params(**prm))
        assert res.dump(max_len) == want, prm


# ---- seeded fuzz against the oracle -----------------------------------------------------------------------
def test_fuzz_default_params(ctx, P, k1path):
    rng = random.Random(101)
    reads = [fuzzgen.fuzz_read(rng, max_len=300) for _ in range(20000)]
    assert check_batch_against_oracle(ctx, P, reads) > 1500
    if k1path == "fast":
        assert 0 < ctx.last_candidates < len(reads)           # the filter really ran and really filtered


@pytest.mark.parametrize("max_len", [100, 105, 112, 150, 153, 160, 249, 256, 297, 304, 305])
def test_fast_path_length_buckets(ctx, P, max_len):
    """Every register-layout instantiation of the filter (and the hand-over to the generic kernel above 304 bp)."""
    rng = random.Random(1000 + max_len)
    reads = []
    for _ in range(3000):
        L = max_len if rng.random() < 0.5 else rng.randint(0, max_len)
        kind = rng.random()
        if L == 0:
            s = b""
        elif kind < 0.3:
            s = fuzzgen.rand_seq(rng, L)
        elif kind < 0.9:
            s = fuzzgen.planted_read(rng, L, sub_rate=rng.choice([0, 0, 0.01]))
        else:
            s = fuzzgen.microsat_read(rng, L)
        if rng.random() < 0.1:
            s = fuzzgen.mutate(rng, s, 0.01, b"NnacgtRY")
        reads.append(s)
    reads[-1] = fuzzgen.planted_read(rng, max_len)               # make sure the longest length is present
    assert check_batch_against_oracle(ctx, P, reads) > 100


def test_fuzz_other_params(ctx, P):
    rng = random.Random(102)
    for _ in range(12):
        prm = dict(window=rng.choice([6, 7, 8, 9]), min_repeats=rng.choice([2, 3, 4]), low_dr=rng.choice([23, 23, 20, 17, 30]),
                   high_dr=rng.choice([47, 40, 60]), low_spacer=rng.choice([26, 20, 30, 10]), high_spacer=rng.choice([50, 60, 40]))
        reads = [fuzzgen.fuzz_read(rng) for _ in range(1500)]
        check_batch_against_oracle(ctx, P, reads, prm)


def test_edge_batches(ctx, P, k1path):
    assert check_batch_against_oracle(ctx, P, []) == 0                        # empty batch
    assert check_batch_against_oracle(ctx, P, [b"", b"A", b"ACGT" * 14, b"", b"N" * 200]) == 0   # empty / short / below 58 bp
    rng = random.Random(103)
    ragged = [fuzzgen.planted_read(rng, rng.choice([58, 59, 60, 61, 100, 333, 1021])) for _ in range(300)]
    check_batch_against_oracle(ctx, P, ragged)


def test_long_reads(ctx, P, k1path):
    """BASELINE.json configs[2]: 1-10 kb reads (warp-per-read kernel on the fast path, generic kernel otherwise)."""
    rng = random.Random(104)
    reads = []
    for _ in range(400):
        L = rng.randint(1000, 10000)
        reads.append(fuzzgen.planted_read(rng, L, sub_rate=rng.choice([0, 0.005, 0.02])) if rng.random() < 0.6 else fuzzgen.rand_seq(rng, L))
    assert check_batch_against_oracle(ctx, P, reads) > 80


def test_long_reads_mixed_lengths_and_odd_content(ctx, P, k1path):
    """Ragged batch around the long-read kernel: empty / sub-58 / short / long reads, microsatellites (window grid keeps
    re-phasing), N and lower case, tandem arrays with many repeats."""
    rng = random.Random(107)
    reads = [b"", b"ACGT", fuzzgen.rand_seq(rng, 57), fuzzgen.rand_seq(rng, 58)]
    for _ in range(500):
        L = rng.choice([60, 150, 305, 306, 777, 1500, 3000, 6000])
        kind = rng.random()
        if kind < 0.25:
            s = fuzzgen.rand_seq(rng, L)
        elif kind < 0.75:
            s = fuzzgen.planted_read(rng, L, sub_rate=rng.choice([0, 0.01]))
        elif kind < 0.9:
            s = fuzzgen.microsat_read(rng, L)
        else:
            s = fuzzgen.planted_read(rng, L, spacer_lo=rng.randint(1, 30), spacer_hi=rng.randint(30, 60), jitter=rng.randint(0, 12))
        if rng.random() < 0.2:
            s = fuzzgen.mutate(rng, s, 0.01, b"NnacgtRY")
        reads.append(s)
    assert check_batch_against_oracle(ctx, P, reads) > 100


def test_mixed_batches_of_mostly_short_reads(ctx, P):
    """A few reads above 304 bases among thousands of short ones: the warp filter keeps everything up to 304 bases (reads of
    tiles that a long read blows up go to the exact kernel unfiltered) and lists the long ones for the warp-per-read kernel --
    one long read used to send the whole batch there.  Hits against the oracle read by read, and phase 2 on the 2-bit stream
    the mixed launch leaves behind against the byte-reading filter."""
    rng = random.Random(114)
    for n_long, long_lens in ((1, [305]), (25, [305, 320, 500, 1200, 4000]), (400, [310, 350, 400])):
        reads = []
        for _ in range(4000):
            L = rng.choice([0, 40, 100, 150, 150, 150, 250, 304])
            reads.append(fuzzgen.planted_read(rng, L, sub_rate=rng.choice([0, 0.01])) if L >= 100 and rng.random() < 0.4 else fuzzgen.rand_seq(rng, L))
        for _ in range(n_long):
            L = rng.choice(long_lens)
            s_ = fuzzgen.planted_read(rng, L) if rng.random() < 0.6 else fuzzgen.rand_seq(rng, L)
            reads.insert(rng.randint(0, len(reads)), s_)
        reads = [fuzzgen.mutate(rng, r, 0.01, b"Nacgt") if rng.random() < 0.1 else r for r in reads]
        assert sum(map(len, reads)) < 400 * len(reads)                    # mostly short: the mixed mode, not the long-read one
        for mixed in ("1", "0"):
            os.environ["CRASS_B200_K1_MIXED"] = mixed
            try:
                assert check_batch_against_oracle(ctx, P, reads) > 300
            finally:
                os.environ.pop("CRASS_B200_K1_MIXED", None)
        # phase 2 of the same (resident) batch: the kept 2-bit stream must be complete whatever tile wrote it
        bases, offs = cb.pack_reads(reads)
        ctx.upload(bases, offs)
        hits, pool, _ = ctx.dr_search_resident(cb.Params())
        pats = api.non_redundant_list(ctx.last_dr_list(), 6)
        ac = cb.Automaton(pats)
        h_packed = ctx.ac_scan_resident(ac, skip_found=False)
        os.environ["CRASS_B200_K2F"] = "bytes"
        try:
            h_bytes = ctx.ac_scan_resident(ac, skip_found=False)
        finally:
            os.environ.pop("CRASS_B200_K2F", None)
        assert hits_by_read(h_packed[0], h_packed[1]) == hits_by_read(h_bytes[0], h_bytes[1]) and len(h_packed[0]) > 300


def test_synthetic_config2_prefix(ctx, P, k1path):
    """A 300k-read prefix of the BASELINE config-2 recipe, compared read by read with the oracle."""
    genome, drs, _ = synth.make_genome(20242)
    n = 300_000
    bases, offs = synth.sample_fixed(genome, n, 150, 21242)
    hits, pool, found = ctx.dr_search(bases, offs)
    want = np.zeros(n, dtype=np.uint8)
    nf = P.lib.orc_phase1_batch(bases.ctypes.data, offs.ctypes.data, n, checkers.params_array(), want.ctypes.data)
    assert nf > 500
    assert np.array_equal(found, want)
    got = hits_by_read(hits, pool)
    for i in np.flatnonzero(want):
        s = bases[int(offs[i]):int(offs[i + 1])].tobytes()
        f, ss, rl = P.search_core(s)
        assert got[int(i)] == (ss, rl)


def test_full_size_config2_properties(ctx, P):
    """BASELINE.json configs[1] at its FULL size (10 M x 150 bp, the bench's recipe and seed): properties that do not need the
    oracle on every read.  (1) determinism: two runs give the same hit set; (2) shard additivity: the hits of the whole
    batch are the hits of its two halves and of an odd three-way cut (no result depends on what lies next to a read: tiles,
    look-ahead words, the kept 2-bit stream); (3) the read-ordered copy is a permutation of the slot-ordered hits, sorted
    and one per flagged read; (4) phase 2 never reports a read phase 1 flagged, and its answers on the kept 2-bit stream
    equal those of the byte-reading filter; (5) the oracle on 3 000 reads drawn from all over the batch agrees flag by flag
    and list by list."""
    import hashlib
    import torch
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    n, L, TOK = 10_000_000, 150, 64
    genome, _, _ = synth.make_genome(20242)
    d_bases, d_offsets = synth.sample_fixed_torch(genome, n, L, 20242 + 1000, dev)
    d_offsets = d_offsets.to(torch.int64)
    cap = n // 4 + 1024
    d_found = torch.empty(n, dtype=torch.uint8, device=dev)
    d_hits = torch.empty(cap * 4, dtype=torch.int32, device=dev)
    d_sorted = torch.empty(cap * 4, dtype=torch.int32, device=dev)
    d_pool = torch.empty(n + 4096, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    prm = cb.Params()
    ctx.keep_packed(True)

    def search(lo, hi):
        """hits of reads [lo, hi) as {global read index: (start/stop tuple, repeat length)} + the sorted copy's read indices"""
        m = hi - lo
        off = d_offsets[lo:hi + 1] - d_offsets[lo]
        bas = d_bases[int(d_offsets[lo]): int(d_offsets[hi])]
        ctx.dr_search_dev(bas, off, m, L, prm, d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
        ctx.sort_hits_dev(d_found, m, d_hits, d_cnt, cap, d_sorted, s.cuda_stream)
        s.synchronize()
        c = d_cnt.cpu().numpy()
        assert c[2] == 0
        nh = int(c[0])
        hits = d_sorted[: nh * 4].cpu().numpy().view(api.HIT_DTYPE)
        raw = d_hits[: nh * 4].cpu().numpy().view(api.HIT_DTYPE)
        pool = d_pool[: int(c[1])].cpu().numpy().view(np.uint32)
        flags = d_found[:m].cpu().numpy()
        # (3) sorted, one per flagged read, a permutation of the slot-ordered records
        assert np.all(np.diff(hits["read_index"].astype(np.int64)) > 0)
        assert np.array_equal(hits["read_index"], np.flatnonzero(flags).astype(np.uint32))
        assert np.array_equal(np.sort(raw, order=["read_index"]), hits)
        out = {int(h["read_index"]) + lo: (tuple(int(x) for x in pool[h["ss_offset"]: h["ss_offset"] + h["n_ss"]]), int(h["repeat_len"])) for h in hits}
        return out, flags

    whole, flags = search(0, n)
    assert len(whole) > 50_000
    digest = hashlib.md5(repr(sorted(whole.items())).encode()).hexdigest()
    again, _ = search(0, n)
    assert hashlib.md5(repr(sorted(again.items())).encode()).hexdigest() == digest                 # (1)
    for cuts in ((0, n // 2, n), (0, 3_333_337, 6_000_001, n)):                                        # (2)
        parts = {}
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            parts.update(search(lo, hi)[0])
        assert parts == whole, cuts
    # (5) the oracle on reads from all over the batch
    rng = np.random.default_rng(7)
    pick = np.unique(np.concatenate([rng.integers(0, n, 2500), np.flatnonzero(flags)[:: max(1, len(whole) // 500)]]))
    h_reads = d_bases.view(n, L)[torch.from_numpy(pick).to(dev)].cpu().numpy()
    for i, row in zip(pick, h_reads):
        f, ss, rl = P.search_core(row.tobytes())
        assert bool(f) == bool(flags[i])
        if f:
            assert whole[int(i)] == (tuple(ss), rl)
    # (4) phase 2 on the whole batch: the matcher from the batch's own DR variants
    whole2, flags = search(0, n)                                                                       # d_found / the 2-bit stream of the WHOLE batch again
    d_tok = torch.empty(cap * TOK, dtype=torch.uint8, device=dev)
    ctx.set_token_output(d_tok, TOK)
    ctx.dr_search_dev(d_bases, d_offsets, n, L, prm, d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
    ctx.set_token_output(None)
    nh = int(d_cnt.cpu()[0])
    blk = torch.empty(api.token_block_bytes(16384, TOK), dtype=torch.uint8, device=dev)
    ctx.unique_tokens_block_dev(d_hits, nh, d_tok, TOK, blk, 16384, s.cuda_stream)
    ac, cnt, fl = ctx.cluster_block_dev(blk, 16384, TOK, 6, s.cuda_stream)
    assert fl == 0 and ac is not None and ac.num_patterns > 500
    d_found2 = torch.empty(n, dtype=torch.uint8, device=dev)
    answers = []
    for mode in (None, "bytes"):
        if mode:
            os.environ["CRASS_B200_K2F"] = mode
        try:
            ctx.ac_scan_dev(ac, d_bases, d_offsets, n, L, d_found, d_found2, d_hits, d_pool, d_cnt, s.cuda_stream)
            s.synchronize()
        finally:
            os.environ.pop("CRASS_B200_K2F", None)
        c = d_cnt.cpu().numpy()
        h2 = np.sort(d_hits[: int(c[0]) * 4].cpu().numpy().view(api.HIT_DTYPE), order=["read_index"])
        p2 = d_pool[: int(c[1])].cpu().numpy().view(np.uint32)
        f2 = d_found2.cpu().numpy()
        assert not np.any(f2 & d_found.cpu().numpy())                      # never a read phase 1 has
        answers.append((h2["read_index"].copy(), np.stack([p2[h2["ss_offset"]], p2[h2["ss_offset"] + 1]], axis=1), f2))
    assert len(answers[0][0]) > 20_000
    assert all(np.array_equal(a, b) for a, b in zip(answers[0], answers[1]))
    ctx.keep_packed(False)


def test_device_dr_tokens_match_host_lowlexi(ctx, k1path):
    """K4: the low-lexi DR token written next to every hit on the device == ReadHolder::DRLowLexi replayed on the host
    (which test_host_logic pins against the reference), including first-appearance order of the distinct tokens."""
    rng = random.Random(106)
    reads = [fuzzgen.planted_read(rng, rng.choice([100, 150, 150, 250]), sub_rate=rng.choice([0, 0.01])) for _ in range(6000)]
    reads += [fuzzgen.mutate(rng, r, 0.01, b"NnacgtRYU") for r in reads[:1500]]
    bases, offs = cb.pack_reads(reads)
    ctx.upload(bases, offs)
    hits, pool, _ = ctx.dr_search_resident(cb.Params())
    assert len(hits) > 2000
    want = api.dr_list_from_hits(bases, offs, hits, pool)
    assert ctx.last_dr_list() == want
    batch = cb.Batch.from_arrays(bases, offs)
    res = cb.Results()
    res.add_phase1(batch, hits, pool)
    assert res.dr_list() == want


def test_token_blocks_of_three_shards_merge_like_one_sequential_run(ctx):
    """K4b in block form + K4c: three contiguous shards of unequal size, searched one after the other on this GPU, are
    merged on the device exactly as the N-rank exchange does after its all-gather; the merged list must be the token
    order of one sequential run over the whole batch (StringCheck numbers by first appearance)."""
    import torch
    rng = random.Random(107)
    pool_drs = [fuzzgen.rand_seq(rng, rng.randint(24, 40)) for _ in range(12)]
    reads = [fuzzgen.planted_read(rng, rng.choice([100, 150, 150, 250]), dr=rng.choice(pool_drs), sub_rate=rng.choice([0, 0.01])) for _ in range(9000)]
    bases, offs = cb.pack_reads(reads)
    ctx.upload(bases, offs)
    hits, pool, _ = ctx.dr_search_resident(cb.Params())
    want = api.dr_list_from_hits(bases, offs, hits, pool)
    assert len(want) > 100
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    TOK, cuts = 64, [0, 2500, 6100, len(reads)]
    shard_reads = max(b - a for a, b in zip(cuts, cuts[1:]))
    o64 = offs.astype(np.int64)

    def shard_blocks(cap):
        blocks = []
        with torch.cuda.stream(s):
            for lo, hi in zip(cuts, cuts[1:]):
                n = hi - lo
                d_b = torch.from_numpy(bases[int(o64[lo]):int(o64[hi])].copy()).to(dev)
                d_o = torch.from_numpy(o64[lo:hi + 1] - o64[lo]).to(dev)
                d_found = torch.empty(n, dtype=torch.uint8, device=dev)
                d_hits = torch.empty((n + 16) * 4, dtype=torch.int32, device=dev)
                d_pool = torch.empty(64 * n, dtype=torch.int32, device=dev)
                d_cnt = torch.zeros(8, dtype=torch.int32, device=dev)
                d_tok = torch.empty((n + 16) * TOK, dtype=torch.uint8, device=dev)
                ctx.set_token_output(d_tok, TOK)
                ctx.dr_search_dev(d_b, d_o, n, 256, cb.Params(), d_found, d_hits, d_pool, d_cnt, s.cuda_stream)
                ctx.set_token_output(None)
                nh = int(d_cnt.cpu()[0])
                blk = torch.empty(api.token_block_bytes(cap, TOK), dtype=torch.uint8, device=dev)
                ctx.unique_tokens_block_dev(d_hits, nh, d_tok, TOK, blk, cap, s.cuda_stream)
                s.synchronize()
                # one shard on its own: the block is that shard's list
                sh = hits[(hits["read_index"] >= lo) & (hits["read_index"] < hi)].copy()
                sh["read_index"] -= lo
                text, count, flags = api.dr_list_from_block(blk.cpu().numpy(), cap, TOK)
                local = api.dr_list_from_hits(bases[int(o64[lo]):int(o64[hi])].copy(), (o64[lo:hi + 1] - o64[lo]).astype(np.uint64), sh, pool)
                assert count == len(local) and (count > cap or text == b"".join(d + b"\n" for d in local))
                blocks.append(blk)
        return blocks

    cap = 4096
    blocks = shard_blocks(cap)
    with torch.cuda.stream(s):
        out = torch.empty(api.token_block_bytes(3 * cap, TOK), dtype=torch.uint8, device=dev)
        ctx.merge_token_blocks_dev(torch.cat(blocks), 3, cap, TOK, shard_reads, out, 3 * cap, s.cuda_stream)
        s.synchronize()
    text, count, flags = api.dr_list_from_block(out.cpu().numpy(), 3 * cap, TOK)
    assert flags == 0 and count == len(want)
    assert text == b"".join(d + b"\n" for d in want)
    # blocks that are too small say so instead of dropping tokens silently
    blocks = shard_blocks(8)
    with torch.cuda.stream(s):
        ctx.merge_token_blocks_dev(torch.cat(blocks), 3, 8, TOK, shard_reads, out, 3 * cap, s.cuda_stream)
        s.synchronize()
    assert api.dr_list_from_block(out.cpu().numpy(), 3 * cap, TOK)[2] & 1
    with torch.cuda.stream(s):
        ctx.merge_token_blocks_dev(torch.cat(shard_blocks(cap)), 3, cap, TOK, shard_reads, out, 16, s.cuda_stream)
        s.synchronize()
    assert api.dr_list_from_block(out.cpu().numpy(), 16, TOK)[1] == len(want)


def _token_block(uniq, keys, slots, cap, stride):
    blk = np.zeros(api.token_block_bytes(cap, stride), dtype=np.uint8)
    blk[:4] = np.frombuffer(np.uint32(len(uniq)).tobytes(), dtype=np.uint8)
    for t, slot in enumerate(slots):
        rec = blk[16 + slot * stride: 16 + (slot + 1) * stride]
        rec[0] = len(uniq[t])
        rec[2:2 + len(uniq[t])] = np.frombuffer(uniq[t], dtype=np.uint8)
        rec[stride - 4:] = np.frombuffer(np.uint32(keys[t]).tobytes(), dtype=np.uint8)
    return blk


def test_clustering_passes_on_the_device_match_the_host(ctx, P):
    """K5: createNonRedundantSet as kernels on a token block (token order, 11-mer keys and first holders, the group walk,
    the substring reduction, survivors + reverse complements) must give exactly the pattern set (same strings, same order)
    of the host passes and of the oracle, incl. N/R letters (string-keyed k-mers), 'U' (declined: host passes), DRs shorter
    than a k-mer, tiny and long lists -- and so must the round-1 splits between device and host."""
    import torch
    rng = random.Random(109)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    dev = torch.device("cuda", 0)
    for n_base, n_var, alphabet, lo, hi in ((1, 1, b"ACGT", 23, 47), (3, 4, b"ACGT", 23, 47), (40, 12, b"ACGTN", 23, 47), (60, 10, b"ACGTUR", 23, 47),
                                           (30, 10, b"ACGTN", 6, 20), (25, 14, b"ACUR", 11, 30), (400, 16, b"ACGTN", 23, 47), (1500, 28, b"ACGT", 23, 47),
                                           (50, 40, b"ACGTNRY", 23, 47), (60, 30, b"ACGTN", 40, 110), (30, 30, b"ACGT", 50, 64),
                                           (2200, 16, b"ACGT", 23, 47)):   # the last one is past the device limit (32768): host passes
        stride = 64 if hi <= 47 else 128 if hi > 64 else 68            # wider DR bounds: longer tokens (pass D on the bytes instead of the 2-bit codes)
        base = [fuzzgen.rand_seq(rng, rng.randint(lo, hi)) for _k in range(n_base)]
        drs = []
        for b in base:
            for _k in range(n_var):
                v = fuzzgen.mutate(rng, b, rng.choice([0, 0.02, 0.05]), alphabet)
                a, e = rng.randint(0, 4), rng.randint(0, 4)
                v = v[a:len(v) - e] if rng.random() < 0.5 else fuzzgen.rand_seq(rng, a) + v + fuzzgen.rand_seq(rng, e)
                drs.append(min(v, v.translate(comp)[::-1])[:stride - 6])
        uniq = [u for u in dict.fromkeys(drs) if u]
        rng.shuffle(uniq)                                                 # token order = order of `uniq`
        keys = sorted(rng.sample(range(50_000_000), len(uniq)))
        slots = list(range(len(uniq)))
        rng.shuffle(slots)                                                # records sit in the block in arbitrary order
        cap = len(uniq) + rng.randint(0, 50)
        blk = _token_block(uniq, keys, slots, cap, stride)
        want, cnt_h, fl_h = api.non_redundant_patterns_from_block(blk, cap, stride, 6)
        assert want == api.non_redundant_patterns(b"".join(d + b"\n" for d in uniq), 6)
        d_blk = torch.from_numpy(blk).to(dev)
        before = ctx.launch_count
        got, cnt, fl = ctx.cluster_block_patterns_dev(d_blk, cap, stride, 6)
        assert (cnt, fl) == (cnt_h, fl_h) == (len(uniq), 0)
        assert got == want, (n_base, n_var, alphabet, lo, hi)
        if (alphabet == b"ACGT" or len(uniq) < 2000) and b"U" not in alphabet and len(uniq) <= 32768:
            assert ctx.launch_count - before == (24 if stride <= 70 else 21)                        # really the kernels, not a quiet detour over the host (which lists with more than 16384 string-keyed k-mers take)
        for mode in ("device-passes", "device-reduce"):                  # the round-1 splits: first passes (and pass D) on the device, rest on the host
            os.environ["CRASS_B200_CLUSTER"] = mode
            try:
                assert ctx.cluster_block_patterns_dev(d_blk, cap, stride, 6)[0] == want
            finally:
                os.environ.pop("CRASS_B200_CLUSTER", None)
        ac, cnt2, fl2 = ctx.cluster_block_dev(d_blk, cap, stride, 6)
        assert ac is not None and ac.num_patterns == want.count(b"\n") and cnt2 == len(uniq)
        assert ac.pattern_text() == want
        if len(uniq) < 3000:                                              # the oracle's clustering is quadratic
            ref = P.non_redundant(uniq)
            assert sorted(l[2:] for l in ref.split("\n") if l.startswith("P\t")) == sorted(want.decode().split("\n")[:-1])
    # an overflowed block is reported, not clustered
    blk[:4] = np.frombuffer(np.uint32(cap + 5).tobytes(), dtype=np.uint8)
    got, cnt, fl = ctx.cluster_block_patterns_dev(torch.from_numpy(blk).to(dev), cap, stride, 6)
    assert got == b"" and cnt == cap + 5


def test_group_walk_on_the_device_follows_long_dependency_chains(ctx, P):
    """The order-dependent walk of clusterDRReads (WorkHorse.cpp:1542-1625) runs as a dependency graph on the device: DR t
    waits for the groups of the earlier DRs its k-mers were first seen in.  Windows sliding over one long sequence make
    that graph a single chain as long as the list (every DR hangs on the one before it), the worst case for it; windows
    over several sequences in interleaved order give many chains at once."""
    import torch
    rng = random.Random(110)
    dev = torch.device("cuda", 0)
    stride = 64
    for n_seq, n_win, step, shuffle in ((1, 3000, 5, False), (1, 800, 9, False), (7, 400, 3, True), (1, 600, 1, False)):
        uniq = []
        seqs = [fuzzgen.rand_seq(rng, n_win * step + 60) for _ in range(n_seq)]
        for w in range(n_win):
            for sq in seqs:
                uniq.append(sq[w * step: w * step + rng.randint(30, 44)])
        uniq = list(dict.fromkeys(uniq))
        if shuffle:
            # keep every chain's order, interleave the chains at random
            chains = [uniq[i::n_seq] for i in range(n_seq)]
            uniq = []
            while any(chains):
                c = rng.choice([c for c in chains if c])
                uniq.append(c.pop(0))
        keys = sorted(rng.sample(range(50_000_000), len(uniq)))
        slots = list(range(len(uniq)))
        rng.shuffle(slots)
        cap = len(uniq) + 7
        blk = _token_block(uniq, keys, slots, cap, stride)
        want, _, _ = api.non_redundant_patterns_from_block(blk, cap, stride, 6)
        before = ctx.launch_count
        got, cnt, fl = ctx.cluster_block_patterns_dev(torch.from_numpy(blk).to(dev), cap, stride, 6)
        assert ctx.launch_count - before == 24 and (cnt, fl) == (len(uniq), 0)
        assert got == want, (n_seq, n_win, step)


def test_singleton_scan_on_the_kept_2bit_stream(ctx, P):
    """Phase 2 of a resident batch reads the 2-bit stream phase 1's filter left in HBM (k_ac_filter_packed); the hits must be
    those of the byte-reading filter and of the oracle, for every read-length bucket and with odd bytes in the reads."""
    rng = random.Random(108)
    for max_len in (100, 150, 250, 300, 1500):                            # 1500: the warp-per-read kernels of the long path
        pool_drs = [fuzzgen.rand_seq(rng, rng.randint(24, 40)) for _ in range(8)]
        reads = [fuzzgen.planted_read(rng, rng.randint(max(60, max_len - 50), max_len), dr=rng.choice(pool_drs), sub_rate=rng.choice([0, 0.01]))
                 for _ in range(3000)]
        reads += [fuzzgen.mutate(rng, r, 0.01, b"NnacgtRYU") for r in reads[:800]]
        reads += [fuzzgen.rand_seq(rng, rng.randint(0, max_len)) for _ in range(1200)]
        rng.shuffle(reads)
        bases, offs = cb.pack_reads(reads)
        ctx.upload(bases, offs)
        hits, pool, _ = ctx.dr_search_resident(cb.Params())
        pats = api.non_redundant_list(ctx.last_dr_list(), 6)
        assert len(pats) >= 8
        if max_len == 250:                                                # enough 16-mers for the 128 KB bitmap (one CTA per SM)
            pats += fuzzgen.dr_like_patterns(rng, 7000)
        ac = cb.Automaton(pats)
        got = {}
        for mode in ("packed", "bytes"):
            if mode == "bytes":
                os.environ["CRASS_B200_K2F"] = "bytes"
            try:
                h2, p2, f2 = ctx.ac_scan_resident(ac, skip_found=True, want_found=True)
            finally:
                os.environ.pop("CRASS_B200_K2F", None)
            got[mode] = (hits_by_read(h2, p2), f2.tobytes())
        assert got["packed"] == got["bytes"]
        by_read = got["packed"][0]
        skip = set(hits["read_index"].tolist())
        h = P.ac_create(pats)
        for i in range(0, len(reads), 7):                                 # oracle on a sample
            m = None if i in skip else P.ac_first_match(h, reads[i])
            assert (m is None) == (i not in by_read)
        P.ac_destroy(h)
        assert len(h2) > 100


def test_matcher_tables_built_on_the_device(ctx, P):
    """k_ac_build fills bitmap, key table and start table with atomics on the device; CRASS_B200_AC_BUILD=host fills them
    sequentially on the host and uploads them.  Both must answer every read like the oracle's automaton -- also for
    patterns with long G runs, whose all-ones 16-mer is the tables' empty marker and goes through the side words."""
    rng = random.Random(111)
    for n_pat in (1, 60, 3000):
        pats = fuzzgen.dr_like_patterns(rng, n_pat)
        pats += [b"G" * rng.randint(23, 40) for _ in range(2)]                                   # every window is all ones
        pats += [fuzzgen.rand_seq(rng, rng.randint(0, 7)) + b"G" * 16 + fuzzgen.rand_seq(rng, rng.randint(7, 20)) for _ in range(6)]   # one all-ones window at offset 0..7
        pats += [b"G" * 16 + fuzzgen.rand_seq(rng, rng.randint(8, 20)) for _ in range(3)]       # ... chained from the side word
        pats += [pats[0][:16] + fuzzgen.rand_seq(rng, 12), pats[0][:16] + fuzzgen.rand_seq(rng, 20)]   # shared first 16-mer: one start-table chain
        pats = [p for p in dict.fromkeys(pats) if len(p) >= 23]
        texts = []
        for _ in range(3000):
            t = bytearray(fuzzgen.rand_seq(rng, rng.choice([30, 100, 150, 150, 300, 900])))
            r = rng.random()
            if r < 0.5:
                p = rng.choice(pats)
                pos = rng.randint(0, len(t) - 1)
                t = (t[:pos] + p + t[pos:])[:len(t)]
            elif r < 0.6:
                pos = rng.randint(0, len(t) - 1)
                t = (t[:pos] + b"G" * rng.randint(10, 60) + t[pos:])[:len(t)]
            texts.append(bytes(t))
        bases, offs = cb.pack_reads(texts)
        h = P.ac_create(pats)
        want = {}
        for i, t in enumerate(texts):
            m = P.ac_first_match(h, t)
            if m is not None:
                dr_end = min(m[0] - 1, len(t) - 1)
                want[i] = ([dr_end - (m[1] - 1), dr_end], 0)
        P.ac_destroy(h)
        assert len(want) > 1000
        for build in ("device", "host"):
            os.environ["CRASS_B200_AC_BUILD"] = build
            try:
                for short_only in (False, True):                            # long reads take the warp filter, reads <= 304 the tile filters
                    sel = [i for i, t in enumerate(texts) if len(t) <= 304] if short_only else list(range(len(texts)))
                    b2, o2 = cb.pack_reads([texts[i] for i in sel])
                    hits, pool, found = ctx.ac_scan(cb.Automaton(pats), b2, o2)
                    got = hits_by_read(hits, pool)
                    assert got == {k: want[i] for k, i in enumerate(sel) if i in want}, (n_pat, build, short_only)
            finally:
                os.environ.pop("CRASS_B200_AC_BUILD", None)


def test_singleton_scan_fuzz(ctx, P, k2path):
    rng = random.Random(105)
    for n_pat in (1, 7, 100, 1500, 12000):
        pats = fuzzgen.dr_like_patterns(rng, n_pat)
        if n_pat == 7:
            pats += [p[2:-3] for p in pats] + [fuzzgen.mutate(rng, p, 0.1, b"ACGTN") for p in pats]
        texts = []
        for _ in range(3000):
            t = fuzzgen.rand_seq(rng, rng.choice([0, 10, 100, 150, 150, 400]), b"ACGTN" if rng.random() < 0.2 else b"ACGT")
            if rng.random() < 0.5 and len(t) > 60:
                p = rng.choice(pats)
                pos = rng.randint(0, len(t) - 1)
                t = (t[:pos] + p + t[pos:])[:len(t)]
            texts.append(t)
        skip = np.array([rng.random() < 0.1 for _ in texts], dtype=np.uint8)
        bases, offs = cb.pack_reads(texts)
        hits, pool, found = ctx.ac_scan(cb.Automaton(pats), bases, offs, skip)
        got = hits_by_read(hits, pool)
        h = P.ac_create(pats)
        for i, t in enumerate(texts):
            m = None if skip[i] else P.ac_first_match(h, t)
            if m is None:
                assert i not in got and found[i] == 0
            else:
                dr_end = min(m[0] - 1, len(t) - 1)
                assert got[i] == ([dr_end - (m[1] - 1), dr_end], 0)
        P.ac_destroy(h)


def test_singleton_scan_short_reads_overlapping_patterns(ctx, P, k2path):
    """Reads up to 304 bp (the filter + verify kernels of the fast path, bytes form and 2-bit-stream form): several planted
    occurrences per read, patterns nested in each other and overlapping, occurrences cut by the read end, N runs -- the
    answer is acism's first callback: earliest end, longest pattern on ties."""
    rng = random.Random(106)
    for n_pat in (3, 40, 2500):
        pats = fuzzgen.dr_like_patterns(rng, n_pat)
        pats += [p[rng.randint(0, 4): len(p) - rng.randint(0, 4)] for p in pats[: max(2, n_pat // 3)] if len(p) >= 31]     # nested, still >= 23
        pats += [fuzzgen.mutate(rng, p, 0.05, b"ACGT") for p in pats[: max(2, n_pat // 3)]]
        pats = [p for p in dict.fromkeys(pats) if len(p) >= 23]
        texts = []
        for _ in range(4000):
            L = rng.choice([0, 15, 23, 40, 100, 150, 150, 151, 250, 304])
            t = bytearray(fuzzgen.rand_seq(rng, L, b"ACGTN" if rng.random() < 0.2 else b"ACGT"))
            for _ in range(rng.choice([0, 1, 1, 2, 3])):
                if L < 23:
                    break
                p = rng.choice(pats)
                pos = rng.randint(-5, L - 10)
                piece = p[max(0, -pos):]
                pos = max(pos, 0)
                piece = piece[: L - pos]
                t[pos: pos + len(piece)] = piece
            texts.append(bytes(t))
        bases, offs = cb.pack_reads(texts)
        ac = cb.Automaton(pats)
        h = P.ac_create(pats)
        want = [P.ac_first_match(h, t) for t in texts]
        P.ac_destroy(h)
        assert sum(m is not None for m in want) > 1000
        # bytes form
        hits, pool, found = ctx.ac_scan(ac, bases, offs, None)
        got = hits_by_read(hits, pool)
        # 2-bit-stream form: phase 1 of the resident batch leaves the stream behind, phase 2 reads it
        ctx.upload(bases, offs)
        _, _, f1 = ctx.dr_search_resident(want_found=True)
        h2, p2, f2 = ctx.ac_scan_resident(ac, skip_found=True, want_found=True)
        got2 = hits_by_read(h2, p2)
        for i, (t, m) in enumerate(zip(texts, want)):
            exp = None
            if m is not None:
                dr_end = min(m[0] - 1, len(t) - 1)
                exp = ([dr_end - (m[1] - 1), dr_end], 0)
            assert got.get(i) == exp and found[i] == (exp is not None)
            if f1[i]:
                assert i not in got2
            else:
                assert got2.get(i) == exp and f2[i] == (exp is not None)


def test_singleton_scan_long_reads(ctx, P, k2path):
    """K2 on 0.3-8 kb reads (warp-per-read q-gram filter + verify, or the generic automaton walk)."""
    rng = random.Random(108)
    pats = fuzzgen.dr_like_patterns(rng, 200)
    texts = []
    for _ in range(600):
        t = fuzzgen.rand_seq(rng, rng.choice([15, 16, 305, 400, 1000, 3000, 8000]), b"ACGTN" if rng.random() < 0.2 else b"ACGT")
        for _k in range(rng.choice([0, 0, 1, 1, 3])):
            if len(t) > 60:
                p = rng.choice(pats)
                if rng.random() < 0.3:
                    p = p[:-1]                                             # near miss
                pos = rng.randint(0, len(t) - 1)
                t = (t[:pos] + p + t[pos:])[:len(t)]
        texts.append(t)
    skip = np.array([rng.random() < 0.1 for _ in texts], dtype=np.uint8)
    bases, offs = cb.pack_reads(texts)
    hits, pool, found = ctx.ac_scan(cb.Automaton(pats), bases, offs, skip)
    got = hits_by_read(hits, pool)
    h = P.ac_create(pats)
    n_match = 0
    for i, t in enumerate(texts):
        m = None if skip[i] else P.ac_first_match(h, t)
        if m is None:
            assert i not in got and found[i] == 0
        else:
            n_match += 1
            dr_end = min(m[0] - 1, len(t) - 1)
            assert got[i] == ([dr_end - (m[1] - 1), dr_end], 0)
    P.ac_destroy(h)
    assert n_match > 150


def test_whole_path_synthetic_against_oracle(ctx, P, tmp_path):
    """Both phases + clustering on a synthetic FASTA file: the product's dump must equal the oracle's."""
    genome, drs, _ = synth.make_genome(777, n_dr_types=12, array_fraction=0.05)
    bases, offs = synth.sample_fixed(genome, 60000, 150, 778)
    path = str(tmp_path / "synth.fa")
    with open(path, "wb") as fh:
        for i in range(60000):
            fh.write(b">r%010d\n" % i + bases[int(offs[i]):int(offs[i + 1])].tobytes() + b"\n")
    want, _ = P.run_files([path])
    res, max_len = ctx.run_files([path])
    got = res.dump(max_len)
    assert got == want
    assert sum(1 for l in got.split("\n") if l.startswith("R\t")) > 1000


# ---- K6: partial-DR recovery (ReadHolder::updateStartStops + smithWaterman) ---------------------------------------
@pytest.fixture(params=["warp", "thread"])
def k6path(request):
    """K6 has two kernels with identical results: one warp per read (lane wavefront, the default) and one thread per
    read (the source tests/hostsim also compiles)."""
    old = os.environ.get("CRASS_B200_K6")
    os.environ["CRASS_B200_K6"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("CRASS_B200_K6", None)
    else:
        os.environ["CRASS_B200_K6"] = old


def test_update_start_stops_golden_vectors(ctx, k6path):
    """The reference's own outputs (tests/golden/make_golden.py uss), all jobs in one launch, several DRs."""
    vec = json.load(open(os.path.join(G, "update_start_stops_vectors.json")))["update_start_stops"]
    reads = [v["seq"].encode() for v in vec]
    bases, offsets = api.pack_reads(reads)
    by_low = {}
    for i, v in enumerate(vec):
        by_low.setdefault(v["low_spacer"], []).append(i)
    for low, idx in by_low.items():
        drs = [vec[i]["dr"].encode() for i in idx]
        jobs = [(i, vec[i]["ss"], vec[i]["front"], k) for k, i in enumerate(idx)]
        got = ctx.update_start_stops(bases, offsets, drs, jobs, low)
        for (st, out), i in zip(got, idx):
            assert (st, out) == (0, vec[i]["ss_out"])


def test_update_start_stops_fuzz_against_oracle(ctx, k6path):
    P = checkers.port()
    rng = random.Random(23)
    cases = [fuzzgen.uss_case(rng) for _ in range(6000)]
    # long reads and long DRs as well: flanks of a few thousand bases, the DR at the 127-byte limit
    for _ in range(40):
        dr = fuzzgen.rand_seq(rng, rng.randint(60, 127))
        lead = fuzzgen.rand_seq(rng, rng.randint(100, 3000)) + dr[len(dr) - rng.randint(5, 60):]
        body = fuzzgen.rand_seq(rng, 40) + dr + fuzzgen.rand_seq(rng, 35) + dr
        tail = fuzzgen.rand_seq(rng, rng.randint(30, 3000)) + dr[: rng.randint(5, 60)]
        seq = lead + body + tail
        s0 = len(lead) + 40
        cases.append((seq, [s0 + 3, s0 + len(dr) - 2, s0 + len(dr) + 35 + 3, s0 + 2 * len(dr) + 35 - 2], 3, dr))
    # bytes outside A/C/G/T/N (row-DP fallback of the similarity test) and every columns-per-lane bucket of the warp kernel
    for _ in range(300):
        seq, ss, front, dr = fuzzgen.uss_case(rng)
        cases.append((seq.replace(b"G", b"g", 3).replace(b"A", b"R", 2), ss, front, dr.replace(b"G", b"g", 1) if rng.random() < 0.5 else dr))
    for _ in range(300):
        dr = fuzzgen.rand_seq(rng, rng.randint(30, 100))
        sp = [fuzzgen.rand_seq(rng, rng.randint(26, 50)) for _ in range(3)]
        lead = fuzzgen.mutate(rng, dr[len(dr) - rng.randint(4, len(dr)):], 0.04, b"ACGT")
        tail = fuzzgen.mutate(rng, dr[: rng.randint(4, len(dr))], 0.04, b"ACGT")
        seq = lead + sp[0] + dr + sp[1] + dr + sp[2] + tail
        s0 = len(lead) + len(sp[0])
        s1 = s0 + len(dr) + len(sp[1])
        cases.append((seq, [s0 + 2, s0 + len(dr) - 3, s1 + 2, s1 + len(dr) - 3], 2, dr))
    reads = [c[0] for c in cases]
    bases, offsets = api.pack_reads(reads)
    for low in (26, 20, 35):
        drs = [c[3] for c in cases]
        jobs = [(i, c[1], c[2], i) for i, c in enumerate(cases)]
        got = ctx.update_start_stops(bases, offsets, drs, jobs, low)
        grew = past = 0
        for (st, out), c in zip(got, cases):
            want = P.update_start_stops(c[0], c[1], c[2], c[3], low)
            if want[0] == -3:
                assert (st, out) == (3, [])
                past += 1
                continue
            assert (st, out) == (0, want[1])
            grew += len(out) > len(c[1])
        assert grew > 2000


def test_update_start_stops_status_codes_and_validation(ctx, k6path):
    bases, offsets = api.pack_reads([b"ACGT" * 40])
    ok = ctx.update_start_stops(bases, offsets, [b"ACGTACGTACGTACGTACGTACGTAC"], [(0, [40, 60], 0, 0)])
    assert ok[0][0] == 0 and len(ok[0][1]) >= 2
    assert ctx.update_start_stops(bases, offsets, [b"ACGTACGTACGTACGTACGTACGTAC"], [(0, [40], 0, 0)])[0] == (1, [])
    assert ctx.update_start_stops(bases, offsets, [b"A" * 128], [(0, [40, 60], 0, 0)])[0] == (2, [])
    assert ctx.update_start_stops(bases, offsets, [b"ACGTACGTACGTACGTACGTACGTAC"], [(0, [40, 60], -200, 0)])[0] == (3, [])
    with pytest.raises(cb.CrassB200Error):
        ctx.update_start_stops(bases, offsets, [b"ACGTACGTACGTACGTACGTACGTAC"], [(1, [40, 60], 0, 0)])      # no such read
    with pytest.raises(cb.CrassB200Error):
        ctx.update_start_stops(bases, offsets, [b"ACGTACGTACGTACGTACGTACGTAC"], [(0, [40, 60], 0, 5)])      # no such DR
    assert ctx.update_start_stops(bases, offsets, [b"ACGT"], []) == []


def test_update_start_stops_on_the_hits_of_a_bundled_read_set(ctx, k6path):
    """K1's own start/stop lists as input: every phase-1 hit of Ill100.fx.gz against its own DR token (front offset 0
    and a few shifted ones), checked against the oracle."""
    path = os.path.join(checkers.REF_DATA, "Ill100.fx.gz")
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged")
    P = checkers.port()
    batch = cb.Batch.from_file(path)
    bases, offsets = batch.bases, batch.offsets
    hits, pool, _ = ctx.dr_search(bases, offsets)
    assert len(hits) > 500
    rng = random.Random(2)
    drs, jobs = [], []
    for h in hits:
        r, o, n = int(h["read_index"]), int(h["ss_offset"]), int(h["n_ss"])
        seq = bytes(bases[int(offsets[r]): int(offsets[r + 1])])
        ss = [int(x) for x in pool[o: o + n]]
        k = 2 if n >= 4 else 0
        drs.append(seq[ss[k]: ss[k + 1] + 1])
        jobs.append((r, ss, rng.choice([0, 0, 1, 3, -2]), len(drs) - 1))
    got = ctx.update_start_stops(bases, offsets, drs, jobs)
    for (st, out), (r, ss, front, d) in zip(got, jobs):
        seq = bytes(bases[int(offsets[r]): int(offsets[r + 1])])
        want = P.update_start_stops(seq, ss, front, drs[d])
        assert (st, out) == ((3, []) if want[0] == -3 else (0, want[1]))


# ---- K7: consensus DR of DR groups (ksw_align + Aligner) --------------------------------------------------------------------

def test_ksw_align_golden_vectors_and_fuzz(ctx, P):
    """k_ksw_align (eight threads per alignment, one per lane of the reference's striped SSE2 kernel) against the vectors the
    compiled reference made and against the oracle on fresh cases: score, both end points, both start points."""
    g = json.load(open(os.path.join(G, "consensus_vectors.json")))
    got = ctx.ksw_align([(v["q"].encode(), v["t"].encode(), False) for v in g["ksw_align"]])
    for v, r in zip(g["ksw_align"], got):
        assert list(r) == v["out"], (v["q"], v["t"])
    rng = random.Random(112)
    nt = bytes.maketrans(b"ACGTN", bytes([0, 1, 2, 3, 4]))
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    for xtra in (0x80000 | 0x40000 | 5, 0x80000, 0x40000 | 12, 0):
        pairs = []
        for _ in range(1500):
            t = fuzzgen.rand_seq(rng, rng.randint(8, 150))
            if rng.random() < 0.75:
                a, b = sorted(rng.sample(range(len(t)), 2))
                q = fuzzgen.mutate(rng, t[a:b + 1][:120], rng.choice([0, 0.05, 0.2]), b"ACGTN")
                if rng.random() < 0.3 and len(q) > 6:
                    k = rng.randint(1, len(q) - 2)
                    q = q[:k] + fuzzgen.rand_seq(rng, rng.randint(1, 3)) + q[k:]
                if rng.random() < 0.3 and len(q) > 8:
                    k = rng.randint(1, len(q) - 4)
                    q = q[:k] + q[k + rng.randint(1, 3):]
            else:
                q = fuzzgen.rand_seq(rng, rng.randint(1, 128))
            if q:
                pairs.append((q, t, rng.random() < 0.4))
        got = ctx.ksw_align(pairs, xtra)
        for (q, t, rc), r in zip(pairs, got):
            qq = q.translate(comp)[::-1] if rc else q
            assert r == P.ksw_align(qq.translate(nt), t.translate(nt), xtra), (q, t, rc, xtra)


def test_consensus_groups_golden_vectors_and_fuzz(ctx, P):
    """K7 on whole DR groups: placements, strands, the extendSlaveDR detour, coverage rows, consensus, conservation bits and the
    DR zone must be those of the reference's Aligner (golden vectors) and of the oracle (fresh groups, many per call)."""
    import hashlib
    import struct
    from test_oracle_golden import consensus_expected
    g = json.load(open(os.path.join(G, "consensus_vectors.json")))
    exp = [consensus_expected(grp) for grp in g["groups"]]
    got, status = ctx.consensus_groups([e[0] for e in exp])
    assert status == 0
    for (case, want, md5, cov), out in zip(exp, got):
        c = out.pop("coverage")
        out.pop("flags")
        assert want.pop("status") == 0 and out == want
        assert hashlib.md5(struct.pack("<%di" % len(c), *c)).hexdigest() == md5
    rng = random.Random(113)
    for batch in range(6):
        cases = [fuzzgen.consensus_case(rng, alphabet=b"ACGTN" if k % 5 == 0 else b"ACGT") for k in range(1 if batch == 0 else 60)]
        got, status = ctx.consensus_groups(cases)
        assert status == 0
        turned = 0
        for case, out in zip(cases, got):
            want = P.consensus_group(case)
            assert want.pop("status") == 0
            flags = out.pop("flags")
            assert out == want
            assert all(((f >> 1) & 1) == (p < 0) for f, p in zip(flags[1:], out["place"][1:]))
            turned += sum(out["reversed"])
        assert batch == 0 or turned > 20


def test_consensus_of_a_bundled_read_set_feeds_partial_repeat_recovery(ctx, P, k6path):
    """The chain behind the scan, on the reference's own reads: K1's hits of Ill100.fx.gz -> low-lexi DR tokens -> the groups
    createNonRedundantSet forms -> K7 (master = longest DR of the group, every other DR aligned against it, consensus and zone
    from the coverage of all their reads) -> the zone's consensus as the DR K6 shifts every read's repeats to.  K7 against the
    oracle's Aligner group by group, K6 against the oracle's updateStartStops read by read."""
    if not os.path.exists(os.path.join(checkers.REF_DATA, "Ill100.fx.gz")):
        pytest.skip("bundled read sets not staged")
    n_groups = n_drs = n_jobs = n_grew = 0
    for name in ("Ill100.fx.gz", "CN_gDC.fa.gz", "front_offset_bug.fa.gz"):
        a, b, c, d = _consensus_chain(ctx, P, os.path.join(checkers.REF_DATA, name))
        n_groups += a; n_drs += b; n_jobs += c; n_grew += d
    assert n_groups >= 3 and n_drs > 50 and n_jobs > 200 and n_grew > 0


def _consensus_chain(ctx, P, path):
    batch = cb.Batch.from_file(path)
    bases, offsets = batch.bases, batch.offsets
    hits, pool, _ = ctx.dr_search(bases, offsets)
    reads_of, order = {}, []                                             # token string -> oriented reads, tokens in first-appearance order
    for h in sorted(hits, key=lambda h: int(h["read_index"])):
        r, o, n = int(h["read_index"]), int(h["ss_offset"]), int(h["n_ss"])
        seq = bytes(bases[int(offsets[r]): int(offsets[r + 1])])
        dr, low, ss2, seq2 = P.dr_lowlexi(seq, [int(x) for x in pool[o: o + n]])
        if dr not in reads_of:
            reads_of[dr] = []
            order.append(dr)
        reads_of[dr].append((seq2, ss2))
    groups = {}
    for line in api.non_redundant_set(order, 6).split("\n"):
        if line.startswith("G\t"):
            _, tok, gid = line.split("\t")
            groups.setdefault(int(gid), []).append(order[int(tok) - 2])
    cases = []
    for gid in sorted(groups):
        drs = groups[gid]
        m = max(range(len(drs)), key=lambda i: (len(drs[i]), -i))         # findMasterDR: the first of the longest
        drs = [drs[m]] + drs[:m] + drs[m + 1:]
        cases.append(dict(drs=drs, reads=[(s_, ss_, d) for d, dr in enumerate(drs) for (s_, ss_) in reads_of[dr]], array_len=4 * batch.max_read_len))
    got, status = ctx.consensus_groups(cases)
    assert status == 0
    jobs, job_drs, all_reads = [], [], []
    for case, out in zip(cases, got):
        want = P.consensus_group(case)
        assert want.pop("status") == 0
        out.pop("flags")
        assert out == want
        true_dr = out["consensus"][out["zone"][0]: out["zone"][1] + 1]
        if not (23 <= len(true_dr) <= 47):
            continue
        job_drs.append(true_dr)
        for (seq, ss, d) in case["reads"]:
            if out["place"][d] < 0 or out["reversed"][d]:
                continue
            front = out["place"][d] - out["zone"][0]                      # where the read's DR starts inside the consensus DR
            jobs.append((len(all_reads), ss, front, len(job_drs) - 1))
            all_reads.append(seq)
    if not jobs:
        return len(cases), sum(len(c["drs"]) for c in cases), 0, 0
    b2, o2 = cb.pack_reads(all_reads)
    res = ctx.update_start_stops(b2, o2, job_drs, jobs)
    grew = 0
    for (st, out), (r, ss, front, d) in zip(res, jobs):
        want = P.update_start_stops(all_reads[r], ss, front, job_drs[d])
        assert (st, out) == ((3, []) if want[0] == -3 else (0, want[1]))
        grew += len(out) > len(ss)
    return len(cases), sum(len(c["drs"]) for c in cases), len(jobs), grew
