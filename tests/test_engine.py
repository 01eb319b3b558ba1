"""The multi-device whole path behind the C-ABI (crass_b200_engine_*).

One caller, one set of containers: whatever the number of devices, the dump (token numbering, DRs, read order,
orientation, start/stops, groups, patterns) must be byte-identical to the reference's / the oracle's / a one-device run.
The emulated tests name the SAME GPU several times -- the shards then share it, but every piece of multi-device logic
(sharding, per-lane threads and streams, token blocks, the gather, the K4c merge with rank-ordered keys, one read-ordered
hit list, replay with duplicate headers) is exercised, and they cannot skip on a one-GPU box.  The last test needs two
GPUs and then also goes through the NCCL all-gather.
"""
import gzip
import os
import random

import numpy as np
import pytest

import checkers
import fuzzgen
import crass_b200 as cb
from crass_b200 import synth

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BUNDLED = ["Ill100.fx.gz", "CN_gDC.fa.gz", "Ill.nr.miss.fa.gz", "front_offset_bug.fa.gz", "poor_dr_ext.fa.gz"]


@pytest.fixture(scope="module")
def P():
    return checkers.port()


def write_fasta(path, reads, names=None):
    with open(path, "wb") as fh:
        for i, s in enumerate(reads):
            fh.write(b">" + (names[i] if names else b"r%07d" % i) + b"\n" + s + b"\n")


@pytest.fixture(params=["peer", "host"])
def exchange(request):
    """how the shards' DR tokens meet: token blocks gathered on the first device (peer copies here; NCCL needs distinct
    devices), or the host containers' token list"""
    old = os.environ.get("CRASS_B200_EXCHANGE")
    os.environ["CRASS_B200_EXCHANGE"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("CRASS_B200_EXCHANGE", None)
    else:
        os.environ["CRASS_B200_EXCHANGE"] = old


@pytest.mark.parametrize("name", BUNDLED)
@pytest.mark.parametrize("devices", [(0,), (0, 0, 0)])
@pytest.mark.parametrize("stream_bytes", [30000, 250000])
def test_bundled_files_streamed(name, devices, stream_bytes):
    """The streamed feed: a file of two ranges or more is parsed range by range while another thread copies and searches the
    range before and a third replays the one before that; the ranges then behave like the files of a multi-file run.  The
    dump must be the reference's whatever the range size (here a few dozen to a few hundred ranges per file), also with
    small parser pieces inside the ranges and on several (emulated) devices."""
    path = os.path.join(checkers.REF_DATA, name)
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged")
    want = gzip.open(os.path.join(G, "bundled", name + ".dump.gz")).read().decode("latin-1")
    old = {k: os.environ.get(k) for k in ("CRASS_B200_STREAM_BYTES", "CRASS_B200_PARSE_CHUNK", "CRASS_B200_PARSE_THREADS", "CRASS_B200_GZ_STREAM_MIN", "CRASS_B200_GZ_STREAM_MARGIN")}
    os.environ.update(CRASS_B200_STREAM_BYTES=str(stream_bytes), CRASS_B200_PARSE_CHUNK=str(stream_bytes // 3), CRASS_B200_PARSE_THREADS="4",
                      CRASS_B200_GZ_STREAM_MIN="1", CRASS_B200_GZ_STREAM_MARGIN="20000")     # the archives are inflated while their first ranges are parsed
    try:
        n_ranges = sum(1 for _ in cb.Batch.stream_file(path, stream_bytes))
        res, max_len = cb.run_files_multi(devices, [path])
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert n_ranges >= 2 or stream_bytes > 30000 or name == "poor_dr_ext.fa.gz"
    assert res.dump(max_len) == want


@pytest.mark.parametrize("name", BUNDLED)
@pytest.mark.parametrize("devices", [(0,), (0, 0), (0, 0, 0, 0, 0)])
def test_bundled_files_on_emulated_devices(name, devices, exchange):
    path = os.path.join(checkers.REF_DATA, name)
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged")
    want = gzip.open(os.path.join(G, "bundled", name + ".dump.gz")).read().decode("latin-1")
    res, max_len = cb.run_files_multi(devices, [path])
    assert res.dump(max_len) == want


def test_synthetic_shards_equal_the_oracle(P, tmp_path, exchange):
    """a config-2 style prefix over 1, 2, 3 and 8 shards: all dumps equal the oracle's whole-path dump"""
    genome, _, _ = synth.make_genome(20242, n_dr_types=12, array_fraction=0.03)
    n = 60_000
    bases, offs = synth.sample_fixed(genome, n, 150, 777)
    reads = [bases[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n)]
    path = str(tmp_path / "reads.fa")
    write_fasta(path, reads)
    want, _ = P.run_files([path])
    for devices in [(0,), (0, 0), (0, 0, 0), (0,) * 8]:
        eng = cb.Engine(devices)
        assert eng.num_devices == len(devices)
        res, max_len = eng.run_files([path])
        assert res.dump(max_len) == want, devices
        h2d, d2h = eng.transfer_bytes()
        assert h2d >= n * 150 and d2h > 0 and eng.launch_count > 0
        eng.close()


def test_duplicate_headers_across_shards_and_several_files(P, tmp_path, exchange):
    """readsFound is keyed by header (libcrispr.cpp:411): a read whose name was found in phase 1 -- in ANOTHER shard or in
    another file -- must not be added by phase 2; several files keep their order; empty and tiny files are fine."""
    rng = random.Random(5)
    pool = [fuzzgen.rand_seq(rng, rng.randint(24, 40)) for _ in range(6)]
    files = []
    for f in range(3):
        reads, names = [], []
        for i in range(4000):
            if rng.random() < 0.5:
                reads.append(fuzzgen.planted_read(rng, rng.choice([100, 150, 200]), dr=rng.choice(pool), sub_rate=rng.choice([0, 0.01])))
            else:
                dr = rng.choice(pool)                                   # a lone repeat: only phase 2 can find it
                reads.append(fuzzgen.rand_seq(rng, 40) + dr + fuzzgen.rand_seq(rng, 60))
            names.append(b"dup%03d" % rng.randint(0, 300) if rng.random() < 0.3 else b"f%d_%05d" % (f, i))
        p = str(tmp_path / ("f%d.fa" % f))
        write_fasta(p, reads, names)
        files.append(p)
    empty = str(tmp_path / "empty.fa")
    open(empty, "wb").close()
    tiny = str(tmp_path / "tiny.fa")
    write_fasta(tiny, [fuzzgen.planted_read(rng, 150, dr=pool[0])])
    paths = [files[0], empty, files[1], tiny, files[2]]
    want, _ = P.run_files(paths)
    for devices in [(0,), (0, 0, 0)]:
        res, max_len = cb.run_files_multi(devices, paths)
        assert res.dump(max_len) == want, devices
    # several files go through the streamed pipeline one after the other: with small ranges every file is several ranges (numbered
    # through, each file a kseq stream of its own), with CRASS_B200_STREAM_MB=0 the files are taken whole, one after the other
    old = {k: os.environ.get(k) for k in ("CRASS_B200_STREAM_BYTES", "CRASS_B200_STREAM_MB")}
    try:
        for env in ({"CRASS_B200_STREAM_BYTES": "150000"}, {"CRASS_B200_STREAM_BYTES": "40000"}, {"CRASS_B200_STREAM_MB": "0"}):
            os.environ.pop("CRASS_B200_STREAM_BYTES", None)
            os.environ.pop("CRASS_B200_STREAM_MB", None)
            os.environ.update(env)
            for devices in [(0,), (0, 0, 0)]:
                res, max_len = cb.run_files_multi(devices, paths)
                assert res.dump(max_len) == want, (env, devices)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_engine_steps_and_resident_reuse(P, tmp_path):
    """the three steps a drop-in caller takes (searchFile / exchange / findSingletons): hit lists in global read order with
    batch-wide indices, the matcher of the exchange equal to the one built from the host token list, and phase 2 on the
    resident shards (no second parse) equal to a one-device scan"""
    genome, _, _ = synth.make_genome(4242, n_dr_types=8, array_fraction=0.05)
    n = 30_000
    bases, offs = synth.sample_fixed(genome, n, 150, 4243)
    path = str(tmp_path / "reads.fa")
    write_fasta(path, [bases[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n)])
    ctx = cb.Context(0)
    hits1, pool1, found1 = ctx.dr_search(bases, offs)
    batch = cb.Batch.from_arrays(bases, offs)
    res = cb.Results()
    res.add_phase1(batch, hits1, pool1)
    pats = res.non_redundant(6)
    ac1 = cb.Automaton(pats)
    hits2, pool2, _ = ctx.ac_scan(ac1, bases, offs, skip=found1)
    eng = cb.Engine((0, 0, 0))
    eh, ep = eng.search_file(path)
    assert list(eh["read_index"]) == list(hits1["read_index"]) and list(eh["read_index"]) == sorted(eh["read_index"])
    assert [list(ep[h["ss_offset"]:h["ss_offset"] + h["n_ss"]]) for h in eh] == [list(pool1[h["ss_offset"]:h["ss_offset"] + h["n_ss"]]) for h in hits1]
    ac, n_variants, n_patterns = eng.exchange(path)
    assert n_variants == res.num_tokens and n_patterns == len(pats)
    os.rename(path, path + ".moved")                                        # phase 2 must not need the file again
    sh, sp = eng.find_singletons(path, ac)
    assert list(sh["read_index"]) == list(hits2["read_index"])
    assert [list(sp[h["ss_offset"]:h["ss_offset"] + 2]) for h in sh] == [list(pool2[h["ss_offset"]:h["ss_offset"] + 2]) for h in hits2]
    eng.release_file(path)
    eng.close()
    ctx.close()


def test_two_real_gpus_through_nccl(P, tmp_path):
    if cb.device_count() < 2:
        pytest.skip("needs two GPUs")
    genome, _, _ = synth.make_genome(20242, n_dr_types=12, array_fraction=0.03)
    n = 200_000
    bases, offs = synth.sample_fixed(genome, n, 150, 778)
    path = str(tmp_path / "reads.fa")
    write_fasta(path, [bases[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n)])
    want, _ = P.run_files([path])
    devices = tuple(range(min(cb.device_count(), 8)))
    eng = cb.Engine(devices)
    res, max_len = eng.run_files([path])
    assert res.dump(max_len) == want
    assert eng.uses_nccl or os.environ.get("CRASS_B200_EXCHANGE") == "peer"
    eng.close()
    os.environ["CRASS_B200_EXCHANGE"] = "peer"                                  # the gather by peer copies gives the same
    try:
        res, max_len = cb.run_files_multi(devices, [path])
        assert res.dump(max_len) == want
    finally:
        os.environ.pop("CRASS_B200_EXCHANGE", None)
