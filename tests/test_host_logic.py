"""Host-side logic of the product library, no GPU needed: feed path (kseq-compatible parser), replay
into the ReadMap/StringCheck mirror, the step between the phases, automaton construction, and the C-ABI
surface.  Device results are stood in for by the oracle here (test_gpu_parity.py uses the real kernels)."""
import ctypes as C
import gzip
import json
import os
import random
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import checkers
import fuzzgen
import crass_b200 as cb
from crass_b200 import api

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BUNDLED = ["Ill100.fx.gz", "CN_gDC.fa.gz", "Ill.nr.miss.fa.gz", "front_offset_bug.fa.gz", "poor_dr_ext.fa.gz"]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(checkers.ROOT, "include", "crass_b200.h")).read()
    names = sorted(set(re.findall(r"\b(crass_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 40
    L = C.CDLL(api.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert cb.lib().crass_b200_abi_version() == 1


def test_no_device_means_loud_failure():
    if cb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(cb.CrassB200Error) as e:
        cb.Context(0)
    assert e.value.status == api.ENODEVICE


def test_default_params_match_reference_defaults():
    p = cb.Params()
    assert p.as_dict() == dict(low_dr=23, high_dr=47, low_spacer=26, high_spacer=50, window=8, min_repeats=2, kmer_clust=6, scan_range=24)


def test_parser_kseq_vectors():
    with tempfile.TemporaryDirectory() as d:
        for name, v in json.load(open(os.path.join(G, "kseq_vectors.json"))).items():
            for gz in (False, True):
                p = os.path.join(d, name + (".gz" if gz else ".fx"))
                with (gzip.open if gz else open)(p, "wb") as fh:
                    fh.write(v["content"].encode())
                assert cb.Batch.from_file(p).record_stream().decode("latin-1") == v["records"], name


@pytest.mark.parametrize("name", BUNDLED)
def test_parser_bundled_files(name):
    path = os.path.join(checkers.REF_DATA, name)
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged")
    b = cb.Batch.from_file(path)
    assert b.record_stream() == checkers.port().kseq_dump(path)
    lens = np.diff(b.offsets.astype(np.int64))
    assert int(lens.max()) == b.max_read_len


def test_parser_fuzz_against_oracle():
    rng = random.Random(3)
    P = checkers.port()
    with tempfile.TemporaryDirectory() as d:
        for it in range(60):
            parts = []
            for _ in range(rng.randint(0, 12)):
                seq = fuzzgen.rand_seq(rng, rng.randint(1, 120), b"ACGTNacgt")
                name = "r%d" % rng.randint(0, 999)
                cmt = rng.choice(["", " c", "\tc d", " ", "  x"])
                wrap = rng.choice([0, 0, 30, 60])
                body = seq.decode() if not wrap else "\n".join(seq.decode()[i:i + wrap] for i in range(0, len(seq), wrap))
                if rng.random() < 0.5:
                    parts.append(">%s%s\n%s\n" % (name, cmt, body))
                else:
                    q = "".join(rng.choice("!#5I@>+") for _ in seq)
                    if rng.random() < 0.05:
                        q = q[:-1]
                    parts.append("@%s%s\n%s\n+%s\n%s\n" % (name, cmt, body, rng.choice(["", name]), q))
            text = "".join(parts)
            if rng.random() < 0.2:
                text = text.replace("\n", "\r\n")
            if rng.random() < 0.2:
                text = text.rstrip("\n")
            p = os.path.join(d, "f%d.fx" % it)
            with open(p, "wb") as fh:
                fh.write(text.encode())
            assert cb.Batch.from_file(p).record_stream() == P.kseq_dump(p), text


_NAIVE = """
import os, sys
sys.path.insert(0, %r)
import crass_b200 as cb
n = 0
for it in range(40):
    p = os.path.join(sys.argv[1], "f%%d.fx" %% it)
    want = open(p + ".want", "rb").read()
    for chunk in (64, 300, 1500):
        os.environ["CRASS_B200_PARSE_CHUNK"] = str(chunk)
        assert cb.Batch.from_file(p).record_stream() == want, (it, chunk)
        # the same through the streamed feed, whose pieces stay where they were parsed (segments): dropped pieces, gaps
        # parsed again in place, pieces whose last record overran their slice
        got = [x.record_stream() for x in cb.Batch.stream_file(p, 5 * chunk + 3)]
        assert b"".join(x[:x.rindex(b"#ret=")] for x in got[:-1]) + got[-1] == want, (it, chunk, "streamed")
    for b in cb.Batch.stream_file(p, 700):                   # record by record == the back-to-back copy
        off, bases = b.offsets, b.bases
        for i in range(len(b)):
            assert b.read(i) == bytes(bases[int(off[i]):int(off[i + 1])]), (it, i)
    n += 1
print("ok", n)
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pieces_env(chunk, threads=4):
    old = {k: os.environ.get(k) for k in ("CRASS_B200_PARSE_CHUNK", "CRASS_B200_PARSE_THREADS")}
    os.environ["CRASS_B200_PARSE_CHUNK"] = str(chunk)
    os.environ["CRASS_B200_PARSE_THREADS"] = str(threads)
    return old


def _restore_env(old):
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def test_parser_pieces_on_worker_threads_match_the_sequential_stream():
    """parse_file cuts large inputs at guessed record starts; whatever the guesses are, the record stream must be the
    one kseq_read (kseq.cpp:171-225) produces sequentially -- stale comments/qualities across the cuts included."""
    rng = random.Random(11)
    P = checkers.port()
    with tempfile.TemporaryDirectory() as d:
        for it in range(40):
            parts = []
            style = rng.choice(["fa", "fq", "mixed", "hostile"])
            for _ in range(rng.randint(20, 200)):
                seq = fuzzgen.rand_seq(rng, rng.randint(1, 150), b"ACGTNacgt").decode()
                name = "r%d" % rng.randint(0, 99999)
                cmt = rng.choice(["", "", " c%d" % rng.randint(0, 99), "\tc d"])
                wrap = rng.choice([0, 0, 0, 40])
                body = seq if not wrap else "\n".join(seq[i:i + wrap] for i in range(0, len(seq), wrap))
                fq = style == "fq" or (style in ("mixed", "hostile") and rng.random() < 0.6)
                if not fq:
                    parts.append(">%s%s\n%s\n" % (name, cmt, body))
                    continue
                alphabet = "@>+I5" if style == "hostile" else "!#5I@>+FFFFFF"
                q = "".join(rng.choice(alphabet) for _ in seq)
                if style == "hostile" and rng.random() < 0.5:
                    q = rng.choice("@>") + q[1:]
                if rng.random() < 0.01:
                    q = q[:-1]
                parts.append("%s%s%s\n%s\n+%s\n%s\n" % (rng.choice("@@@>"), name, cmt, body, rng.choice(["", name]), q))
            text = "".join(parts)
            if rng.random() < 0.15:
                text = text.replace("\n", "\r\n")
            if rng.random() < 0.1:
                cut = rng.randint(0, len(text))
                text = text[:cut] + "\xff" + text[cut:]
            p = os.path.join(d, "f%d.fx" % it)
            with open(p, "wb") as fh:
                fh.write(text.encode("latin-1"))
            want = P.kseq_dump(p)
            with open(p + ".want", "wb") as fh:
                fh.write(want)
            for chunk in (64, 257, 1500, 6000):
                old = _pieces_env(chunk)
                try:
                    b = cb.Batch.from_file(p)
                    got = b.record_stream()
                finally:
                    _restore_env(old)
                assert got == want, (style, chunk, it)
                lens = np.diff(b.offsets.astype(np.int64))
                assert (int(lens.max()) if len(lens) else 0) == b.max_read_len
                # the streamed feed: the same records range by range (a range = a few pieces), each range inheriting kseq's
                # stale strings; only the last batch carries the stream's status
                old = _pieces_env(chunk)
                try:
                    parts_got = [x.record_stream() for x in cb.Batch.stream_file(p, 3 * chunk + 17)]
                finally:
                    _restore_env(old)
                assert len(parts_got) >= 1 and all(x.endswith(b"#ret=0\n") for x in parts_got[:-1])
                joined = b"".join(x[:x.rindex(b"#ret=")] for x in parts_got[:-1]) + parts_got[-1]
                assert joined == want, (style, chunk, it, "streamed")
        # wrong guesses on purpose (any '>'/'@' byte is taken for a record start): pieces are dropped and the gaps parsed
        # again from the true position.  The switch is read once per process, hence the child.
        r = subprocess.run([sys.executable, "-c", _NAIVE, d], env=dict(os.environ, CRASS_B200_PARSE_GUESS="naive", CRASS_B200_PARSE_THREADS="4"),
                           capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.strip() == "ok 40", r.stdout + r.stderr


def _write_bgzf(path, data, block=0xff00, corrupt_block=None):
    """what bgzip writes: gzip members of at most 64 KB with the BC extra field (total member size - 1) and an empty last one"""
    import struct
    import zlib

    def member(chunk):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp +
                struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    with open(path, "wb") as fh:
        for k, i in enumerate(range(0, len(data), block)):
            m = member(data[i:i + block])
            if corrupt_block == k:
                m = m[:40] + bytes([m[40] ^ 0x55]) + m[41:]
            fh.write(m)
        fh.write(member(b""))


def test_parser_inflates_a_bgzf_archive_on_several_threads():
    """A BGZF archive (bgzip / htslib) says in every member's header how long the member is, so its blocks are inflated by
    several threads, each block straight to its place in the output, while the parser takes the ranges that are complete
    (the reference inflates inside its read loop, SeqUtils.cpp:100-125).  Same record stream as zlib's sequential read;
    an archive that stops being BGZF half way is read by the one-thread path; a damaged block ends the stream where it
    ends for the reference (that stretch is read again through zlib, 4096 bytes per call)."""
    rng = random.Random(14)
    P = checkers.port()
    with tempfile.TemporaryDirectory() as d:
        parts = []
        for k in range(50000):
            seq = fuzzgen.rand_seq(rng, rng.randint(30, 150)).decode()
            if rng.random() < 0.5:
                parts.append("@r%d c%d\n%s\n+\n%s\n" % (k, k, seq, "".join(rng.choice("!5I@>F") for _ in seq)))
            else:
                parts.append(">r%d\n%s\n" % (k, seq))
        text = "".join(parts).encode()
        p = os.path.join(d, "b.fx.gz")
        _write_bgzf(p, text)
        assert gzip.open(p).read() == text                               # (a valid multi-member gzip file)
        want = P.kseq_dump(p)
        keys = ("CRASS_B200_GZ_STREAM_MIN", "CRASS_B200_PARSE_CHUNK", "CRASS_B200_PARSE_THREADS", "CRASS_B200_GZ_THREADS", "CRASS_B200_GZ_SERIAL")
        old = {k: os.environ.get(k) for k in keys}
        os.environ.update(CRASS_B200_GZ_STREAM_MIN="1", CRASS_B200_PARSE_CHUNK="200000", CRASS_B200_PARSE_THREADS="4")
        try:
            for threads, range_bytes in (("1", 900000), ("3", 300000), ("8", 2500000), ("5", 0)):
                os.environ["CRASS_B200_GZ_THREADS"] = threads
                got = [x.record_stream() for x in cb.Batch.stream_file(p, range_bytes)]
                assert b"".join(x[:x.rindex(b"#ret=")] for x in got[:-1]) + got[-1] == want, (threads, range_bytes)
            assert cb.Batch.from_file(p).record_stream() == want         # the whole-file form inflates BGZF in parallel too
            os.environ["CRASS_B200_GZ_SERIAL"] = "1"                     # the one-thread path on the same archive
            got = [x.record_stream() for x in cb.Batch.stream_file(p, 900000)]
            assert b"".join(x[:x.rindex(b"#ret=")] for x in got[:-1]) + got[-1] == want
            os.environ.pop("CRASS_B200_GZ_SERIAL")
            # BGZF blocks followed by an ordinary gzip member: not a BGZF archive, zlib reads it all
            p2 = os.path.join(d, "mixed.fx.gz")
            with open(p2, "wb") as fh:
                fh.write(open(p, "rb").read())
                fh.write(gzip.compress(b">tail\nACGTACGTACGTACGT\n"))
            got = [x.record_stream() for x in cb.Batch.stream_file(p2, 900000)]
            assert b"".join(x[:x.rindex(b"#ret=")] for x in got[:-1]) + got[-1] == P.kseq_dump(p2)
            # a damaged block
            p3 = os.path.join(d, "bad.fx.gz")
            _write_bgzf(p3, text, corrupt_block=20)
            want3 = P.kseq_dump(p3)                                       # (pinned against the reference in test_oracle_vs_ref.py)
            for threads in ("1", "4"):
                os.environ["CRASS_B200_GZ_THREADS"] = threads
                got = [x.record_stream() for x in cb.Batch.stream_file(p3, 900000)]
                assert b"".join(x[:x.rindex(b"#ret=")] for x in got[:-1]) + got[-1] == want3, threads
            assert cb.Batch.from_file(p3).record_stream() == want3
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v


def test_parser_streams_a_gz_archive_while_it_inflates():
    """A gz input of a parse stream is inflated by a thread of its own while the ranges that are already there are parsed
    (the reference inflates inside its read loop); the record stream must not depend on how far the inflater is when a range
    is cut -- ranges much smaller than the inflater's steps, records that straddle what has arrived, a truncated last record."""
    import gzip
    rng = random.Random(12)
    P = checkers.port()
    with tempfile.TemporaryDirectory() as d:
        for it, style in enumerate(["fa", "fq", "mixed"]):
            parts = []
            for k in range(60000):
                seq = fuzzgen.rand_seq(rng, rng.randint(30, 150)).decode()
                fq = style == "fq" or (style == "mixed" and rng.random() < 0.5)
                cmt = "" if rng.random() < 0.7 else " c%d" % k
                if fq:
                    parts.append("@r%d%s\n%s\n+\n%s\n" % (k, cmt, seq, "".join(rng.choice("!5I@>F") for _ in seq)))
                else:
                    parts.append(">r%d%s\n%s\n" % (k, cmt, seq))
            text = "".join(parts)
            if it == 1:
                text = text[:-40]                                        # the last quality string is cut short: kseq returns -2
            p = os.path.join(d, "s%d.fx.gz" % it)
            with gzip.open(p, "wb", compresslevel=1) as fh:
                fh.write(text.encode())
            want = P.kseq_dump(p)
            old = {k: os.environ.get(k) for k in ("CRASS_B200_GZ_STREAM_MIN", "CRASS_B200_PARSE_CHUNK", "CRASS_B200_PARSE_THREADS", "CRASS_B200_GZ_STREAM_MARGIN")}
            os.environ.update(CRASS_B200_GZ_STREAM_MIN="1", CRASS_B200_PARSE_CHUNK="200000", CRASS_B200_PARSE_THREADS="4")
            try:
                for range_bytes, margin in ((700000, 100), (700000, 50000), (3000000, 1 << 25)):
                    os.environ["CRASS_B200_GZ_STREAM_MARGIN"] = str(margin)   # tiny margins: views end inside records, ranges are parsed again
                    got = [x.record_stream() for x in cb.Batch.stream_file(p, range_bytes)]
                    assert len(got) >= 2 and all(x.endswith(b"#ret=0\n") for x in got[:-1])
                    assert b"".join(x[:x.rindex(b"#ret=")] for x in got[:-1]) + got[-1] == want, (style, range_bytes)
            finally:
                for k, v in old.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v


@pytest.mark.parametrize("name", BUNDLED)
def test_parser_pieces_bundled_files(name):
    path = os.path.join(checkers.REF_DATA, name)
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged")
    want = cb.Batch.from_file(path)
    old = _pieces_env(20000, threads=3)
    try:
        got = cb.Batch.from_file(path)
    finally:
        _restore_env(old)
    assert got.record_stream() == want.record_stream() == checkers.port().kseq_dump(path)
    assert np.array_equal(got.offsets, want.offsets) and got.max_read_len == want.max_read_len
    assert np.array_equal(got.bases, want.bases)


def test_parse_missing_file_is_an_error():
    with pytest.raises(cb.CrassB200Error):
        cb.Batch.from_file("/nonexistent/file.fa")


def test_non_redundant_set_fuzz_against_oracle():
    rng = random.Random(17)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    P = checkers.port()
    for it in range(240):
        alphabet = (b"ACGT", b"ACGT", b"ACGTN", b"ACGTUR")[it % 4]          # 'U' complements to 'A': not an involution
        base = [fuzzgen.rand_seq(rng, rng.randint(23, 47)) for _k in range(rng.randint(1, 12))]
        drs = []
        for b in base:
            for _k in range(rng.randint(1, 8)):
                v = fuzzgen.mutate(rng, b, rng.choice([0, 0.02, 0.05]), alphabet)
                a, e = rng.randint(0, 4), rng.randint(0, 4)
                v = v[a:len(v) - e] if rng.random() < 0.5 else fuzzgen.rand_seq(rng, a) + v + fuzzgen.rand_seq(rng, e)
                drs.append(min(v, v.translate(comp)[::-1]))
        uniq = list(dict.fromkeys(drs))
        rng.shuffle(uniq)
        a, b = P.non_redundant(uniq), cb.non_redundant_set(uniq)
        assert [l for l in a.split("\n") if l.startswith("G")] == [l for l in b.split("\n") if l.startswith("G")]
        assert sorted(l for l in a.split("\n") if l.startswith("P")) == sorted(l for l in b.split("\n") if l.startswith("P"))


def test_non_redundant_set_long_list_worker_threads():
    """A list long enough (> 2^17 k-mers) to take the multi-threaded clustering passes; a few DRs carry 'N'."""
    rng = random.Random(23)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    P = checkers.port()
    base = [fuzzgen.rand_seq(rng, rng.randint(23, 47)) for _k in range(300)]
    drs = []
    for b in base:
        for _k in range(20):
            v = fuzzgen.mutate(rng, b, rng.choice([0, 0.02, 0.05]), b"ACGT" if rng.random() < 0.97 else b"ACGTN")
            a, e = rng.randint(0, 4), rng.randint(0, 4)
            v = v[a:len(v) - e] if rng.random() < 0.5 else fuzzgen.rand_seq(rng, a) + v + fuzzgen.rand_seq(rng, e)
            drs.append(min(v, v.translate(comp)[::-1]))
    uniq = list(dict.fromkeys(drs))
    rng.shuffle(uniq)
    assert sum(max(0, len(d) - 10) for d in uniq) > (1 << 17)
    a = P.non_redundant(uniq)
    for _rep in range(3):                                                 # the table is reused across calls
        b = cb.non_redundant_set(uniq)
        assert [l for l in a.split("\n") if l.startswith("G")] == [l for l in b.split("\n") if l.startswith("G")]
        assert sorted(l for l in a.split("\n") if l.startswith("P")) == sorted(l for l in b.split("\n") if l.startswith("P"))


def test_pattern_text_round_trip():
    """The two halves of ac_build_from_dr_list used by the root-clusters-for-all exchange."""
    rng = random.Random(37)
    base = [fuzzgen.rand_seq(rng, rng.randint(23, 47)) for _k in range(30)]
    drs = list(dict.fromkeys(fuzzgen.mutate(rng, b, rng.choice([0, 0.03]), b"ACGTN") for b in base for _k in range(6)))
    text = b"".join(d + b"\n" for d in drs)
    pats = api.non_redundant_patterns(text, 6)
    assert pats.split(b"\n")[:-1] == api.non_redundant_list(drs, 6)             # same set, same order
    a, b = cb.Automaton.from_pattern_text(pats), cb.Automaton.from_dr_list(text, 6)
    assert a.num_patterns == b.num_patterns == pats.count(b"\n")
    assert api.non_redundant_patterns(b"", 6) == b""


def test_sort_hits_puts_device_order_into_read_order():
    rng = np.random.default_rng(31)
    for n, top in ((0, 10), (1, 10), (300, 5000), (2000, 2**11), (5000, 2**22 + 5), (70000, 10_000_000), (40000, 2**32 - 1)):
        hits = np.zeros(n, dtype=api.HIT_DTYPE)
        hits["read_index"] = rng.choice(top, size=n, replace=False) if top < 2**31 else rng.integers(0, top, size=n, dtype=np.uint64)
        hits["ss_offset"] = np.arange(n)
        hits["n_ss"] = 2 * (1 + np.arange(n) % 5)
        want = hits[np.argsort(hits["read_index"], kind="stable")]
        api.sort_hits(hits)
        assert np.array_equal(hits, want)


def test_dr_list_from_token_block():
    """Host side of the exchange: a token block (header, records with the order key in their last four bytes)."""
    rng = random.Random(29)
    cap, stride = 50, 64
    drs = [fuzzgen.rand_seq(rng, rng.randint(23, 47)) for _ in range(40)]
    keys = rng.sample(range(10_000_000), len(drs))
    blk = np.zeros(api.token_block_bytes(cap, stride), dtype=np.uint8)
    assert blk.size == 16 + cap * stride
    blk[:4] = np.frombuffer(np.uint32(len(drs)).tobytes(), dtype=np.uint8)
    for slot, (d, k) in enumerate(zip(drs, keys)):
        rec = blk[16 + slot * stride: 16 + (slot + 1) * stride]
        rec[0] = len(d)
        rec[2:2 + len(d)] = np.frombuffer(d, dtype=np.uint8)
        rec[stride - 4:] = np.frombuffer(np.uint32(k).tobytes(), dtype=np.uint8)
    text, count, flags = api.dr_list_from_block(blk, cap, stride)
    assert count == len(drs) and flags == 0
    assert text == b"".join(d + b"\n" for _k, d in sorted(zip(keys, drs)))
    blk[:4] = np.frombuffer(np.uint32(1000).tobytes(), dtype=np.uint8)      # an overflowed block reports its true count
    blk[4] = 1
    text, count, flags = api.dr_list_from_block(blk, cap, stride)
    assert count == 1000 and flags == 1 and text.count(b"\n") == len(drs)   # slots beyond the 40 written ones are empty


def oracle_hits_phase1(batch, params=None):
    """Stand-in for kernel K1 on a box without a GPU: the oracle decides, the product replays."""
    P = checkers.port()
    offs, bases = batch.offsets, batch.bases
    hits, pool = [], []
    for i in range(len(batch)):
        s = bases[int(offs[i]):int(offs[i + 1])].tobytes()
        f, ss, rl = P.search_core(s, params)
        if f == 1:
            hits.append((i, len(ss), len(pool), rl))
            pool += ss
    return np.array(hits, dtype=api.HIT_DTYPE), np.array(pool + [0], dtype=np.uint32)


def oracle_hits_phase2(batch, patterns, skip):
    P = checkers.port()
    h = P.ac_create(patterns)
    offs, bases = batch.offsets, batch.bases
    hits, pool = [], []
    for i in range(len(batch)):
        if skip[i]:
            continue
        s = bases[int(offs[i]):int(offs[i + 1])].tobytes()
        m = P.ac_first_match(h, s)
        if m:
            end, plen = m
            dr_end = min(end - 1, len(s) - 1)
            hits.append((i, 2, len(pool), 0))
            pool += [dr_end - (plen - 1), dr_end]
    P.ac_destroy(h)
    return np.array(hits, dtype=api.HIT_DTYPE), np.array(pool + [0], dtype=np.uint32)


@pytest.mark.parametrize("name", BUNDLED)
def test_replay_reproduces_reference_dump(name):
    """Parser + replay (addReadHolder/DRLowLexi/tokens) + createNonRedundantSet + on_match bookkeeping of the
    PRODUCT, fed with oracle-decided hits, must reproduce the reference's dump byte for byte."""
    path = os.path.join(checkers.REF_DATA, name)
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged")
    want = gzip.open(os.path.join(G, "bundled", name + ".dump.gz")).read().decode("latin-1")
    b = cb.Batch.from_file(path)
    res = cb.Results()
    hits, pool = oracle_hits_phase1(b)
    res.add_phase1(b, hits, pool)
    pats = res.non_redundant(6)
    skip = np.zeros(len(b), dtype=np.uint8)
    skip[hits["read_index"]] = 1
    if pats:
        h2, p2 = oracle_hits_phase2(b, pats, skip)
        res.add_phase2(b, h2, p2)
    assert res.dump(b.max_read_len) == want


def test_adopt_tokens_renumbers_like_a_sequential_run():
    """Multi-GPU merge: shard-local token numbers -> global first-appearance order."""
    path = os.path.join(checkers.REF_DATA, "Ill100.fx.gz")
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged")
    b = cb.Batch.from_file(path)
    hits, pool = oracle_hits_phase1(b)
    whole = cb.Results()
    whole.add_phase1(b, hits, pool)
    n = len(b)
    cut = n // 2
    offs, bases = b.offsets, b.bases
    names = [b.name(i) for i in range(n)]
    shards = []
    for lo, hi in ((0, cut), (cut, n)):
        sb = cb.Batch.from_arrays(bases[int(offs[lo]):int(offs[hi])], offs[lo:hi + 1] - offs[lo], names[lo:hi])
        sh = hits[(hits["read_index"] >= lo) & (hits["read_index"] < hi)].copy()
        sh["read_index"] -= lo
        r = cb.Results()
        r.add_phase1(sb, sh, pool)
        shards.append(r)
    gathered = shards[0].dr_list() + shards[1].dr_list()
    for r in shards:
        r.adopt_tokens(gathered)
    assert shards[0].dr_list() == whole.dr_list() == shards[1].dr_list()
    assert shards[0].non_redundant() == whole.non_redundant()


def test_automaton_shape():
    ac = cb.Automaton([b"ACGTACGTACGTACGTACGTACG", b"ACGTACGTACGTACGTACGTACGTT", b"TTTTTTTTTTTTTTTTTTTTTTTTT"])
    assert ac.num_states == 1 + 25 + 25
    assert ac.table_bytes == ac.num_states * 4 * 4
    with pytest.raises(cb.CrassB200Error):
        cb.Automaton([])


def test_parser_records_longer_than_pieces_and_ranges():
    """Contig-sized records: a record may span many pieces and several ranges' worth of bytes.  Pieces that begin inside it are
    dropped, the piece that holds its header overruns its slice of the shared buffer and goes on in a buffer of its own, a
    range ends on the next true record start however far that is."""
    rng = random.Random(3)
    P = checkers.port()
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "contigs.fa")
        with open(p, "w") as fh:
            for k in range(10):
                L = rng.choice([50, 2000, 150000, 400000, 700000])
                seq = "".join(rng.choice("ACGT") for _ in range(L))
                if k % 3 == 0:
                    fh.write(">c%d single line\n%s\n" % (k, seq))
                else:
                    fh.write(">c%d\n%s\n" % (k, "\n".join(seq[i:i + 60] for i in range(0, L, 60))))
        want = P.kseq_dump(p)
        for chunk, range_bytes in ((100000, 350000), (5000, 20000), (300000, 0)):
            old = _pieces_env(chunk)
            try:
                assert cb.Batch.from_file(p).record_stream() == want, chunk
                got = [x.record_stream() for x in cb.Batch.stream_file(p, range_bytes)]
                assert b"".join(x[:x.rindex(b"#ret=")] for x in got[:-1]) + got[-1] == want, (chunk, range_bytes)
            finally:
                _restore_env(old)


def test_parser_own_inflate_against_zlib():
    """An ordinary .gz of a streamed input is inflated by the library's own DEFLATE decoder (fast_inflate.h; zlib when it gives
    up).  Archives that exercise every block type and header field -- stored, fixed and dynamic Huffman blocks, Huffman-only /
    RLE / filtered strategies, a 512-byte window, sync and full flushes, several members (an empty one among them), a member
    with a file name, trailing garbage, a truncated archive -- must give the record stream zlib's gzread gives."""
    import io
    import zlib
    rng = random.Random(31)
    P = checkers.port()
    parts = []
    for k in range(9000):
        seq = fuzzgen.rand_seq(rng, rng.randint(30, 150)).decode()
        if rng.random() < 0.6:
            parts.append("@read%d/1 x\n%s\n+\n%s\n" % (k, seq, "".join(rng.choice("FFFFFFF:,#") for _ in seq)))
        else:
            parts.append(">r%d\n%s\n" % (k, seq * rng.choice([1, 1, 3])))
    text = "".join(parts).encode()
    archives = {}
    for lvl in (0, 1, 6, 9):
        archives["level%d" % lvl] = gzip.compress(text, compresslevel=lvl)
    for strat, nm in ((zlib.Z_FIXED, "fixed"), (zlib.Z_HUFFMAN_ONLY, "huffman"), (zlib.Z_RLE, "rle"), (zlib.Z_FILTERED, "filtered")):
        c = zlib.compressobj(6, zlib.DEFLATED, 31, 8, strat)
        archives[nm] = c.compress(text) + c.flush()
    c = zlib.compressobj(6, zlib.DEFLATED, 25)
    archives["window512"] = c.compress(text) + c.flush()
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    buf = b""
    for i in range(0, len(text), 7001):
        buf += c.compress(text[i:i + 7001]) + c.flush(zlib.Z_SYNC_FLUSH if (i // 7001) % 2 else zlib.Z_FULL_FLUSH)
    archives["flushes"] = buf + c.flush()
    third = len(text) // 3
    archives["members"] = gzip.compress(text[:third]) + gzip.compress(b"") + gzip.compress(text[third:2 * third], 1) + gzip.compress(text[2 * third:], 0)
    b = io.BytesIO()
    with gzip.GzipFile(filename="reads.fq", mode="wb", fileobj=b, mtime=12345) as g:
        g.write(text)
    archives["named"] = b.getvalue()
    archives["garbage"] = gzip.compress(text) + b"\0\0\0\0 not gzip"
    archives["truncated"] = gzip.compress(text)[:len(gzip.compress(text)) // 2]
    keys = ("CRASS_B200_GZ_STREAM_MIN", "CRASS_B200_GZ_SERIAL", "CRASS_B200_PARSE_CHUNK", "CRASS_B200_PARSE_THREADS")
    old = {k: os.environ.get(k) for k in keys}
    os.environ.update(CRASS_B200_GZ_STREAM_MIN="1", CRASS_B200_PARSE_CHUNK="100000", CRASS_B200_PARSE_THREADS="4")
    try:
        with tempfile.TemporaryDirectory() as d:
            for name, data in archives.items():
                p = os.path.join(d, name + ".fx.gz")
                with open(p, "wb") as fh:
                    fh.write(data)
                want = P.kseq_dump(p)
                assert len(want) > 100000, name
                for range_bytes in (250000, 0):
                    got = [x.record_stream() for x in cb.Batch.stream_file(p, range_bytes)]
                    assert b"".join(x[:x.rindex(b"#ret=")] for x in got[:-1]) + got[-1] == want, (name, range_bytes)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_parser_damaged_archive_fixtures():
    """the product's parser on tests/golden/damaged_gz (record streams of the reference, see test_oracle_golden.py): whole-file
    and streamed, own decoder / parallel BGZF blocks and zlib only"""
    import base64
    import hashlib
    d = os.path.join(G, "damaged_gz")
    want = json.load(open(os.path.join(d, "expected.json")))
    keys = ("CRASS_B200_GZ_STREAM_MIN", "CRASS_B200_GZ_SERIAL")
    old = {k: os.environ.get(k) for k in keys}
    os.environ["CRASS_B200_GZ_STREAM_MIN"] = "1"
    try:
        for name, w in want.items():
            p = os.path.join(d, name)
            for serial in (False, True):
                os.environ.pop("CRASS_B200_GZ_SERIAL", None)
                if serial:
                    os.environ["CRASS_B200_GZ_SERIAL"] = "1"
                whole = cb.Batch.from_file(p).record_stream()
                parts = [x.record_stream() for x in cb.Batch.stream_file(p, 40000)]
                streamed = b"".join(x[:x.rindex(b"#ret=")] for x in parts[:-1]) + parts[-1]
                for got in (whole, streamed):
                    assert len(got) == w["records_len"] and got[-160:] == base64.b64decode(w["tail"]), (name, serial)
                    assert hashlib.md5(got).hexdigest() == w["records_md5"], (name, serial)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
