"""Pins the plain-C oracle against the compiled, unmodified reference (oracle/_ref/libcrass_ref.so).

Skipped when the reference library was not built (it is built wherever /root/reference exists and
travels to the GPU box inside oracle/_ref/).
"""
import os
import random

import pytest

import checkers
import fuzzgen

pytestmark = pytest.mark.skipif(not checkers.have_ref(), reason="oracle/_ref/libcrass_ref.so not built")


@pytest.fixture(scope="module")
def R():
    return checkers.ref()


@pytest.fixture(scope="module")
def P():
    return checkers.port()


def test_reference_catch_binary_passes():
    exe = os.path.join(checkers.ORACLE_DIR, "_ref", "crass-test")
    if not os.path.exists(exe):
        pytest.skip("crass-test not built")
    import subprocess
    # The reference's test binary never initialises its global logger (LoggerSimp members are read
    # uninitialised), so it can crash depending on heap garbage.  glibc's
    # MALLOC_PERTURB_=255 makes fresh allocations zero-filled => log level 0 => deterministic.
    out = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, MALLOC_PERTURB_="255"))
    assert out.returncode == 0 and "139 assertions in 7 test cases" in out.stdout


def test_search_core_fuzz(R, P):
    rng = random.Random(11)
    hits = 0
    for _ in range(30000):
        s = fuzzgen.fuzz_read(rng)
        prm = None
        if rng.random() < 0.3:
            prm = dict(window=rng.choice([6, 7, 8, 9]), min_repeats=rng.choice([2, 3, 4]), low_dr=rng.choice([23, 23, 20, 17, 30]),
                       high_dr=rng.choice([47, 40, 60]), low_spacer=rng.choice([26, 20, 30, 10]), high_spacer=rng.choice([50, 60, 40]))
        a = R.search_core(s, prm)
        assert a == P.search_core(s, prm), (s, prm)
        hits += a[0] == 1
    assert hits > 2000


def test_long_read_fuzz(R, P):
    rng = random.Random(12)
    hits = 0
    for _ in range(300):
        L = rng.randint(1000, 6000)
        s = fuzzgen.planted_read(rng, L, sub_rate=rng.choice([0, 0.005, 0.02])) if rng.random() < 0.6 else fuzzgen.rand_seq(rng, L)
        a = R.search_core(s)
        assert a == P.search_core(s)
        hits += a[0] == 1
    assert hits > 50


def test_edit_distance_fuzz(R, P):
    rng = random.Random(13)
    for _ in range(20000):
        a = fuzzgen.rand_seq(rng, rng.randint(0, 55), b"ACGTN")
        b = fuzzgen.mutate(rng, a, 0.2) if rng.random() < 0.6 else fuzzgen.rand_seq(rng, rng.randint(0, 55))
        if rng.random() < 0.3:
            b = b[rng.randint(0, 3):]
        assert R.edit_distance(a, b) == P.edit_distance(a, b)
        assert R.similarity(a, b) == P.similarity(a, b)


def test_qc_and_extend_fuzz(R, P):
    rng = random.Random(14)
    for _ in range(8000):
        L = rng.randint(120, 400)
        s = fuzzgen.planted_read(rng, L, sub_rate=rng.choice([0, 0.02]))
        n = rng.choice([2, 2, 3, 4])
        starts, pos = [], rng.randint(0, 10)
        for _k in range(n):
            starts.append(pos)
            pos += rng.randint(34, 90)
        if starts[-1] + 8 >= L:
            continue
        ss = []
        for st in starts:
            ss += [st, st + 7]
        assert R.extend_pre_repeat(s, ss, 8, 26) == P.extend_pre_repeat(s, ss, 8, 26)
        ln = rng.randint(20, 40)
        ss2 = []
        for st in starts:
            ss2 += [st, min(L - 1, st + ln - 1)]
        assert R.qc_found_repeats(s, ss2) == P.qc_found_repeats(s, ss2)


def test_smith_waterman_and_update_start_stops_fuzz(R, P):
    rng = random.Random(17)
    grew = 0
    for _ in range(3000):
        seq, ss, front, dr = fuzzgen.uss_case(rng)
        start = rng.randint(0, len(seq) - 1)
        length = rng.randint(1, len(seq) - start)
        sim = 0.85 if rng.random() < 0.8 else 0.0
        assert R.smith_waterman(seq, dr, start, length, sim) == P.smith_waterman(seq, dr, start, length, sim)
        low = rng.choice([26, 26, 20, 35])
        got = P.update_start_stops(seq, ss, front, dr, low)
        if got[0] == -3:
            continue
        want = R.update_start_stops(seq, ss, front, dr, low)
        assert got == want
        grew += len(want[1]) > len(ss)
    assert grew > 1000


def test_ksw_align_and_consensus_fuzz(R, P):
    """The consensus step (second half): the lane-by-lane restatement of ksw_i16 / ksw_align and of the Aligner
    against the reference's own ksw.c and Aligner.cpp -- alignment end points, tie breaks of the striped kernel, second-best
    scores, both strands, the extendSlaveDR detour, coverage, consensus, conservation bits and the DR zone."""
    rng = random.Random(18)
    nt = bytes.maketrans(b"ACGTN", bytes([0, 1, 2, 3, 4]))
    for _ in range(2000):
        t = fuzzgen.rand_seq(rng, rng.randint(20, 70))
        if rng.random() < 0.75:
            a, b = sorted(rng.sample(range(len(t)), 2))
            q = fuzzgen.mutate(rng, t[a:b + 1], rng.choice([0, 0.05, 0.2]), b"ACGTN")
            if rng.random() < 0.3 and len(q) > 6:
                k = rng.randint(1, len(q) - 2)
                q = q[:k] + fuzzgen.rand_seq(rng, rng.randint(1, 3)) + q[k:]
            if rng.random() < 0.3 and len(q) > 8:
                k = rng.randint(1, len(q) - 4)
                q = q[:k] + q[k + rng.randint(1, 3):]
        else:
            q = fuzzgen.rand_seq(rng, rng.randint(1, 60))
        if not q:
            continue
        xtra = rng.choice([0x80000 | 0x40000 | 5, 0x80000, 0x40000 | 12, 0])
        assert R.ksw_align(q.translate(nt), t.translate(nt), xtra) == P.ksw_align(q.translate(nt), t.translate(nt), xtra)
    turned = failed = 0
    for it in range(500):
        case = fuzzgen.consensus_case(rng, alphabet=b"ACGTN" if it % 5 == 0 else b"ACGT")
        want, got = R.consensus_group(case), P.consensus_group(case)
        assert want["status"] == 0 and got == want
        turned += sum(want["reversed"])
        failed += sum(1 for p in want["place"] if p < 0)
    assert turned > 200 and failed > 200


def test_lowlexi_fuzz(R, P):
    rng = random.Random(15)
    for _ in range(10000):
        L = rng.randint(60, 300)
        s = fuzzgen.rand_seq(rng, L, b"ACGTNRYacgtU`")
        n = rng.choice([1, 2, 2, 3, 4])
        pts = sorted(rng.sample(range(0, L), 2 * n))
        if rng.random() < 0.3:
            pts[0] = 0
        if rng.random() < 0.3:
            pts[-1] = L - 1
        assert R.dr_lowlexi(s, pts) == P.dr_lowlexi(s, pts)


def test_ac_fuzz(R, P):
    rng = random.Random(16)
    matches = 0
    for _ in range(120):
        pats = fuzzgen.dr_like_patterns(rng, rng.randint(1, 150))
        if rng.random() < 0.3:
            pats += [p[rng.randint(0, 5):len(p) - rng.randint(0, 5)] for p in pats[:10]]
        if rng.random() < 0.2:
            pats = [fuzzgen.mutate(rng, p, 0.05, b"ACGTN") for p in pats]
        hr, hp = R.ac_create(pats), P.ac_create(pats)
        for _k in range(150):
            t = fuzzgen.rand_seq(rng, rng.randint(0, 200), b"ACGTN" if rng.random() < 0.2 else b"ACGT")
            if rng.random() < 0.6 and len(t) > 60:
                p = rng.choice(pats)
                pos = rng.randint(0, len(t) - 1)
                t = (t[:pos] + p + t[pos:])[:len(t)]
            a = R.ac_first_match(hr, t)
            assert a == P.ac_first_match(hp, t)
            matches += a is not None
        R.ac_destroy(hr)
        P.ac_destroy(hp)
    assert matches > 3000


def test_non_redundant_fuzz(R, P):
    rng = random.Random(17)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for _ in range(150):
        base = [fuzzgen.rand_seq(rng, rng.randint(23, 47)) for _k in range(rng.randint(1, 12))]
        drs = []
        for b in base:
            for _k in range(rng.randint(1, 8)):
                v = fuzzgen.mutate(rng, b, rng.choice([0, 0.02, 0.05]), b"ACGT")
                a, e = rng.randint(0, 4), rng.randint(0, 4)
                v = v[a:len(v) - e] if rng.random() < 0.5 else fuzzgen.rand_seq(rng, a) + v + fuzzgen.rand_seq(rng, e)
                rc = v.translate(comp)[::-1]
                drs.append(min(v, rc))
        rng.shuffle(drs)
        seen, uniq = set(), []
        for d in drs:
            if d not in seen:
                seen.add(d)
                uniq.append(d)
        a, b = R.non_redundant(uniq), P.non_redundant(uniq)
        ga = [l for l in a.split("\n") if l.startswith("G")]
        gb = [l for l in b.split("\n") if l.startswith("G")]
        assert ga == gb
        assert sorted(l for l in a.split("\n") if l.startswith("P")) == sorted(l for l in b.split("\n") if l.startswith("P"))


@pytest.mark.parametrize("name", ["Ill100.fx.gz", "CN_gDC.fa.gz", "Ill.nr.miss.fa.gz", "front_offset_bug.fa.gz", "poor_dr_ext.fa.gz"])
def test_bundled_files_ref_vs_port(R, P, name):
    path = os.path.join(checkers.REF_DATA, name)
    assert R.kseq_dump(path) == P.kseq_dump(path)
    for prm in (None, dict(window=6), dict(min_repeats=3), dict(window=9, low_dr=20)):
        a, _ = R.run_files([path], prm)
        b, _ = P.run_files([path], prm)
        assert a == b, (name, prm)


def test_the_step_between_the_phases_is_the_references_own_code(R, P):
    """clusterDRReads / removeRedundantRepeats / createNonRedundantSet in oracle/_ref are cut verbatim out of the
    reference's WorkHorse.cpp at build time (oracle/refshim/gen_workhorse_excerpt.py) -- the checker does not restate them.
    The restatement kept in the harness (CRASS_REF_CLUSTER=restated) and the plain-C oracle must agree with it, also on
    letters outside A/C/G/T and on lists of the size a multi-GPU run merges."""
    import ctypes as C
    import os
    R.lib.ref_cluster_impl.restype = C.c_char_p
    assert b"WorkHorse.cpp" in R.lib.ref_cluster_impl()
    rng = random.Random(23)
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    for trial in range(6):
        base = [fuzzgen.rand_seq(rng, rng.randint(23, 47)) for _k in range(rng.randint(3, 40))]
        drs = set()
        for b in base:
            for _k in range(rng.randint(1, 60 if trial == 5 else 12)):
                v = fuzzgen.mutate(rng, b, rng.choice([0, 0.02, 0.05]), b"ACGTN" if trial % 2 else b"ACGT")
                a, e = rng.randint(0, 4), rng.randint(0, 4)
                v = v[a:len(v) - e] if rng.random() < 0.5 else fuzzgen.rand_seq(rng, a) + v + fuzzgen.rand_seq(rng, e)
                drs.add(min(v, v.translate(comp)[::-1]))
        drs = list(drs)
        rng.shuffle(drs)
        real = R.non_redundant(drs)
        os.environ["CRASS_REF_CLUSTER"] = "restated"
        try:
            assert b"restatement" in R.lib.ref_cluster_impl()
            restated = R.non_redundant(drs)
        finally:
            os.environ.pop("CRASS_REF_CLUSTER", None)
        port = P.non_redundant(drs)
        for other in (restated, port):
            assert [l for l in real.split("\n") if l.startswith("G")] == [l for l in other.split("\n") if l.startswith("G")]
            assert sorted(l for l in real.split("\n") if l.startswith("P")) == sorted(l for l in other.split("\n") if l.startswith("P"))


def test_kseq_on_damaged_archives(R, P, tmp_path):
    """A gzread that FAILS (a damaged archive) is taken for a short read by kstream (kseq.cpp:55-96): the stream ends at the
    4096-byte call that met the damage, and the ks_getc that met it hands out the read buffer's first byte once more.  The port
    and the product's parser restate that; both must give the reference's record stream, byte for byte -- ordinary gzip
    archives and BGZF ones, damage in names, sequences and qualities."""
    import gzip
    import crass_b200 as cb
    from test_host_logic import _write_bgzf
    rng = random.Random(21)
    parts = []
    for k in range(12000):
        seq = fuzzgen.rand_seq(rng, rng.randint(30, 150)).decode()
        if rng.random() < 0.5:
            parts.append("@r%d c%d\n%s\n+\n%s\n" % (k, k, seq, "".join(rng.choice("!5I@>F") for _ in seq)))
        else:
            parts.append(">r%d\n%s\n" % (k, seq))
    text = "".join(parts).encode()
    plain = gzip.compress(text, compresslevel=6)
    cases = 0
    old = {k: os.environ.get(k) for k in ("CRASS_B200_GZ_STREAM_MIN", "CRASS_B200_GZ_SERIAL")}
    os.environ["CRASS_B200_GZ_STREAM_MIN"] = "1"
    try:
        for it in range(10):
            p = str(tmp_path / ("bad%d.fx.gz" % it))
            if it % 2 == 0:
                at = rng.randint(len(plain) // 10, len(plain) - 100)
                with open(p, "wb") as fh:
                    fh.write(plain[:at] + bytes([plain[at] ^ (1 << rng.randint(0, 7))]) + plain[at + 1:])
            else:
                _write_bgzf(p, text, corrupt_block=rng.randint(1, len(text) // 0xff00 - 1))
            want = R.kseq_dump(p)
            assert P.kseq_dump(p) == want, it
            assert cb.Batch.from_file(p).record_stream() == want, it           # (a damaged BGZF archive falls back to zlib's read)
            for serial in (False, True):                                       # streamed: BGZF blocks on several threads / everything through zlib
                if serial:
                    os.environ["CRASS_B200_GZ_SERIAL"] = "1"
                got = [x.record_stream() for x in cb.Batch.stream_file(p, 300000)]
                os.environ.pop("CRASS_B200_GZ_SERIAL", None)
                assert b"".join(x[:x.rindex(b"#ret=")] for x in got[:-1]) + got[-1] == want, (it, serial)
            cases += 1
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert cases == 10
