"""The plain-C oracle (oracle/crass_oracle.c) against the committed golden fixtures.

The fixtures were produced by RUNNING THE REFERENCE (tests/golden/make_golden.py) or restate the
reference's own Catch known-answer tests (catch_vectors.json).  No GPU, no /root/reference needed.
"""
import gzip
import hashlib
import json
import os
import struct
import tempfile

import pytest

import checkers

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return json.load(open(os.path.join(G, name)))


@pytest.fixture(scope="module")
def P():
    return checkers.port()


def test_catch_scan_right(P):
    for v in load("catch_vectors.json")["scan_right"]:
        got = P.scan_right(v["seq"].encode(), v["ss"], v["pattern"].encode(), v["min_spacer"], v["scan_range"])
        assert got == v["expect"], v["cite"]


def test_catch_extend_pre_repeat(P):
    for v in load("catch_vectors.json")["extend_pre_repeat"]:
        got = P.extend_pre_repeat(v["seq"].encode(), v["ss"], v["window"], v["min_spacer"])
        assert got == (v["expect_len"], v["expect"]), v["cite"]


def test_search_core_vectors(P):
    vec = load("search_core_vectors.json")
    assert sum(v["found"] == 1 for v in vec) > 300
    for v in vec:
        assert P.search_core(v["seq"].encode("latin-1"), v["params"]) == (v["found"], v["ss"], v["replen"])


def test_edit_distance_vectors(P):
    for a, b, d, simhex in load("edit_distance_vectors.json"):
        assert P.edit_distance(a.encode(), b.encode()) == d
        assert struct.pack(">f", P.similarity(a.encode(), b.encode())).hex() == simhex


def test_lowlexi_vectors(P):
    for v in load("lowlexi_vectors.json"):
        dr, low, ss, seq = P.dr_lowlexi(v["seq"].encode(), v["ss"])
        assert (dr.decode(), low, ss, seq.decode()) == (v["dr"], v["lowlexi"], v["ss_out"], v["seq_out"])


def test_update_start_stops_vectors(P):
    g = load("update_start_stops_vectors.json")
    grew = 0
    for v in g["update_start_stops"]:
        st, out = P.update_start_stops(v["seq"].encode(), v["ss"], v["front"], v["dr"].encode(), v["low_spacer"])
        assert (st, out) == (0, v["ss_out"])
        grew += len(out) > len(v["ss"])
    assert grew > 100                                    # the fixture does exercise the partial-repeat branches
    for v in g["smith_waterman"]:
        assert list(P.smith_waterman(v["a"].encode(), v["b"].encode(), v["start"], v["len"], v["similarity"])) == v["out"]


def consensus_expected(g):
    """golden group entry -> (case, expected output without the coverage array, md5 of the coverage array)"""
    case = dict(reads=[(r[0].encode(), r[1], r[2]) for r in g["reads"]], drs=[d.encode() for d in g["drs"]], array_len=g["array_len"])
    want = {k: v for k, v in g["out"].items() if k not in ("coverage", "coverage_md5")}
    want["consensus"] = want["consensus"].encode()
    return case, want, g["out"]["coverage_md5"], g["out"].get("coverage")


def test_consensus_vectors(P):
    """ksw_align and the Aligner (consensus DR of a group) against vectors made by the compiled reference."""
    import hashlib
    import struct
    g = load("consensus_vectors.json")
    nt = bytes.maketrans(b"ACGTN", bytes([0, 1, 2, 3, 4]))
    for v in g["ksw_align"]:
        assert list(P.ksw_align(v["q"].encode().translate(nt), v["t"].encode().translate(nt))) == v["out"]
    placed = turned = 0
    for grp in g["groups"]:
        case, want, md5, cov = consensus_expected(grp)
        got = P.consensus_group(case)
        c = got.pop("coverage")
        assert got == want
        assert hashlib.md5(struct.pack("<%di" % len(c), *c)).hexdigest() == md5
        if cov is not None:
            assert c == cov
        placed += sum(1 for p in want["place"][1:] if p >= 0)
        turned += sum(want["reversed"])
    assert placed > 100 and turned > 20                  # the fixture does exercise both strands


def test_ac_vectors(P):
    for case in load("ac_vectors.json"):
        h = P.ac_create([p.encode() for p in case["patterns"]])
        for text, expect in case["texts"]:
            got = P.ac_first_match(h, text.encode())
            assert (list(got) if got else None) == expect
        P.ac_destroy(h)


def test_kseq_vectors(P):
    with tempfile.TemporaryDirectory() as d:
        for name, v in load("kseq_vectors.json").items():
            p = os.path.join(d, name + ".fx")
            with open(p, "wb") as fh:
                fh.write(v["content"].encode())
            assert P.kseq_dump(p).decode("latin-1") == v["records"], name
            with gzip.open(p + ".gz", "wb") as g:
                g.write(v["content"].encode())
            assert P.kseq_dump(p + ".gz").decode("latin-1") == v["records"], name + ".gz"


@pytest.mark.parametrize("name", ["Ill100.fx.gz", "CN_gDC.fa.gz", "Ill.nr.miss.fa.gz", "front_offset_bug.fa.gz", "poor_dr_ext.fa.gz"])
def test_bundled_dumps(P, name):
    """BASELINE.json configs[0]: bit-exact DR/spacer calls on the reference's bundled read sets."""
    path = os.path.join(checkers.REF_DATA, name)
    if not os.path.exists(path):
        pytest.skip("bundled read sets not staged (oracle/_ref/data is filled by `make -C oracle ref`)")
    want = gzip.open(os.path.join(G, "bundled", name + ".dump.gz")).read()
    sums = dict(l.split()[::-1] for l in open(os.path.join(G, "bundled", "MD5SUMS")).read().splitlines())
    assert hashlib.md5(want).hexdigest() == sums[name + ".dump"]
    got, _ = P.run_files([path])
    assert got.encode("latin-1") == want


def test_golden_counts():
    """SURVEY.md section 6 golden counts, read back from the committed dumps."""
    want = {"Ill100.fx.gz": (4324, 837, 140, 140, 101), "CN_gDC.fa.gz": (4740, 2761, 92, 97, 150),
            "Ill.nr.miss.fa.gz": (395, 10, 9, 9, 101), "front_offset_bug.fa.gz": (618, 54, 50, 51, 150),
            "poor_dr_ext.fa.gz": (8, 6, 4, 4, 1161)}
    for name, (_, hits, variants, phash, maxlen) in want.items():
        lines = gzip.open(os.path.join(G, "bundled", name + ".dump.gz")).read().decode("latin-1").split("\n")
        m = [l for l in lines if l.startswith("M\t")][0].split("\t")
        assert (int(m[1]), int(m[2]), int(m[3])) == (maxlen, hits, phash)
        p1 = [l for l in lines if l.startswith("R\t") and l.split("\t")[2] == "1"]
        assert len(p1) == hits
        assert len({l.split("\t")[1] for l in p1}) == variants


def test_kseq_damaged_archive_fixtures():
    """tests/golden/damaged_gz: damaged gzip / BGZF archives and a truncated one, with the record streams the REFERENCE reads
    from them (make_golden.py damaged): a failed gzread is taken for a short read by kstream (kseq.cpp:55-96)."""
    import base64
    import hashlib
    d = os.path.join(G, "damaged_gz")
    want = json.load(open(os.path.join(d, "expected.json")))
    assert len(want) == 6
    P = checkers.port()
    for name, w in want.items():
        got = P.kseq_dump(os.path.join(d, name))
        assert len(got) == w["records_len"] and got[-160:] == base64.b64decode(w["tail"]), name
        assert hashlib.md5(got).hexdigest() == w["records_md5"], name
