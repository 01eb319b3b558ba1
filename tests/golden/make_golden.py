#!/usr/bin/env python
"""Regenerates the committed golden fixtures by RUNNING THE COMPILED REFERENCE
(oracle/_ref/libcrass_ref.so, built from /root/reference by `make -C oracle ref`).

Outputs (all under tests/golden/):
  bundled/<name>.dump.gz     "crass-dump v1" of searchFile -> createNonRedundantSet -> findSingletons
                             on each read set the reference ships in test/ (default options)
  bundled/MD5SUMS            md5 of the uncompressed dumps
  search_core_vectors.json   seeded per-read vectors: read, params -> (found, start/stops, repeat length)
  edit_distance_vectors.json seeded string pairs -> (distance, similarity as float32 hex)
  lowlexi_vectors.json       seeded (read, start/stops) -> (DR, flag, mirrored start/stops)
  ac_vectors.json            seeded pattern sets + texts -> first match (end, length) or null
  kseq_vectors.json          hand-made FASTA/FASTQ edge-case files -> record stream seen by searchFile
  update_start_stops_vectors.json  seeded (read, start/stops, front offset, DR, lowSpacerSize) -> new start/stops
                             (ReadHolder::updateStartStops), and smithWaterman calls -> (ok, start, end,
                             lengths of the two strings, find/rfind of the second in the DR);
                             `python make_golden.py uss` regenerates this file alone

  consensus_vectors.json     ksw_align calls (nt4 codes of q / t -> score, te, qe, score2, te2, tb, qb) and DR groups through
                             Aligner::setMasterDR / alignSlave / generateConsensus -> placements, strands, zone, consensus,
                             conservation bits, coverage; `python make_golden.py consensus` regenerates this file alone

catch_vectors.json is NOT generated: it restates the known-answer tests of the reference's own
src/test/test_libcrispr.cpp by hand (each entry cites its line range).
"""
import gzip
import hashlib
import json
import os
import random
import struct
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import checkers  # noqa: E402
import fuzzgen  # noqa: E402

KSEQ_CASES = {
    "plain_fasta": ">r1 first comment\nACGTACGT\nACGT\n>r2\nGGGG\n>r3\tx y\nTTTT",
    "fastq_then_fasta_stale_qual": "@q1 c1\nACGTAC\n+\nIIIIII\n>f1\nACGTACGTAA\n>f2 c2\nAC\n@q2\nAAAA\n+q2\n!!!!\n",
    "crlf": ">r1 c\r\nACGT\r\nAC\r\n>r2\r\nGG\r\n",
    "multiline_fastq": "@m1\nACGT\nACGT\n+\nIIII\nIIII\n@m2\nAC\n+\nII\n",
    "qual_with_at": "@a1\nACGTA\n+\n@@@@@\n@a2\nCC\n+\n@I\n",
    "truncated_qual": "@t1\nACGTACGT\n+\nIIII\n@t2\nAC\n+\nII\n",
    "leading_garbage": "garbage line\n\n>r1\nAC GT\n>r2\nA+C\nGG\n",
    "plus_in_fasta": ">p1\nACGT\n+\nIIII\n>p2\nTT\n",
    "no_trailing_newline": ">r1\nACGT\n>r2 last",
    "empty_comment_then_stale": ">r1 \nAAAA\n>r2 real\nCCCC\n>r3\nGGGG\n",
    "lowercase_and_iupac": ">l1\nacgtNNRYacgt\n>l2\nACGTUacgtu\n",
}
# kstream's 4096-byte buffer: inputs whose size is a multiple of it and whose last byte is a header character yield one more
# (empty) record than the same bytes one byte longer or shorter
_fa = "".join(">q%03d\nACGTTGCAACGT\n" % i for i in range(100))
KSEQ_CASES["exactly_4096_ending_in_gt"] = _fa + ">z c\n" + "A" * (4096 - len(_fa) - 5 - 2) + "\n>"
_fq = "".join("@f%03d\nACGTAC\n+\nIIIIII\n" % i for i in range(300))
_m = (8192 - len(_fq) - 3 - 3 - 2) // 2
KSEQ_CASES["exactly_8192_fastq_then_at"] = _fq + "@p\n" + "C" * _m + "\n+\n" + "I" * _m + "\n@" + ("" if (8192 - len(_fq)) % 2 == 0 else "@")
KSEQ_CASES["one_more_than_4096_ending_in_gt"] = KSEQ_CASES["exactly_4096_ending_in_gt"][:-2] + "A\n>"
assert len(KSEQ_CASES["exactly_4096_ending_in_gt"]) == 4096 and len(KSEQ_CASES["exactly_8192_fastq_then_at"]) == 8192 and len(KSEQ_CASES["one_more_than_4096_ending_in_gt"]) == 4097


def gen_update_start_stops(R):
    rng = random.Random(20240)
    uss, sw = [], []
    while len(uss) < 400:
        seq, ss, front, dr = fuzzgen.uss_case(rng)
        low = rng.choice([26, 26, 26, 20, 35])
        if checkers.port().update_start_stops(seq, ss, front, dr, low)[0] == -3:
            continue                                 # shifted start past the read: the reference reads out of bounds
        st, out = R.update_start_stops(seq, ss, front, dr, low)
        assert st == 0
        uss.append(dict(seq=seq.decode(), ss=ss, front=front, dr=dr.decode(), low_spacer=low, ss_out=out))
    for _ in range(400):
        seq, ss, front, dr = fuzzgen.uss_case(rng)
        start = rng.randint(0, len(seq) - 1)
        length = rng.randint(1, len(seq) - start)
        sim = 0.85 if rng.random() < 0.8 else 0.0
        sw.append(dict(a=seq.decode(), b=dr.decode(), start=start, len=length, similarity=sim,
                       out=list(R.smith_waterman(seq, dr, start, length, sim))))
    json.dump(dict(update_start_stops=uss, smith_waterman=sw), open(os.path.join(HERE, "update_start_stops_vectors.json"), "w"))


def gen_consensus(R):
    """ksw_align calls and whole DR groups through the reference's Aligner (`python make_golden.py consensus`)."""
    rng = random.Random(20241)
    nt = bytes.maketrans(b"ACGTN", bytes([0, 1, 2, 3, 4]))
    ksw = []
    while len(ksw) < 600:
        t = fuzzgen.rand_seq(rng, rng.randint(20, 60))
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
            q = fuzzgen.rand_seq(rng, rng.randint(1, 50))
        if q:
            ksw.append(dict(q=q.decode(), t=t.decode(), out=list(R.ksw_align(q.translate(nt), t.translate(nt)))))
    groups = []
    for it in range(40):
        case = fuzzgen.consensus_case(rng, read_len=(70, 90), alphabet=b"ACGTN" if it % 4 == 0 else b"ACGT")
        out = R.consensus_group(case)
        assert out["status"] == 0
        cov = out.pop("coverage")
        out["coverage_md5"] = hashlib.md5(struct.pack("<%di" % len(cov), *cov)).hexdigest()
        if it < 4:
            out["coverage"] = cov
        out["consensus"] = out["consensus"].decode()
        groups.append(dict(reads=[[r[0].decode(), r[1], r[2]] for r in case["reads"]], drs=[d.decode() for d in case["drs"]],
                           array_len=case["array_len"], out=out))
    json.dump(dict(ksw_align=ksw, groups=groups), open(os.path.join(HERE, "consensus_vectors.json"), "w"))


def gen_damaged(R):
    """Small damaged archives (ordinary gzip with a flipped bit, BGZF with a damaged block, a truncated one) and the record stream
    the REFERENCE's kseq loop reads from each: what a failed gzread leaves behind (kseq.cpp:55-96) pinned without the compiled
    reference.  The archives are committed as they are (a few KB each)."""
    import base64
    import zlib
    sys.path.insert(0, os.path.dirname(HERE))
    from test_host_logic import _write_bgzf
    rng = random.Random(515)
    parts = []
    for k in range(900):
        seq = fuzzgen.rand_seq(rng, rng.randint(30, 150)).decode()
        if rng.random() < 0.5:
            parts.append("@r%d c%d\n%s\n+\n%s\n" % (k, k, seq, "".join(rng.choice("!5I@>F") for _ in seq)))
        else:
            parts.append(">r%d\n%s\n" % (k, seq))
    text = "".join(parts).encode()
    plain = gzip.compress(text, compresslevel=6, mtime=0)
    out = {}
    d = os.path.join(HERE, "damaged_gz")
    os.makedirs(d, exist_ok=True)
    for it in range(6):
        name = "bad%d.fx.gz" % it
        p = os.path.join(d, name)
        if it < 3:
            at = rng.randint(len(plain) // 8, len(plain) - 64)
            data = plain[:at] + bytes([plain[at] ^ (1 << rng.randint(0, 7))]) + plain[at + 1:]
            with open(p, "wb") as fh:
                fh.write(data)
        elif it < 5:
            _write_bgzf(p, text, block=9000, corrupt_block=rng.randint(1, len(text) // 9000 - 1))
        else:
            with open(p, "wb") as fh:
                fh.write(plain[:len(plain) * 2 // 3])
        want = R.kseq_dump(p)
        out[name] = dict(records_md5=hashlib.md5(want).hexdigest(), records_len=len(want), tail=base64.b64encode(want[-160:]).decode())
    json.dump(out, open(os.path.join(d, "expected.json"), "w"), indent=1)


def main():
    R = checkers.ref()
    if sys.argv[1:] == ["damaged"]:
        gen_damaged(R)
        return
    if sys.argv[1:] == ["uss"]:
        gen_update_start_stops(R)
        return
    if sys.argv[1:] == ["consensus"]:
        gen_consensus(R)
        return
    if sys.argv[1:] == ["kseq"]:
        gen_kseq(R)
        return
    os.makedirs(os.path.join(HERE, "bundled"), exist_ok=True)
    sums = []
    for f in sorted(os.listdir(checkers.REF_DATA)):
        if not f.endswith(".gz"):
            continue
        dump, _ = R.run_files([os.path.join(checkers.REF_DATA, f)])
        raw = dump.encode("latin-1")
        sums.append("%s  %s.dump" % (hashlib.md5(raw).hexdigest(), f))
        with gzip.GzipFile(os.path.join(HERE, "bundled", f + ".dump.gz"), "wb", mtime=0) as g:
            g.write(raw)
    with open(os.path.join(HERE, "bundled", "MD5SUMS"), "w") as g:
        g.write("\n".join(sums) + "\n")

    rng = random.Random(20240)
    vec = []
    while len(vec) < 1500:
        s = fuzzgen.fuzz_read(rng)
        prm = None
        if rng.random() < 0.25:
            prm = dict(window=rng.choice([6, 7, 8, 9]), min_repeats=rng.choice([2, 3, 4]), low_dr=rng.choice([23, 20, 30]),
                       high_dr=rng.choice([47, 60]), low_spacer=rng.choice([26, 20, 10]), high_spacer=rng.choice([50, 60]))
        found, ss, rl = R.search_core(s, prm)
        if found != 1 and rng.random() < 0.5 and len(vec) > 200:
            continue
        vec.append(dict(seq=s.decode("latin-1"), params=prm, found=found, ss=ss, replen=rl))
    json.dump(vec, open(os.path.join(HERE, "search_core_vectors.json"), "w"), indent=0)

    vec = []
    for _ in range(2000):
        a = fuzzgen.rand_seq(rng, rng.randint(0, 55), b"ACGTN")
        b = fuzzgen.mutate(rng, a, 0.2) if rng.random() < 0.6 else fuzzgen.rand_seq(rng, rng.randint(0, 55))
        if rng.random() < 0.3:
            b = b[rng.randint(0, 3):]
        vec.append([a.decode(), b.decode(), R.edit_distance(a, b), struct.pack(">f", R.similarity(a, b)).hex()])
    json.dump(vec, open(os.path.join(HERE, "edit_distance_vectors.json"), "w"))

    vec = []
    for _ in range(600):
        L = rng.randint(60, 300)
        s = fuzzgen.rand_seq(rng, L, b"ACGTNRYacgtU")
        n = rng.choice([1, 2, 2, 3, 4])
        pts = sorted(rng.sample(range(0, L), 2 * n))
        if rng.random() < 0.3:
            pts[0] = 0
        if rng.random() < 0.3:
            pts[-1] = L - 1
        dr, low, ss2, seq2 = R.dr_lowlexi(s, pts)
        vec.append(dict(seq=s.decode(), ss=pts, dr=dr.decode(), lowlexi=low, ss_out=ss2, seq_out=seq2.decode()))
    json.dump(vec, open(os.path.join(HERE, "lowlexi_vectors.json"), "w"))

    vec = []
    for _ in range(40):
        pats = fuzzgen.dr_like_patterns(rng, rng.randint(1, 60))
        if rng.random() < 0.3:
            pats += [p[rng.randint(0, 5):len(p) - rng.randint(0, 5)] for p in pats[:10]]
        h = R.ac_create(pats)
        texts = []
        for _k in range(60):
            t = fuzzgen.rand_seq(rng, rng.randint(0, 200), b"ACGTN" if rng.random() < 0.2 else b"ACGT")
            if rng.random() < 0.6 and len(t) > 60:
                p = rng.choice(pats)
                pos = rng.randint(0, len(t) - 1)
                t = (t[:pos] + p + t[pos:])[:len(t)]
            texts.append([t.decode(), R.ac_first_match(h, t)])
        R.ac_destroy(h)
        vec.append(dict(patterns=[p.decode() for p in pats], texts=texts))
    json.dump(vec, open(os.path.join(HERE, "ac_vectors.json"), "w"))

    gen_kseq(R)
    gen_update_start_stops(R)
    print("golden fixtures regenerated from", checkers.REF_SO)


def gen_kseq(R):
    vec = {}
    with tempfile.TemporaryDirectory() as d:
        for name, content in KSEQ_CASES.items():
            for gz in (False, True):
                p = os.path.join(d, name + (".gz" if gz else ".fx"))
                if gz:
                    with gzip.open(p, "wb") as g:
                        g.write(content.encode())
                else:
                    open(p, "wb").write(content.encode())
                out = R.kseq_dump(p).decode("latin-1")
                if name in vec:
                    assert vec[name]["records"] == out
                vec[name] = dict(content=content, records=out)
    json.dump(vec, open(os.path.join(HERE, "kseq_vectors.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
