#!/usr/bin/env python
"""Compressed input on the feed path (the reference inflates inside its read loop, SeqUtils.cpp:100-125 +
kseq.cpp:55-69).  Needs no GPU.

  python tools/bench_gz.py [--reads N] [--repeat R]

Writes N synthetic 150 bp FASTQ records (Illumina-style names, mostly-'F' qualities) as `gzip -6` and as BGZF, then times
crass_b200.Batch.stream_file (inflate || parse, ranges of 128 MB) with
  * zlib's sequential read (CRASS_B200_GZ_SERIAL=1: what the library did before, and what any gzread-based reader gets),
  * the library's own DEFLATE decoder on the ordinary archive (one stream, one inflating thread),
  * BGZF blocks on 1 / 4 / 8 / all threads.
One JSON line per case: best-of-R seconds, MB/s of inflated data, M records/s, and the md5 of the record stream, which must be
the same in every case.  Not the headline bench.
"""
import argparse
import hashlib
import json
import os
import struct
import subprocess
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def write_bgzf(path, data, block=0xff00):
    """what bgzip writes: gzip members of at most 64 KB with the BC extra field (member size - 1) and an empty last one"""
    def member(chunk):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp +
                struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    with open(path, "wb") as fh:
        for i in range(0, len(data), block):
            fh.write(member(data[i:i + block]))
        fh.write(member(b""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    import crass_b200 as cb
    n = args.reads
    rng = np.random.default_rng(1)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, 150), dtype=np.uint8)]
    qual = np.frombuffer(b"FFFFFFFFFFFF::,#", dtype=np.uint8)[rng.integers(0, 16, size=(n, 150), dtype=np.uint8)]
    hdr = np.frombuffer(b"".join(b"@A00123:45:HXXXX:1:1101:%05d:%07d 1:N:0:ACGT\n" % (i % 99999, i // 7) for i in range(n)), dtype=np.uint8)
    nl = np.full((n, 1), 10, dtype=np.uint8)
    plus = np.full((n, 1), ord("+"), dtype=np.uint8)
    data = np.concatenate([hdr.reshape(n, len(hdr) // n), seq, nl, plus, nl, qual, nl], axis=1).tobytes()
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        plain = os.path.join(d, "reads.fq")
        with open(plain, "wb") as fh:
            fh.write(data)
        subprocess.run(["gzip", "-6", "-k", "-f", plain], check=True)
        gz = plain + ".gz"
        bgzf = os.path.join(d, "reads.bgzf.fq.gz")
        write_bgzf(bgzf, data)
        cores = os.cpu_count() or 1
        cases = [("gzip -6, zlib sequential read", gz, {"CRASS_B200_GZ_SERIAL": "1"}),
                 ("gzip -6, own DEFLATE decoder", gz, {}),
                 ("BGZF, zlib sequential read", bgzf, {"CRASS_B200_GZ_SERIAL": "1"})]
        for t in sorted({1, 4, 8, max(2, min(16, cores - 2))}):
            cases.append(("BGZF, blocks on %d threads" % t, bgzf, {"CRASS_B200_GZ_THREADS": str(t)}))
        for what, path, env in cases:
            for k in ("CRASS_B200_GZ_SERIAL", "CRASS_B200_GZ_THREADS"):
                os.environ.pop(k, None)
            os.environ.update(env)
            best, md5, nrec = 1e30, None, 0
            for _ in range(args.repeat):
                h = hashlib.md5()
                t0 = time.time()
                nrec = 0
                for b in cb.Batch.stream_file(path, 128 << 20):
                    nrec += len(b)
                best = min(best, time.time() - t0)
            for b in cb.Batch.stream_file(path, 128 << 20):            # (the checksum outside the clock)
                rs = b.record_stream()
                h.update(rs[:rs.rindex(b"#ret=")])
            md5 = h.hexdigest()
            print(json.dumps({"tool": "bench_gz", "case": what, "host_threads": cores, "reads": nrec, "archive_MB": round(os.path.getsize(path) / 1e6, 1),
                              "inflated_MB": round(len(data) / 1e6, 1), "seconds": round(best, 3), "MB_per_s_inflated": round(len(data) / best / 1e6),
                              "Mreads_per_s": round(nrec / best / 1e6, 2), "record_stream_md5": md5}), flush=True)


if __name__ == "__main__":
    main()
