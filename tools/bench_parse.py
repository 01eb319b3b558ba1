#!/usr/bin/env python
"""Feed-path timing: crass_b200_parse_file on a synthetic FASTA and FASTQ file of 150 bp reads, by
number of parser threads (CRASS_B200_PARSE_THREADS).  Needs no GPU.

  python tools/bench_parse.py [--reads N] [--threads 1,2,4,8,16] [--repeat R]

Prints one JSON line per (format, threads): best-of-R wall time, MB/s, Mreads/s and the md5 of the record stream, which
must not depend on the thread count (the pieces are checked against each other inside parse_file; this is the outside
check).  The reference's kseq_read loop runs at about 115 MB/s on one core.  Not the headline bench.
"""
import argparse
import hashlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def write_files(d, n, L=150):
    rng = np.random.default_rng(20242)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    paths = {}
    for fmt in ("fasta", "fastq"):
        p = os.path.join(d, "reads." + fmt)
        with open(p, "wb") as fh:
            for lo in range(0, n, 1 << 18):
                m = min(1 << 18, n - lo)
                names = np.char.add("r", np.char.zfill(np.arange(lo, lo + m).astype(str), 10)).astype("S11")
                seq = lut[rng.integers(0, 4, size=(m, L), dtype=np.uint8)]
                nl = np.full((m, 1), 10, dtype=np.uint8)
                head = np.full((m, 1), ord(">" if fmt == "fasta" else "@"), dtype=np.uint8)
                cols = [head, names.view(np.uint8).reshape(m, 11), nl, seq, nl]
                if fmt == "fastq":
                    qual = rng.integers(33, 74, size=(m, L), dtype=np.uint8)
                    cols += [np.full((m, 1), ord("+"), dtype=np.uint8), nl, qual, nl]
                fh.write(np.concatenate(cols, axis=1).tobytes())
        paths[fmt] = p
    return paths


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--threads", default="1,2,4,8,16")
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    import crass_b200 as cb

    with tempfile.TemporaryDirectory() as d:
        paths = write_files(d, args.reads)
        for fmt, p in paths.items():
            size = os.path.getsize(p)
            sums = set()
            for t in [int(x) for x in args.threads.split(",")]:
                os.environ["CRASS_B200_PARSE_THREADS"] = str(t)
                best = None
                for _ in range(args.repeat):
                    t0 = time.perf_counter()
                    b = cb.Batch.from_file(p)
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                md5 = hashlib.md5(b.record_stream()).hexdigest()
                sums.add(md5)
                print(json.dumps({"tool": "bench_parse", "format": fmt, "threads": t, "host_threads": os.cpu_count(),
                                  "reads": args.reads, "file_MB": round(size / 1e6, 1), "seconds": round(best, 4),
                                  "MB_per_s": round(size / 1e6 / best, 1), "Mreads_per_s": round(args.reads / 1e6 / best, 2),
                                  "record_stream_md5": md5}), flush=True)
                del b
            assert len(sums) == 1, "record stream depends on the thread count"


if __name__ == "__main__":
    main()
