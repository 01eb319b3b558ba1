#!/usr/bin/env python
"""File in -> containers out through the engine (crass_b200_engine_run_files) on a config-2 style FASTA in tmpfs, with the
stage trace of the library on stderr (CRASS_B200_TRACE=1).   python tools/e2e_trace.py [--reads N] [--repeat R] [--devices 0,0]"""
import argparse
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--fastq", action="store_true", help="write the reads as FASTQ (one fixed quality string)")
    ap.add_argument("--files", type=int, default=1, help="split the reads over this many files (paired-end style run)")
    args = ap.parse_args()
    os.environ.setdefault("CRASS_B200_TRACE", "1")
    import crass_b200 as cb
    from crass_b200 import synth
    genome, _, _ = synth.make_genome(20242)
    n = args.reads
    bases, offs = synth.sample_fixed(genome, n, 150, 21242)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        path = os.path.join(d, "reads.fa")
        hdr = np.frombuffer(b"".join(b">r%010d\n" % i for i in range(n)), dtype=np.uint8).reshape(n, 13).copy()
        nl = np.full((n, 1), 10, dtype=np.uint8)
        cols = [hdr, bases.reshape(n, 150), nl]
        if args.fastq:
            hdr[:, 0] = ord("@")
            qual = np.frombuffer(bytes(33 + (7 * k) % 41 for k in range(150)), dtype=np.uint8)
            cols += [np.full((n, 1), ord("+"), dtype=np.uint8), nl, np.broadcast_to(qual, (n, 150)), nl]
        rows = np.concatenate(cols, axis=1)
        paths = []
        for f in range(args.files):
            paths.append(os.path.join(d, "reads%d.fx" % f))
            rows[n * f // args.files:n * (f + 1) // args.files].tofile(paths[-1])
        del rows
        eng = cb.Engine(tuple(int(x) for x in args.devices.split(",")))
        for it in range(args.repeat):
            t0 = time.time()
            res, ml = eng.run_files(paths)
            dt = time.time() - t0
            print("run %d: %.3f s, %.2f M reads/s, %d found reads, %d tokens, stages %s" % (it, dt, n / dt / 1e6, res.num_reads, res.num_tokens, eng.stage_ms()), flush=True)
            del res
        eng.close()


if __name__ == "__main__":
    main()
