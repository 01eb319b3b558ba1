#!/usr/bin/env python
"""The step between the phases on its own: createNonRedundantSet + matcher build from a token block on the device.

  python tools/bench_cluster.py [--list tools/data/dr_list_8x10M.txt.gz] [--iters 20]

The lists under tools/data are the merged DR lists of bench.py's config 2 at 1 and 8 ranks (tools/merged_dr_list.py made
them on a B200): 6 466 and 16 012 DR variants.  Prints one JSON line per list and mode: wall time of the call
(crass_b200_cluster_block_dev: kernels, the one host synchronisation, matcher tables), the pattern count, and whether the
pattern set equals the host passes' (it must)."""
import argparse
import gzip
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def block_from_list(drs, cap, stride):
    from crass_b200 import api
    blk = np.zeros(api.token_block_bytes(cap, stride), dtype=np.uint8)
    blk[:4] = np.frombuffer(np.uint32(len(drs)).tobytes(), dtype=np.uint8)
    recs = blk[16:16 + cap * stride].reshape(cap, stride)
    for t, d in enumerate(drs):
        recs[t, 0] = len(d)
        recs[t, 2:2 + len(d)] = np.frombuffer(d, dtype=np.uint8)
    recs[:len(drs), stride - 4:] = np.arange(len(drs), dtype=np.uint32).view(np.uint8).reshape(-1, 4)
    return blk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--list", action="append")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import torch
    import crass_b200 as cb
    from crass_b200 import api
    lists = args.list or [os.path.join(ROOT, "tools", "data", n) for n in ("dr_list_1x10M.txt.gz", "dr_list_8x10M.txt.gz")]
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    ctx = cb.Context(0)
    for path in lists:
        raw = gzip.open(path).read() if path.endswith(".gz") else open(path, "rb").read()
        drs = [d for d in raw.split(b"\n") if d]
        cap, stride = 16384 * (1 if len(drs) <= 16384 else 4), 64
        blk = block_from_list(drs, cap, stride)
        want = api.non_redundant_patterns(b"".join(d + b"\n" for d in drs), 6)
        d_blk = torch.from_numpy(blk).to(dev)
        for mode in ("device", "device-passes"):
            if mode == "device":
                os.environ.pop("CRASS_B200_CLUSTER", None)
            else:
                os.environ["CRASS_B200_CLUSTER"] = mode
            times = []
            for it in range(args.iters + 3):
                s.synchronize()
                t0 = time.perf_counter()
                ac, cnt, fl = ctx.cluster_block_dev(d_blk, cap, stride, 6, s.cuda_stream)
                s.synchronize()
                if it >= 3:
                    times.append((time.perf_counter() - t0) * 1e3)
            same = ac.pattern_text() == want
            print(json.dumps({"list": os.path.basename(path), "variants": len(drs), "cap": cap, "mode": mode, "patterns": ac.num_patterns,
                              "ms_median": float(np.median(times)), "ms_min": float(np.min(times)), "identical_to_host_passes": bool(same)}), flush=True)
            assert same
    os.environ.pop("CRASS_B200_CLUSTER", None)
    ctx.close()


if __name__ == "__main__":
    main()
