"""K1 time on a 1/N row shard of the C3 database, measured on ONE GPU (rank 0's shard of an N-way split) — the per-GPU
share of a sharded step without paying for an N-GPU box.  Prints one JSON line per N with the library's own CUDA-event
K1 time, the ideal (the N = 1 time / N) and the merged keys' checksum for a cheap cross-version comparison.

    python tools/k1_shard_sweep.py [--frames 64] [--shards 1,2,4,8] [--reps 10]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tod_b200 import DescriptorMatcher, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--keypoints", type=int, default=2000)
    ap.add_argument("--objects", type=int, default=100)
    ap.add_argument("--rows", type=int, default=10000)
    ap.add_argument("--k", type=int, default=2)
    ap.add_argument("--shards", default="1,2,4,8")
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    descs, points = synth.make_db(a.objects, a.rows, seed=11)
    nq = a.frames * a.keypoints
    q = synth.make_queries(descs, nq, seed=12)[0]
    dq = torch.from_numpy(q).cuda()
    base = None
    for n in [int(x) for x in a.shards.split(",")]:
        m = DescriptorMatcher(k=a.k, radius=0, device=0, shard_rank=0, shard_count=n)
        for i, (d, p) in enumerate(zip(descs, points)):
            m.add_object("o%d" % i, d, p)
        m.train()
        m.reserve(nq)
        keys = torch.empty((nq, a.k), dtype=torch.int32, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ts = []
        for r in range(a.reps + 2):
            flush.zero_()
            torch.cuda.synchronize()
            m.knn_keys_device(dq.data_ptr(), nq, keys.data_ptr())
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(m.last_k1_ms)
        ms = float(np.median(ts))
        if base is None:
            base = ms * n
        csum = int(keys.to(torch.int64).sum().item())
        print(json.dumps({"shards": n, "shard_rows": m.shard_rows, "nq": nq, "k1_ms": round(ms, 4),
                          "k1_ms_min": round(min(ts), 4), "ideal_ms": round(base / n, 4),
                          "efficiency": round(base / n / ms, 4), "keys_checksum": csum}), flush=True)
        m.close()


if __name__ == "__main__":
    main()
