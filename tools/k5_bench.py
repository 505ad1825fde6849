"""K5 alone (tod_gate_search_device): time of one launch over a batch of random graphs of one size class — where the
thread-per-search kernel spends its time (the slowest single search bounds a launch).
usage: python tools/k5_bench.py"""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tod_b200 import capi  # noqa: E402


def batch(rng, count, lo, hi, p):
    nv, offs, edges = [], [0], []
    for _ in range(count):
        n = int(rng.integers(lo, hi + 1))
        a = np.triu(rng.random((n, n)) < p, 1)
        e = np.argwhere(a).astype(np.int32)
        nv.append(n)
        edges.append(e)
        offs.append(offs[-1] + e.shape[0])
    return np.array(nv, np.int32), np.array(offs, np.int32), np.ascontiguousarray(np.concatenate(edges), np.int32)


def main():
    lib = capi.load()
    rng = np.random.default_rng(1)
    for lo, hi, count in [(24, 64, 4000), (65, 128, 2000), (129, 256, 600), (129, 256, 8)]:
        for p in (0.35, 0.6, 0.85):
            nv, off, ed = batch(rng, count, lo, hi, p)
            res = np.zeros(count, np.int32)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                capi.check(lib.tod_gate_search_device(0, count, capi._ptr(nv), capi._ptr(off), capi._ptr(ed), capi._ptr(res)))
                ts.append(time.perf_counter() - t0)
            steps = []
            for g in range(min(count, 200)):
                st = ctypes.c_int32(0)
                e = ed[off[g]:off[g + 1]]
                lib.tod_clique_gate_small(int(nv[g]), capi._ptr(np.ascontiguousarray(e)), e.shape[0], 100000, ctypes.byref(st))
                steps.append(st.value)
            print(json.dumps({"vertices": [lo, hi], "p": p, "graphs": count, "call_ms_incl_copies": round(1e3 * min(ts), 3),
                              "passes": int((res == 1).sum()), "to_host": int((res < 0).sum()),
                              "steps_mean": float(np.mean(steps)), "steps_max": int(np.max(steps))}), flush=True)


if __name__ == "__main__":
    main()
