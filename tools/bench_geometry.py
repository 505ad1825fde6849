#!/usr/bin/env python
"""Stage numbers for the geometry half of the hot path on one B200 (BASELINE configs C4 / C5):

  K2  tod_fill_adjacency      achieved GB/s = algorithmic bytes (32 n in + 2 n W 4 out per cluster) / kernel time
  K3  tod_score_hypotheses    achieved GB/s = algorithmic bytes (3 W 4 rows + 2 W 4 masks + 72 + 16 in, 52 out) / time
  GuessGenerator.process      whole cell (K2 + rounds of K3 + host replay / gate / refinement), wall clock, host buffers

next to the reference's own geometry code (oracle/_ref: src/common compiled unmodified) timed on ONE host core — the
reference is single-threaded — on a bounded sample of the same inputs.  Kernel times are CUDA events recorded by the
library around the launch (tod_last_stage_ms / tod_guess_last_stats).  Prints one JSON object; diagnostic tool.
usage: python tools/bench_geometry.py [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tod_b200 import GuessGenerator, capi, fill_adjacency, score_hypotheses, synth  # noqa: E402


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p)).get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def clusters(sizes, inlier_fraction, seed):
    qs, ts, ps = [], [], []
    for i, n in enumerate(sizes):
        q, t, px, _, _ = synth.make_cluster(n, inlier_fraction, seed=seed + i)
        qs.append(q); ts.append(t); ps.append(px)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    return off, np.concatenate(qs), np.concatenate(ts), np.concatenate(ps)


def bench_k2(name, sizes, inlier_fraction, reps=5):
    lib = capi.load()
    off, q, t, px = clusters(sizes, inlier_fraction, 9000)
    spans = np.full(len(sizes), 0.25, np.float32)
    ms = []
    for _ in range(reps):
        P, S, mo = fill_adjacency(off, q, t, px, spans, 0.01)
        ms.append(float(lib.tod_last_stage_ms()))
    ms = float(np.median(ms[1:]))
    words = sum(int(n) * capi.adjacency_row_words(int(n)) for n in sizes)
    alg = 32.0 * sum(sizes) + 2.0 * words * 4.0
    pairs = sum(n * (n - 1) // 2 for n in sizes)
    peak, src = hbm_peak()
    return {"case": name, "clusters": len(sizes), "n": int(sizes[0]), "kernel_ms": ms, "algorithmic_bytes": alg,
            "achieved_gbs": alg / (ms * 1e-3) / 1e9, "peak_gbs": peak, "frac": alg / (ms * 1e-3) / 1e9 / peak,
            "peak_source": src, "pair_tests": pairs, "gpairs_per_s": pairs / (ms * 1e-3) / 1e9,
            "edges_physical": int(np.bitwise_count(P).sum()), "edges_sample": int(np.bitwise_count(S).sum())}, (q, t, px, P, S, off)


def triangles(Sdense, rng, H):
    n = Sdense.shape[0]
    out = []
    tries = 0
    while len(out) < H and tries < 50 * H:
        tries += 1
        a = int(rng.integers(0, n))
        na = np.nonzero(Sdense[a])[0]
        if len(na) == 0:
            continue
        b = int(na[rng.integers(0, len(na))])
        nab = np.nonzero(Sdense[a] & Sdense[b])[0]
        if len(nab) == 0:
            continue
        out.append((int(nab[rng.integers(0, len(nab))]), b, a))
    return np.array(out, np.uint32).reshape(-1, 3)


def bench_k3(n, inlier_fraction, H, reps=5):
    lib = capi.load()
    q, t, px, _, _ = synth.make_cluster(n, inlier_fraction, seed=9100)
    P, S, _ = fill_adjacency([0, n], q, t, px, [0.25], 0.01)
    W = capi.adjacency_row_words(n)
    Sd = np.unpackbits(S.reshape(n, W).view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)
    rng = np.random.default_rng(1)
    base = triangles(Sd, rng, 2048)
    tri = np.ascontiguousarray(np.tile(base, ((H + len(base) - 1) // len(base), 1))[:H])
    V = np.packbits(np.concatenate([np.ones(n, bool), np.zeros(W * 32 - n, bool)]), bitorder="little").view("<u4")
    out = {}
    for mode, thr in (("reference_faithful_inf", float("inf")), ("finite_threshold_2e", 0.02)):
        ms = []
        for _ in range(reps):
            counts, R, T = score_hypotheses(q, t, P, V, tri, threshold=thr)
            ms.append(float(lib.tod_last_stage_ms()))
        ms = float(np.median(ms[1:]))
        cand = float(counts.mean())
        per_h = 3.0 * W * 4 + 2.0 * W * 4 + 72 + 16 + 52 + (0.0 if thr == float("inf") else 24.0 * cand)
        peak, src = hbm_peak()
        out[mode] = {"n": n, "hypotheses": int(H), "kernel_ms": ms, "hyp_per_s": H / (ms * 1e-3),
                     "algorithmic_bytes_per_hypothesis": per_h, "achieved_gbs": per_h * H / (ms * 1e-3) / 1e9,
                     "peak_gbs": peak, "frac": per_h * H / (ms * 1e-3) / 1e9 / peak, "peak_source": src,
                     "mean_count": cand}
    return out


def bench_guess(n_objects, n_per_object, inlier_fraction, iters, ref_objects=3, host_threads=0):
    g = synth.make_guess_inputs(n_objects, n_per_object, inlier_fraction, seed=synth.BASE_SEED + 5, k=1, height=960,
                                width=1280)
    gg = GuessGenerator(min_inliers=15, n_ransac_iterations=iters, sensor_error=0.01, seed=11,
                        host_threads=host_threads)
    res = None
    wall = []
    for _ in range(3):
        t0 = time.perf_counter()
        res = gg.process(g["keypoints_xy"], g["cloud"], g["matches"], g["counts"], g["points3d"], g["spans"],
                         max_poses=32 * n_objects)
        wall.append(time.perf_counter() - t0)
    st = gg.last_stats()
    out = {"objects": n_objects, "correspondences_per_object": n_per_object, "inlier_fraction": inlier_fraction,
           "n_ransac_iterations": iters, "wall_ms": 1e3 * float(np.median(wall[1:])), "k2_ms": st["k2_ms"],
           "k3_ms": st["k3_ms"], "hypotheses_scored": st["n_hypotheses"], "rounds": st["n_rounds"],
           "host_threads": host_threads or min(16, os.cpu_count() or 1), "host_ms": st["host_ms"],
           "gate_calls": st["gate_calls"], "gate_proved_empty": st["gate_proved_empty"],
           "gate_core_rejects": st["gate_core_rejects"],
           "gate_thread_ms": st["gate_thread_ms"],
           "poses": int(len(res["pose_results"])),
           "objects_recovered": int(len(set(int(p["object_index"]) for p in res["pose_results"])))}
    try:
        from oracle import ref
        if ref.available() and ref_objects > 0:
            sub = synth.make_guess_inputs(ref_objects, n_per_object, inlier_fraction, seed=synth.BASE_SEED + 5, k=1,
                                          height=960, width=1280)
            t0 = time.perf_counter()
            exp = ref.process(sub["keypoints_xy"], sub["cloud"], sub["matches"], sub["counts"], sub["points3d"],
                              sub["spans"], 15, iters, 0.01, seed=11)
            dt = time.perf_counter() - t0
            out["cpu_reference"] = {"kind": "reference", "cores": 1, "sample": "%d of the %d objects" % (
                ref_objects, n_objects), "ms_per_object": 1e3 * dt / ref_objects, "poses": len(exp),
                "extrapolated_ms_all_objects": 1e3 * dt / ref_objects * n_objects}
            out["speedup_vs_cpu_reference_1core"] = out["cpu_reference"]["extrapolated_ms_all_objects"] / out["wall_ms"]
    except Exception as e:  # the checker is optional here
        out["cpu_reference"] = {"unavailable": str(e)[:200]}
    return out


def main():
    res = {"gpu": "B200", "note": "kernel_ms = CUDA events around the launch; see tools/bench_geometry.py"}
    res["k2"] = []
    for name, sizes, frac in (("C5: 100 objects x 2000 correspondences, 90% outliers", [2000] * 100, 0.1),
                              ("C4-like: 64 frames x 16 objects x 256 correspondences", [256] * 1024, 0.5),
                              ("worst case: one cluster of 20480 correspondences", [20480], 0.3)):
        r, _ = bench_k2(name, sizes, frac)
        res["k2"].append(r)
        print(json.dumps(r), file=sys.stderr)
    res["k3"] = bench_k3(2000, 0.1, 409600)
    print(json.dumps(res["k3"]), file=sys.stderr)
    res["guess_c5"] = bench_guess(100, 2000, 0.1, 4096)
    print(json.dumps(res["guess_c5"]), file=sys.stderr)
    res["guess_c5_1thread"] = bench_guess(100, 2000, 0.1, 4096, ref_objects=0, host_threads=1)
    print(json.dumps(res["guess_c5_1thread"]), file=sys.stderr)
    res["guess_conf"] = bench_guess(10, 400, 0.5, 2500, ref_objects=10)
    print(json.dumps(res["guess_conf"]), file=sys.stderr)
    s = json.dumps(res, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(s)
    print(s)


if __name__ == "__main__":
    main()
