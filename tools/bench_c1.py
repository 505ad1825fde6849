#!/usr/bin/env python
"""BASELINE config C1 — the reference's own CPU-runnable case: one synthetic 640x480 RGB-D frame, a 1-object DB of 5000
ORB descriptors, 5000 keypoints, the `.ork` parameters (LSH search, k = 5, radius 35, 2500 RANSAC iterations, 8
inliers).  Reference arm on the host: cv2 FlannBasedMatcher + LSH (what DescriptorMatcher.cpp:175-181 builds) then the
reference's own geometry code (oracle/_ref), single frame latency.  Our arm: DescriptorMatcher.process +
GuessGenerator.process through the C-ABI with host buffers, single frame latency.  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tod_b200 import DescriptorMatcher, GuessGenerator, capi, synth  # noqa: E402


def main():
    descs, points = synth.make_db(1, 5000, seed=synth.BASE_SEED)
    fr = synth.make_frame(descs, points, [0], 5000, seed=synth.BASE_SEED + 9)
    m = DescriptorMatcher(search_json_params='{"type": "LSH", "key_size": 16, "multi_probe_level": 1, "n_tables": 10, '
                                             '"radius": 35, "ratio": 0.8}')
    m.add_object("object_0", descs[0], points[0])
    m.train()
    g = GuessGenerator(min_inliers=8, n_ransac_iterations=2500, sensor_error=0.01, seed=5)
    tm, tg, res = [], [], None
    for _ in range(8):
        t0 = time.perf_counter()
        out = m.process(fr["descriptors"])
        t1 = time.perf_counter()
        res = g.process(fr["keypoints_xy"], fr["cloud"], out["matches"], out["counts"], out["matches_3d"],
                        m.spans_by_index)
        t2 = time.perf_counter()
        tm.append(t1 - t0)
        tg.append(t2 - t1)
    R, T = fr["poses"][0]
    ok = any(np.abs(p["R"].reshape(3, 3) - R).max() < 0.02 and np.abs(p["T"] - T).max() < 0.01
             for p in res["pose_results"])
    d = {"config": "C1: 640x480 frame, 1 object x 5000 descriptors, 5000 keypoints, .ork parameters",
         "ours": {"matcher_ms": 1e3 * float(np.median(tm[2:])), "guess_ms": 1e3 * float(np.median(tg[2:])),
                  "frame_ms": 1e3 * float(np.median(np.array(tm[2:]) + np.array(tg[2:]))),
                  "k1_kernel": m.last_kernel, "k1_ms": m.last_k1_ms, "poses": int(len(res["pose_results"])),
                  "planted_pose_recovered": bool(ok), "matches": int(out["counts"].sum()),
                  "guess_stats": {k: v for k, v in g.last_stats().items() if k in ("host_ms", "gate_shape", "k2_ms",
                                  "k3_ms", "n_hypotheses", "n_rounds", "gate_thread_ms")}}}
    try:
        import cv2
        from oracle import ref
        lsh = cv2.FlannBasedMatcher(dict(algorithm=6, table_number=10, key_size=16, multi_probe_level=1), dict())
        lsh.add([np.ascontiguousarray(descs[0])])
        lsh.train()
        t0 = time.perf_counter()
        raw = lsh.knnMatch(fr["descriptors"], 5)
        t_lsh = time.perf_counter() - t0
        nq = fr["descriptors"].shape[0]
        mt = np.zeros((nq, 5), capi.MATCH_DTYPE)
        cnt = np.zeros(nq, np.int32)
        p3 = np.zeros((nq, 5, 3), np.float32)
        for q, lst in enumerate(raw):                       # radius cut, DescriptorMatcher.cpp:212-220
            for j, dm in enumerate(lst[:5]):
                if dm.distance > 35:
                    break
                mt[q, j] = (q, dm.trainIdx, dm.imgIdx, dm.distance)
                p3[q, j] = points[dm.imgIdx][dm.trainIdx]
                cnt[q] = j + 1
        t0 = time.perf_counter()
        exp = ref.process(fr["keypoints_xy"], fr["cloud"], mt, cnt, p3, m.spans_by_index, 8, 2500, 0.01, seed=5) \
            if ref.available() else None
        t_geo = time.perf_counter() - t0
        d["cpu_reference"] = {"matcher": "cv2 %s FlannBasedMatcher LSH(10, 16, 1), knnMatch(5) + radius 35" % cv2.__version__,
                              "matcher_ms": 1e3 * t_lsh, "geometry": "src/common compiled unmodified (oracle/_ref), 1 core",
                              "geometry_ms": 1e3 * t_geo if exp is not None else None,
                              "frame_ms": 1e3 * (t_lsh + (t_geo if exp is not None else 0.0)),
                              "poses": len(exp) if exp is not None else None, "matches": int(cnt.sum())}
        d["speedup_vs_cpu_reference"] = d["cpu_reference"]["frame_ms"] / d["ours"]["frame_ms"]
    except Exception as e:
        d["cpu_reference"] = {"unavailable": str(e)[:200]}
    s = json.dumps(d, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(s)
    print(s)


if __name__ == "__main__":
    main()
