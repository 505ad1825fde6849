#!/usr/bin/env python
"""BASELINE config C4 on one B200: a batch of 64 synthetic 1280x960 RGB-D frames, 4096 keypoints each, against the
100-object / 1M-descriptor DB — DescriptorMatcher.process (K1, k = 5, radius 35 as in conf/detection.ork) followed by
the batched GuessGenerator (K2 + K3 rounds + host replay / gate / refinement), host buffers in and out, wall clock.
The CPU reference beside it: cv2 BFMatcher on a sample of one frame's keypoints (all cores) and the reference's own
geometry code (oracle/_ref, 1 core — it is single-threaded) on one frame's matches.  Prints one JSON object.
usage: python tools/bench_pipeline.py [out.json] [n_frames]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tod_b200 import DescriptorMatcher, GuessGenerator, synth  # noqa: E402


def main():
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    n_kp, H, W, K, RADIUS, ITERS = 4096, 960, 1280, 5, 35, 2500
    descs, points = synth.make_db(100, 10000, seed=synth.BASE_SEED + 2)
    rng = np.random.default_rng(4)
    frames = []
    for f in range(n_frames):
        vis = sorted(int(x) for x in rng.choice(100, 4, replace=False))
        frames.append(synth.make_frame(descs, points, vis, n_kp, height=H, width=W, seed=synth.BASE_SEED + 400 + f))
    q_all = np.ascontiguousarray(np.concatenate([f["descriptors"] for f in frames]))
    clouds = np.stack([f["cloud"] for f in frames])
    m = DescriptorMatcher(k=K, radius=RADIUS)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("object_%03d" % i, d, p)
    m.train()
    spans = m.spans_by_index
    gg = GuessGenerator(min_inliers=15, n_ransac_iterations=ITERS, sensor_error=0.01, seed=9)
    t_match, t_guess = [], []
    res = out = None
    for rep in range(4):
        t0 = time.perf_counter()
        out = m.process(q_all)
        t1 = time.perf_counter()
        res = gg.process_batch([f["keypoints_xy"] for f in frames], clouds, out["matches"], out["counts"],
                               out["matches_3d"], spans, max_poses=64 * n_frames)
        t2 = time.perf_counter()
        if rep:
            t_match.append(t1 - t0)
            t_guess.append(t2 - t1)
    st = gg.last_stats()
    # ---- streaming mode: the matcher works on batch i + 1 (GPU) while the guess generator finishes batch i (host) ----
    import threading
    n_batches = 6
    t0 = time.perf_counter()
    prev = [None]

    def guess_job(o):
        prev[0] = gg.process_batch([f["keypoints_xy"] for f in frames], clouds, o["matches"], o["counts"],
                                   o["matches_3d"], spans, max_poses=64 * n_frames)
    worker = None
    for b in range(n_batches):
        o = m.process(q_all)                      # ctypes releases the GIL: runs beside the previous batch's guess
        o = {k_: (v.copy() if isinstance(v, np.ndarray) else v) for k_, v in o.items()}
        if worker is not None:
            worker.join()
        worker = threading.Thread(target=guess_job, args=(o,))
        worker.start()
    worker.join()
    t_stream = time.perf_counter() - t0
    # planted poses recovered?
    want = got = 0
    for f, r in zip(frames, res):
        for o, (R, T) in f["poses"].items():
            want += 1
            for p in r["pose_results"]:
                if int(p["object_index"]) == o and np.abs(p["R"].reshape(3, 3) - R).max() < 0.02 and \
                        np.abs(p["T"] - T).max() < 0.01:
                    got += 1
                    break
    tm, tg = float(np.median(t_match)), float(np.median(t_guess))
    d = {"config": "C4: %d frames x %d keypoints, %dx%d clouds, 1M-descriptor DB (100 objects), k=%d radius=%d, "
                   "n_ransac_iterations=%d" % (n_frames, n_kp, W, H, K, RADIUS, ITERS),
         "frames_per_s": n_frames / (tm + tg), "matcher_ms_per_batch": 1e3 * tm, "guess_ms_per_batch": 1e3 * tg,
         "frames_per_s_streaming": n_batches * n_frames / t_stream,
         "streaming_note": "matcher of batch i+1 overlapped with the guess generator of batch i (two host threads)",
         "k1_ms": m.last_k1_ms, "k1_kernel": m.last_kernel, "k2_ms": st["k2_ms"], "k3_ms": st["k3_ms"],
         "hypotheses": st["n_hypotheses"], "rounds": st["n_rounds"], "guess_host_ms": st["host_ms"],
         "gate_calls": st["gate_calls"], "poses_found": int(sum(len(r["pose_results"]) for r in res)),
         "planted_objects": want, "planted_recovered": got,
         "matches_per_frame": float(out["counts"].sum()) / n_frames}
    # ---- CPU reference on a bounded sample ----
    try:
        import cv2
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        bf.add([np.ascontiguousarray(x) for x in descs])
        ns = 256
        bf.knnMatch(np.ascontiguousarray(frames[0]["descriptors"][:32]), K)
        t0 = time.perf_counter()
        bf.knnMatch(np.ascontiguousarray(frames[0]["descriptors"][:ns]), K)
        dt = time.perf_counter() - t0
        cpu_match_ms = 1e3 * dt * n_kp / ns
        d["cpu_reference"] = {"matcher": {"kind": "reference", "what": "cv2 %s BFMatcher.knnMatch" % cv2.__version__,
                                          "cores": cv2.getNumThreads(), "sample": "%d of %d keypoints of one frame" % (ns, n_kp),
                                          "ms_per_frame_extrapolated": cpu_match_ms}}
        from oracle import ref
        if ref.available():
            o0 = 0
            f0 = frames[0]
            t0 = time.perf_counter()
            exp = ref.process(f0["keypoints_xy"], f0["cloud"], out["matches"][:n_kp], out["counts"][:n_kp],
                              out["matches_3d"][:n_kp], spans, 15, ITERS, 0.01, seed=9)
            dt = time.perf_counter() - t0
            d["cpu_reference"]["geometry"] = {"kind": "reference", "what": "src/common compiled unmodified (oracle/_ref)",
                                              "cores": 1, "sample": "frame 0 of the batch", "ms_per_frame": 1e3 * dt,
                                              "poses": len(exp)}
            d["cpu_reference"]["frames_per_s"] = 1e3 / (cpu_match_ms + 1e3 * dt)
            d["speedup_vs_cpu_reference"] = d["frames_per_s"] / d["cpu_reference"]["frames_per_s"]
    except Exception as e:
        d["cpu_reference"] = {"unavailable": str(e)[:200]}
    s = json.dumps(d, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(s)
    print(s)


if __name__ == "__main__":
    main()
