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


def run_distributed(n_frames):
    """torchrun mode (one process per GPU): the DB is sharded over the ranks for K1 (NCCL all-gather of the packed keys,
    merge on every rank), then the batch's frames are split over the ranks for the guess generator — frames are
    independent — and the pose counts are gathered.  Throughput = frames / max-over-ranks wall time."""
    import torch
    import torch.distributed as dist
    from tod_b200 import capi
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_kp, H, W, K, RADIUS, ITERS = 4096, 960, 1280, 5, 35, 2500
    descs, points = synth.make_db(100, 10000, seed=synth.BASE_SEED + 2)
    rng = np.random.default_rng(4)
    vis_all = [sorted(int(x) for x in rng.choice(100, 4, replace=False)) for _ in range(n_frames)]
    f0, f1 = rank * n_frames // world, (rank + 1) * n_frames // world      # this rank's frames for the geometry half
    # every rank needs every frame's descriptors (queries are replicated) but only its own frames' clouds / keypoints
    frames = [synth.make_frame(descs, points, vis_all[f], n_kp, height=H, width=W, seed=synth.BASE_SEED + 400 + f)
              for f in range(n_frames)]
    q_all = np.ascontiguousarray(np.concatenate([f["descriptors"] for f in frames]))
    clouds = np.stack([frames[f]["cloud"] for f in range(f0, f1)])
    kps = [frames[f]["keypoints_xy"] for f in range(f0, f1)]
    planted = [frames[f]["poses"] for f in range(f0, f1)]
    del frames
    m = DescriptorMatcher(k=K, radius=RADIUS, device=local, shard_rank=rank, shard_count=world)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("object_%03d" % i, d, p)
    m.train()
    spans = m.spans_by_index
    gg = GuessGenerator(min_inliers=15, n_ransac_iterations=ITERS, sensor_error=0.01, seed=9, device=local,
                        host_threads=max(1, min(16, (os.cpu_count() or 16) // world)))
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    nq = q_all.shape[0]
    q_host = torch.from_numpy(q_all).pin_memory()
    q_dev = torch.empty((nq, 32), dtype=torch.uint8, device=dev)
    keys = torch.empty((nq, K), dtype=torch.int32, device=dev)
    keys_all = torch.empty((world, nq, K), dtype=torch.int32, device=dev)
    matches = torch.empty((nq, K, 4), dtype=torch.int32, device=dev)
    counts = torch.empty((nq,), dtype=torch.int32, device=dev)
    pts3d = torch.empty((nq, K, 3), dtype=torch.float32, device=dev)
    lo, hi = f0 * n_kp, f1 * n_kp
    m_host = torch.empty((hi - lo, K, 4), dtype=torch.int32).pin_memory()
    c_host = torch.empty((hi - lo,), dtype=torch.int32).pin_memory()
    p_host = torch.empty((hi - lo, K, 3), dtype=torch.float32).pin_memory()
    times, res = [], None
    for rep in range(4):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        q_dev.copy_(q_host, non_blocking=True)
        m.knn_keys_device(q_dev.data_ptr(), nq, keys.data_ptr(), sptr)
        dist.all_gather_into_tensor(keys_all.view(-1), keys.view(-1))
        m.merge_device(keys_all.data_ptr(), world, nq, matches.data_ptr(), counts.data_ptr(), pts3d.data_ptr(), sptr)
        m_host.copy_(matches[lo:hi], non_blocking=True)      # only this rank's frames come back to the host
        c_host.copy_(counts[lo:hi], non_blocking=True)
        p_host.copy_(pts3d[lo:hi], non_blocking=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        mm = m_host.numpy().view(capi.MATCH_DTYPE).reshape(hi - lo, K).copy()
        mm["queryIdx"] = np.where(mm["queryIdx"] >= 0, mm["queryIdx"] - lo, mm["queryIdx"])
        res = gg.process_batch(kps, clouds, mm, c_host.numpy(), p_host.numpy(), spans, max_poses=64 * (f1 - f0))
        t2 = time.perf_counter()
        t = torch.tensor([t2 - t0, t1 - t0, t2 - t1], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep:
            times.append(t.cpu().numpy())
    want = got = 0
    for pl, r in zip(planted, res):
        for o, (R, T) in pl.items():
            want += 1
            got += any(int(p["object_index"]) == o and np.abs(p["R"].reshape(3, 3) - R).max() < 0.02 and
                       np.abs(p["T"] - T).max() < 0.01 for p in r["pose_results"])
    tot = torch.tensor([want, got, sum(len(r["pose_results"]) for r in res)], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    tm = np.median(np.array(times), axis=0)
    d = {"config": "C4 on %d GPUs: %d frames x %d keypoints, %dx%d clouds, 1M-descriptor DB sharded by rows for K1, "
                   "frames split over the ranks for the guess generator" % (world, n_frames, n_kp, W, H),
         "n_gpus": world, "frames_per_s": n_frames / float(tm[0]), "matcher_ms_per_batch_max_rank": 1e3 * float(tm[1]),
         "guess_ms_per_batch_max_rank": 1e3 * float(tm[2]), "planted_objects": int(tot[0]),
         "planted_recovered": int(tot[1]), "poses_found": int(tot[2]), "host_cores": os.cpu_count()}
    if rank == 0:
        s = json.dumps(d, indent=1)
        if len(sys.argv) > 1:
            open(sys.argv[1], "w").write(s)
        print(s)
    dist.destroy_process_group()


def main():
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_distributed(n_frames)
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
    # host buffers the way a capture pipeline would hold them: pinned (two sets, for the streaming mode below)
    import torch
    from tod_b200 import capi
    nq_all = q_all.shape[0]
    q_pin = torch.from_numpy(q_all).pin_memory()

    def pinned_out():
        mt = torch.empty((nq_all, K, 4), dtype=torch.int32).pin_memory()
        ct = torch.empty((nq_all,), dtype=torch.int32).pin_memory()
        pt = torch.empty((nq_all, K, 3), dtype=torch.float32).pin_memory()
        return {"matches": mt.numpy().view(capi.MATCH_DTYPE).reshape(nq_all, K), "counts": ct.numpy(),
                "matches_3d": pt.numpy(), "_keep": (mt, ct, pt)}
    bufs = [pinned_out(), pinned_out()]
    t_match, t_guess = [], []
    res = out = None
    for rep in range(4):
        t0 = time.perf_counter()
        out = m.process(q_pin.numpy(), out=bufs[0])
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
    n_batches = 8
    t0 = time.perf_counter()
    prev = [None]

    def guess_job(o):
        prev[0] = gg.process_batch([f["keypoints_xy"] for f in frames], clouds, o["matches"], o["counts"],
                                   o["matches_3d"], spans, max_poses=64 * n_frames)
    worker = None
    for b in range(n_batches):
        o = m.process(q_pin.numpy(), out=bufs[b & 1])   # ctypes releases the GIL: runs beside the previous batch's guess
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
