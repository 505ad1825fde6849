#!/usr/bin/env python
"""bench.py — headline benchmark of the TOD detection hot path on B200.

Metric (BASELINE.json): frames/sec at 2k keypoints x 1M-descriptor DB (100 objects x 10k ORB descriptors), exact
Hamming k-NN k=2, on 1/2/4/8 B200; Hamming Gcmp/s vs the matching roofline.

A "step" = one batch of `--frames` synthetic frames (2000 query descriptors each) through the hot path:
  value : frames/s with the queries already resident in HBM (K1 k-NN -> [when the DB is sharded: top-k reduction fused
          with the exchange of the packed keys over peer memory, or ncclAllGather] -> merge/radius/decode/3-D gather),
          timed with CUDA events, max over ranks;
  e2e   : the same metric through the reference-facing call with HOST buffers (pinned): H2D of the descriptors,
          DescriptorMatcher.process, D2H of matches / counts / matches_3d, wall clock around synchronous calls
          (N > 1: streamed, every rank uploads the batch and reads back the frames it owns).
Beside the headline (N = 1): `parity` of the timed result against live cv2, `roofline` of K1, `cpu_baseline` (+ the
reference's LSH matcher), `e2e_pipeline` = BASELINE config C4 through both cells (matcher + batched guess generator,
back to back and streamed) with `stages` (K2 / K3 rooflines, host split), `stages_c5` = the RANSAC stress
configuration, `feature_stage` = GPU ORB + DepthTo3d beside cv2.ORB.
N > 1 shards the DB rows over the ranks (strong scaling: the 1M-descriptor DB is fixed).

`--impl reference` times the reference's own CPU implementation of the path on the host cores: OpenCV's
cv::BFMatcher(NORM_HAMMING) (the exact matcher north_star names; through cv2) on all 2000 keypoints of a frame per
step, plus the matcher the reference ships (FlannBasedMatcher + LSH) as `reference_lsh`.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec at 2k kpts x 1M-desc DB (exact Hamming kNN k=2)"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="frames per step (batch)")
    ap.add_argument("--keypoints", type=int, default=2000)
    ap.add_argument("--objects", type=int, default=100)
    ap.add_argument("--rows", type=int, default=10000, help="descriptors per object")
    ap.add_argument("--k", type=int, default=2)
    ap.add_argument("--radius", type=int, default=0)
    ap.add_argument("--kernel", default="auto", choices=["auto", "popc", "mma"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-queries", type=int, default=0,
                    help="keypoints per CPU-baseline repetition (0 = all keypoints of a frame)")
    ap.add_argument("--no-lsh", action="store_true", help="skip the FlannBasedMatcher+LSH leg of the CPU baseline")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the C4 (matcher + guess generator) leg")
    ap.add_argument("--pipeline-frames", type=int, default=64)
    ap.add_argument("--c5-objects", type=int, default=100, help="objects in the C5 RANSAC-stress leg (0 = skip)")
    ap.add_argument("--exchange", default="library", choices=["library", "nccl", "torch"],
                    help="N > 1: 'library' (default) = the exchange inside the library, over peer memory where the GPUs "
                         "can map each other (reduce+push kernel over NVLink) else ncclAllGather; 'nccl' = force the "
                         "in-library ncclAllGather; 'torch' = device-timed loop only, the key all-gather moved out of "
                         "the library (stage calls + torch.distributed), for A/B comparison")
    ap.add_argument("--trace", action="store_true", help="progress lines on stderr (debugging a multi-rank run)")
    return ap.parse_args()


def workload(args):
    from tod_b200 import synth
    descs, points = synth.make_db(args.objects, args.rows, seed=synth.BASE_SEED + 2)
    frames = []
    for f in range(args.frames):
        q, _, _ = synth.make_queries(descs, args.keypoints, seed=synth.BASE_SEED + 102 + f)
        frames.append(q)
    return descs, points, np.ascontiguousarray(np.concatenate(frames))


def config_dict(args, world):
    return {"workload": "C3: %d keypoints/frame x %d-descriptor DB (%d objects x %d), exact Hamming kNN k=%d, "
                        "radius %d" % (args.keypoints, args.objects * args.rows, args.objects, args.rows, args.k,
                                       args.radius),
            "frames_per_step": args.frames, "db_sharding": "rows/%d" % world,
            "l2": "flushed between timed steps (256 MiB memset outside the timed events)"}


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def load_int_peaks():
    """Measured INT-pipe peaks from tools/microbench (committed under profiles/), else the documented fallback."""
    p = os.path.join(ROOT, "profiles", "int_peaks.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["source"] = "profiles/int_peaks.json (tools/microbench on this pool's B200)"
        return d
    return {"xor_popc_gcmp": 148 * 16 * 1.965 / 8.0 * 1.0, "source": "fallback: 16 POPC/clk/SM x 148 SMs x 1.965 GHz / 8"}


def load_measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_knn(descs, queries, k, threads=None):
    """The reference's exact matcher on the host: cv2.BFMatcher(NORM_HAMMING) if importable, else the oracle C port.
    Returns (seconds, kind, cores, what, (trainIdx, imgIdx, distance) arrays)."""
    try:
        import cv2
        if threads:
            cv2.setNumThreads(threads)
        m = cv2.BFMatcher(cv2.NORM_HAMMING)
        m.add([np.ascontiguousarray(d) for d in descs])
        t0 = time.perf_counter()
        res = m.knnMatch(np.ascontiguousarray(queries), k)
        dt = time.perf_counter() - t0
        assert len(res) == queries.shape[0]
        trn = np.array([[x.trainIdx for x in r] + [-1] * (k - len(r)) for r in res], np.int32).reshape(-1, k)
        img = np.array([[x.imgIdx for x in r] + [-1] * (k - len(r)) for r in res], np.int32).reshape(-1, k)
        dist = np.array([[x.distance for x in r] + [0] * (k - len(r)) for r in res], np.float32).reshape(-1, k)
        return dt, "reference", cv2.getNumThreads(), "cv2 %s BFMatcher(NORM_HAMMING).knnMatch" % cv2.__version__, \
            (trn, img, dist)
    except ImportError:
        from oracle import hamming_knn as hk
        t0 = time.perf_counter()
        em, ec = hk.knn_c(queries, descs, k)
        dt = time.perf_counter() - t0
        return dt, "port", os.cpu_count(), "oracle/hamming_knn.c (OpenMP)", (em["trainIdx"], em["imgIdx"],
                                                                              em["distance"])


def lsh_reference(descs, queries, exact, radius=35):
    """The matcher the reference actually ships (DescriptorMatcher.cpp:175-181 with conf/detection.ork:32-39):
    cv::FlannBasedMatcher + LshIndexParams(10 tables, key 16, multi-probe 1), knnMatch(5) + radius cut, on the host —
    timed on one frame, with its recall of the exact within-radius matches `exact` = (trainIdx, imgIdx, distance)."""
    try:
        import cv2
    except ImportError:
        return {"unavailable": "cv2 not importable"}
    fl = cv2.FlannBasedMatcher(dict(algorithm=6, table_number=10, key_size=16, multi_probe_level=1), dict())
    fl.add([np.ascontiguousarray(d) for d in descs])
    t0 = time.perf_counter()
    fl.train()
    t_train = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = fl.knnMatch(np.ascontiguousarray(queries), 5)
    t_knn = time.perf_counter() - t0
    got = set()
    for qi, r in enumerate(res):
        for x in r[:5]:
            if x.distance > radius:          # DescriptorMatcher.cpp:212-220
                break
            got.add((qi, x.imgIdx, x.trainIdx))
    trn, img, dist = exact
    want = set((qi, int(img[qi, j]), int(trn[qi, j])) for qi in range(trn.shape[0]) for j in range(trn.shape[1])
               if trn[qi, j] >= 0 and dist[qi, j] <= radius)
    return {"what": "cv2 FlannBasedMatcher LSH(table_number 10, key_size 16, multi_probe_level 1), knnMatch(5) + "
                    "radius %d (conf/detection.ork:32-39)" % radius,
            "frames_per_s": 1.0 / t_knn, "knn_ms_per_frame": 1e3 * t_knn, "train_ms_once": 1e3 * t_train,
            "cores": 1, "sample": "one whole frame (%d keypoints)" % queries.shape[0],
            "recall_of_exact_within_radius": (len(got & want) / float(len(want))) if want else None,
            "exact_within_radius": len(want), "lsh_within_radius": len(got)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    descs, points, queries = workload(args)
    nsamp = args.keypoints if args.cpu_sample_queries <= 0 else min(args.cpu_sample_queries, args.keypoints)
    times = []
    info = None
    exact = None
    for s in range(args.warmup + args.steps):
        f = s % args.frames
        q = queries[f * args.keypoints: f * args.keypoints + nsamp]
        dt, kind, cores, what, ex = cpu_reference_knn(descs, q, args.k)
        info = (kind, cores, what)
        if s >= args.warmup:
            times.append(dt)
        if f == 0:
            exact = ex
    total = float(sum(times))
    frames = args.steps * nsamp / float(args.keypoints)
    value = frames / total
    sample = "%d of the %d keypoints of one frame per step vs the full %d-descriptor DB; %s" % (
        nsamp, args.keypoints, args.objects * args.rows, info[2])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info[1], "kind": info[0], "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gcmp_per_s": value * args.keypoints * args.objects * args.rows / 1e9}
    if not args.no_lsh and nsamp == args.keypoints:
        # second leg, reported beside the headline: the matcher the reference ships (LSH), with its recall
        k5 = cpu_reference_knn(descs, queries[:args.keypoints], 5)[4] if args.k != 5 else exact
        line["reference_lsh"] = lsh_reference(descs, queries[:args.keypoints], k5)
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
def parity_sample(args, m_np, c_np, p_np, descs, points, queries):
    """Bit-exact check of the result of the TIMED configuration (the last end-to-end step) on a fixed sample of its
    queries — first and last query tile, plus queries spread over every frame — against cv::BFMatcher(NORM_HAMMING)
    (DescriptorMatcher.cpp:211-220 semantics) and the matches_3d gather (:232-244)."""
    from oracle import hamming_knn as hk
    nqt, k = queries.shape[0], args.k
    rng = np.random.default_rng(12345)
    idx = set(range(0, min(128, nqt))) | set(range(max(0, nqt - 128), nqt))
    per_frame = max(1, 320 // max(1, args.frames))
    for f in range(args.frames):
        lo = f * args.keypoints
        idx |= set(int(x) for x in lo + rng.choice(args.keypoints, min(per_frame, args.keypoints), replace=False))
    idx = np.array(sorted(idx), np.int64)
    _, kind, _, what, (trn, img, dist) = cpu_reference_knn(descs, queries[idx], k)
    cnt = (trn >= 0).sum(axis=1).astype(np.int32)
    if args.radius > 0:
        keep = np.cumprod((dist <= args.radius) & (trn >= 0), axis=1).astype(bool)
        cnt = keep.sum(axis=1).astype(np.int32)
    mask = np.arange(k)[None, :] < cnt[:, None]
    got_m, got_c, got_p = m_np[idx], c_np[idx], p_np[idx]
    bad = (got_c != cnt)
    for f, exp in (("trainIdx", trn), ("imgIdx", img), ("distance", dist)):
        bad |= ((got_m[f] != exp) & mask).any(axis=1)
    bad |= ((got_m["queryIdx"] != idx[:, None]) & mask).any(axis=1)
    em = np.zeros((idx.shape[0], k), got_m.dtype)
    em["trainIdx"], em["imgIdx"] = np.where(mask, trn, 0), np.where(mask, img, 0)
    e3 = hk.gather_points3d(em, cnt, points)
    bad |= ((got_p != e3) & mask[:, :, None]).any(axis=(1, 2))
    return {"checked": int(idx.shape[0]), "mismatches": int(bad.sum()), "against": what, "kind": kind,
            "fields": "counts, trainIdx, imgIdx, distance, queryIdx, matches_3d",
            "where": "result of the last timed end-to-end step (merged result when sharded): first and last 128 "
                     "queries + %d per frame" % per_frame}


# what ncu says really bounds the geometry kernels (profiles/r2_geometry_ncu.json, one `--set full` capture each): their
# matrices stay in the 126 MB L2, so the HBM fraction north_star asks for is small by construction
NCU_NOTES = {
    "k2_adjacency_kernel": {"issue_slots_busy_pct": 76.8, "dram_throughput_pct": 0.18, "dram_bytes_per_launch": 2418432,
                            "bound_by": "instruction issue (about 89 warp instructions per 32 pair tests, no FMA "
                                        "contraction allowed); the output stays in L2",
                            "source": "profiles/r2_geometry_ncu.json"},
    "k3_score_kernel": {"issue_slots_busy_pct": 18.0, "dram_throughput_pct": 0.63, "dram_bytes_per_launch": 1630464,
                        "bound_by": "launch latency at C4 (16 384 hypotheses, 32 us); L2 bandwidth at C5",
                        "source": "profiles/r2_geometry_ncu.json"},
}


def hbm_roofline(kernel, alg_bytes, ms, hbm_peak, hbm_src, extra=None):
    gbs = alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    r = {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
         "peak_source": hbm_src, "kernel_ms": ms, "algorithmic_bytes": alg_bytes, "traffic": None}
    if kernel in NCU_NOTES:
        r["ncu"] = NCU_NOTES[kernel]
    if extra:
        r.update(extra)
    return r


def pipeline_leg(args, descs, points, hbm_peak, hbm_src):
    """BASELINE config C4 through the reference-facing calls with HOST buffers: a batch of 64 synthetic 1280x960
    RGB-D frames x 4096 keypoints, DescriptorMatcher.process (k = 5, radius 35 as in conf/detection.ork) then the
    batched GuessGenerator.process — the whole north_star path.  Returns the `e2e_pipeline` and `stages` objects."""
    import torch
    from tod_b200 import DescriptorMatcher, GuessGenerator, capi, synth
    n_frames, n_kp, H, W, K, RADIUS, ITERS = args.pipeline_frames, 4096, 960, 1280, 5, 35, 2500
    rng = np.random.default_rng(4)
    frames = []
    for f in range(n_frames):
        vis = sorted(int(x) for x in rng.choice(len(descs), min(4, len(descs)), replace=False))
        frames.append(synth.make_frame(descs, points, vis, n_kp, height=H, width=W, seed=synth.BASE_SEED + 400 + f))
    q_all = np.ascontiguousarray(np.concatenate([f["descriptors"] for f in frames]))
    clouds = np.stack([f["cloud"] for f in frames])
    # keypoints as the reference's feature cell hands them over: one cv::KeyPoint-shaped record each
    kps = GuessGenerator.pack_keypoints([f["keypoints_xy"] for f in frames])
    m5 = DescriptorMatcher(search_json_params='{"type": "LSH", "radius": %d, "ratio": 0.8, "n_tables": 10, '
                                              '"key_size": 16, "multi_probe_level": 1}' % RADIUS)
    assert m5.k == K
    for i, (d, p) in enumerate(zip(descs, points)):
        m5.add_object("object_%03d" % i, d, p)
    m5.train()
    nq = q_all.shape[0]
    m5.reserve(nq)
    spans = m5.spans_by_index
    gg = GuessGenerator(min_inliers=15, n_ransac_iterations=ITERS, sensor_error=0.01, seed=9)
    q_pin = torch.from_numpy(q_all).pin_memory()
    mt = torch.empty((nq, K, 4), dtype=torch.int32).pin_memory()
    ct = torch.empty((nq,), dtype=torch.int32).pin_memory()
    pt = torch.empty((nq, K, 3), dtype=torch.float32).pin_memory()
    out = {"matches": mt.numpy().view(capi.MATCH_DTYPE).reshape(nq, K), "counts": ct.numpy(), "matches_3d": pt.numpy()}
    t_m, t_g, k1, stats, res = [], [], [], None, None
    for rep in range(4):
        t0 = time.perf_counter()
        o = m5.process(q_pin.numpy(), out=out)
        t1 = time.perf_counter()
        res = gg.process_batch(kps, clouds, o["matches"], o["counts"], o["matches_3d"], spans, max_poses=64 * n_frames)
        t2 = time.perf_counter()
        if rep:
            t_m.append(t1 - t0)
            t_g.append(t2 - t1)
            k1.append(m5.last_k1_ms)
            stats = gg.last_stats()
    # the same two calls as a stream of batches: the matcher of batch i + 1 (GPU-bound) runs on a second host thread
    # while the guess generator of batch i (host-bound) runs on this one, two sets of pinned buffers
    streamed = None
    try:
        import queue
        import threading
        n_batches = 6
        outs = [out, None]
        mt2 = torch.empty((nq, K, 4), dtype=torch.int32).pin_memory()
        ct2 = torch.empty((nq,), dtype=torch.int32).pin_memory()
        pt2 = torch.empty((nq, K, 3), dtype=torch.float32).pin_memory()
        outs[1] = {"matches": mt2.numpy().view(capi.MATCH_DTYPE).reshape(nq, K), "counts": ct2.numpy(),
                   "matches_3d": pt2.numpy()}
        free_q, full_q = queue.Queue(), queue.Queue()
        free_q.put(0)
        free_q.put(1)
        err = []

        def produce():
            try:
                for _ in range(n_batches):
                    b = free_q.get()
                    m5.process(q_pin.numpy(), out=outs[b])
                    full_q.put(b)
            except Exception as e:          # surfaced by the consumer
                err.append(e)
                full_q.put(None)

        th = threading.Thread(target=produce)
        t0 = time.perf_counter()
        th.start()
        n_poses = 0
        for _ in range(n_batches):
            b = full_q.get()
            if b is None:
                raise err[0]
            o = outs[b]
            r = gg.process_batch(kps, clouds, o["matches"], o["counts"], o["matches_3d"], spans, max_poses=64 * n_frames)
            n_poses += sum(len(x["pose_results"]) for x in r)
            free_q.put(b)
        th.join()
        dt = time.perf_counter() - t0
        streamed = {"value": n_batches * n_frames / dt, "unit": UNIT, "batches": n_batches,
                    "ms_per_batch": 1e3 * dt / n_batches, "poses_per_batch": n_poses / n_batches,
                    "scope": "the same two C-ABI calls on a stream of batches: DescriptorMatcher.process of batch i + 1 on "
                             "a second host thread while GuessGenerator.process(batch) of batch i runs (two sets of "
                             "pinned buffers); includes the pipeline fill of the first batch"}
    except Exception as e:
        streamed = {"unavailable": str(e)[:200]}
    want = got = 0
    for f, r in zip(frames, res):
        for o_, (R, T) in f["poses"].items():
            want += 1
            got += any(int(p["object_index"]) == o_ and np.abs(p["R"].reshape(3, 3) - R).max() < 0.02 and
                       np.abs(p["T"] - T).max() < 0.01 for p in r["pose_results"])
    tm, tg = float(np.median(t_m)), float(np.median(t_g))
    k1_ms = float(np.median(k1))
    e2e = {"value": n_frames / (tm + tg), "unit": UNIT,
           "workload": "C4: %d frames x %d keypoints, %dx%d clouds, %d-descriptor DB, k=%d radius=%d, "
                       "n_ransac_iterations=%d, min_inliers=15" % (n_frames, n_kp, W, H, sum(d.shape[0] for d in descs),
                                                                   K, RADIUS, ITERS),
           "scope": "DescriptorMatcher.process + GuessGenerator.process(batch) through the C-ABI, pinned host buffers, "
                    "sequential (no overlap between the two cells)",
           "matcher_ms_per_batch": 1e3 * tm, "guess_ms_per_batch": 1e3 * tg,
           "h2d_bytes_per_step": int(nq * 32), "d2h_bytes_per_step": int(nq * K * 16 + nq * 4 + nq * K * 12),
           "poses_found": int(sum(len(r["pose_results"]) for r in res)), "planted_objects": want,
           "planted_recovered": got, "matches_per_frame": float(out["counts"].sum()) / n_frames,
           "streamed": streamed}
    stages = {"k1": {"kernel_ms": k1_ms, "gcmp_per_s": nq * float(m5.num_descriptors) / (k1_ms * 1e-3) / 1e9},
              "k2": hbm_roofline("k2_adjacency_kernel", stats["k2_bytes"], stats["k2_ms"], hbm_peak, hbm_src,
                                 {"clusters": stats["n_clusters"], "correspondences": stats["n_correspondences"]}),
              "k3": hbm_roofline("k3_score_kernel", stats["k3_bytes"], stats["k3_ms"], hbm_peak, hbm_src,
                                 {"hypotheses": stats["n_hypotheses"], "rounds": stats["n_rounds"]}),
              "guess_host_ms": stats["host_ms"], "gate_calls": stats["gate_calls"], "gate_shape": stats["gate_shape"]}
    # the reference's CPU pipeline on a bounded sample of the same batch: exact matcher on one whole frame (all host
    # cores) + the reference's own geometry code (oracle/_ref, single-threaded like the reference) on that frame
    cpu = None
    try:
        dt, kind, cores, what, _ = cpu_reference_knn(descs, frames[0]["descriptors"], K)
        cpu = {"matcher": {"kind": kind, "what": what, "cores": cores, "ms_per_frame": 1e3 * dt,
                           "sample": "frame 0 (all %d keypoints)" % n_kp}}
        from oracle import ref
        if ref.available():
            t0 = time.perf_counter()
            exp = ref.process(frames[0]["keypoints_xy"], frames[0]["cloud"], out["matches"][:n_kp],
                              out["counts"][:n_kp], out["matches_3d"][:n_kp], spans, 15, ITERS, 0.01, seed=9)
            dg = time.perf_counter() - t0
            cpu["geometry"] = {"kind": "reference", "what": "src/common compiled unmodified (oracle/_ref)", "cores": 1,
                               "sample": "frame 0", "ms_per_frame": 1e3 * dg, "poses": len(exp)}
            cpu["value"] = 1.0 / (dt + dg)
            cpu["unit"] = UNIT
            # parity of the pipeline on the frame the CPU reference just processed
            mine = res[0]
            same = len(mine["pose_results"]) == len(exp)
            if same:
                for p, inl, (eo, eR, eT, einl) in zip(mine["pose_results"], mine["inliers"], exp):
                    same = same and int(p["object_index"]) == eo and list(inl) == list(einl) and \
                        float(np.abs(p["R"].reshape(3, 3) - eR).max()) < 1e-4 and float(np.abs(p["T"] - eT).max()) < 1e-4
            e2e["parity"] = {"checked": "frame 0: poses, inlier keypoint sets, R/T within 1e-4 vs oracle/_ref",
                             "poses": len(exp), "mismatches": 0 if same else 1}
    except Exception as e:                                      # the checker is optional on a box without cv2 / _ref
        cpu = {"unavailable": str(e)[:200]}
    e2e["cpu_baseline"] = cpu
    m5.close()
    gg.close()
    return e2e, stages


def feature_leg(args):
    """The stage in front of the hot path: cv::ORB(5000, 1.2, 3) on one synthetic 1280 x 960
    frame and DepthTo3d, on the GPU through the C-ABI with host buffers, beside cv2.ORB on the host cores, with the
    parity of this very frame (keypoint set and descriptors)."""
    from tod_b200 import FeatureDescriptor, depth_to_3d, synth
    h, w, nf = 960, 1280, 5000
    img = synth.make_textured_image(h, w, seed=synth.BASE_SEED + 41, n_shapes=1500)
    zf, _ = synth.make_depth_image(h, w, seed=synth.BASE_SEED + 42)
    K = np.array([[1050.0, 0, 639.5], [0, 1050.0, 479.5], [0, 0, 1]], np.float32)
    fd = FeatureDescriptor(n_features=nf)
    t_orb, t_d3 = [], []
    import torch
    cloud = torch.empty((h, w, 3), dtype=torch.float32).pin_memory().numpy()   # a frame loop reuses one pinned cloud
    for _ in range(6):
        t0 = time.perf_counter()
        kp, desc = fd.process(img)
        t1 = time.perf_counter()
        depth_to_3d(zf, K, out=cloud)
        t2 = time.perf_counter()
        t_orb.append(t1 - t0)
        t_d3.append(t2 - t1)
    fd.close()
    out = {"workload": "one %dx%d frame, ORB n_features %d, n_levels 3, scale_factor 1.2 (conf/detection.ork:23-31)" % (w, h, nf),
           "orb_ms_per_frame": 1e3 * float(np.median(t_orb[1:])), "keypoints": int(kp.shape[0]),
           "depth_to_3d_ms_per_frame": 1e3 * float(np.median(t_d3[1:])),
           "scope": "tod_orb_detect_and_compute / tod_depth_to_3d with host buffers (H2D of the frame, D2H of keypoints, "
                    "descriptors and the point image inside; the 14.7 MB point image lands in a pinned buffer)"}
    try:
        import cv2
        orb = cv2.ORB_create(nf, 1.2, 3)
        tt = []
        for _ in range(4):
            t0 = time.perf_counter()
            kps, des = orb.detectAndCompute(img, None)
            tt.append(time.perf_counter() - t0)
        ref = {(k.octave, float(k.pt[0]), float(k.pt[1])): i for i, k in enumerate(kps)}
        mine = [(int(k["octave"]), float(k["x"]), float(k["y"])) for k in kp]
        same_set = set(mine) == set(ref)
        bad = -1
        if same_set:
            bad = int((desc != des[[ref[m] for m in mine]]).any(axis=1).sum())
        out["cpu_baseline"] = {"value": 1e3 * float(np.median(tt[1:])), "unit": "ms/frame", "kind": "reference",
                               "cores": cv2.getNumThreads(), "what": "cv2 %s ORB_create(5000, 1.2, 3).detectAndCompute" % cv2.__version__}
        out["parity"] = {"keypoints_cv2": len(kps), "same_keypoint_set": bool(same_set), "descriptor_rows_differing": bad}
    except ImportError:
        out["cpu_baseline"] = None
    return out


def c5_leg(args, hbm_peak, hbm_src):
    """BASELINE config C5 (RANSAC stress): 100 objects x 2000 correspondences, 90 % outlier matches injected at the
    GuessGenerator boundary, 4096 iterations per object."""
    from tod_b200 import GuessGenerator, synth
    n_obj, n_per, iters = args.c5_objects, 2000, 4096
    g = synth.make_guess_inputs(n_obj, n_per, 0.1, seed=synth.BASE_SEED + 5, k=1, height=960, width=1280)
    gg = GuessGenerator(min_inliers=15, n_ransac_iterations=iters, sensor_error=0.01, seed=11)
    wall, res = [], None
    for _ in range(3):
        t0 = time.perf_counter()
        res = gg.process(g["keypoints_xy"], g["cloud"], g["matches"], g["counts"], g["points3d"], g["spans"],
                         max_poses=32 * n_obj)
        wall.append(time.perf_counter() - t0)
    st = gg.last_stats()
    out = {"workload": "C5: %d objects x %d correspondences, 90%% outliers, %d iterations/object" % (n_obj, n_per, iters),
           "wall_ms": 1e3 * float(np.median(wall[1:])), "poses": int(len(res["pose_results"])),
           "objects_recovered": int(len(set(int(p["object_index"]) for p in res["pose_results"]))),
           "k2": hbm_roofline("k2_adjacency_kernel", st["k2_bytes"], st["k2_ms"], hbm_peak, hbm_src,
                              {"clusters": st["n_clusters"], "correspondences": st["n_correspondences"]}),
           "k3": hbm_roofline("k3_score_kernel", st["k3_bytes"], st["k3_ms"], hbm_peak, hbm_src,
                              {"hypotheses": st["n_hypotheses"], "rounds": st["n_rounds"]}),
           "guess_host_ms": st["host_ms"], "gate_calls": st["gate_calls"], "gate_shape": st["gate_shape"],
           "gate_thread_ms": st["gate_thread_ms"]}
    gg.close()
    return out


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from tod_b200 import DescriptorMatcher, capi, comm_unique_id

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback")

    def trace(msg):
        if args.trace:
            print("[bench rank %d %.1fs] %s" % (rank, time.perf_counter() - t_begin, msg), file=sys.stderr, flush=True)
    t_begin = time.perf_counter()
    if args.trace:
        import faulthandler
        faulthandler.dump_traceback_later(90, exit=False, file=sys.stderr)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)   # plumbing only: id broadcast, barriers, max over ranks
    lib = capi.load()
    trace("process group up")

    descs, points, queries = workload(args)
    kernel = {"auto": capi.TOD_KERNEL_AUTO, "popc": capi.TOD_KERNEL_POPC, "mma": capi.TOD_KERNEL_MMA}[args.kernel]
    m = DescriptorMatcher(k=args.k, radius=args.radius, device=local_rank, shard_rank=rank, shard_count=world,
                          kernel=kernel)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("object_%03d" % i, d, p)
    m.train()
    k = args.k
    nqt = queries.shape[0]
    m.reserve(nqt)
    if world > 1:
        # the exchange lives INSIDE the library: an NCCL communicator owned by the matcher handle (unique id created by
        # rank 0's library, handed over by the host-side plumbing)
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        trace("unique id received")
        m.set_comm(box[0])
        if args.exchange == "nccl":
            m.set_exchange(False)
        trace("communicator up, mode %d" % m.comm_mode)

    stream = torch.cuda.Stream(device=dev)     # explicit, non-legacy: K1 / NCCL / merge are all ordered on it
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    q_dev = torch.from_numpy(queries).to(dev)
    matches = torch.empty((nqt, k, 4), dtype=torch.int32, device=dev)
    counts = torch.empty((nqt,), dtype=torch.int32, device=dev)
    pts3d = torch.empty((nqt, k, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    if world > 1 and args.exchange == "torch":
        keys = torch.empty((nqt, k), dtype=torch.int32, device=dev)
        keys_all = torch.empty((world, nqt, k), dtype=torch.int32, device=dev)

        def step():   # A/B variant: the same kernels, the all-gather issued by torch.distributed between two stage calls
            m.knn_keys_device(q_dev.data_ptr(), nqt, keys.data_ptr(), sptr)
            dist.all_gather_into_tensor(keys_all.view(-1), keys.view(-1))
            m.merge_device(keys_all.data_ptr(), world, nqt, matches.data_ptr(), counts.data_ptr(), pts3d.data_ptr(), sptr)
    else:
        def step():
            # DescriptorMatcher.process on device buffers, one C-ABI call: K1 -> [top-k reduction -> ncclAllGather of
            # the packed keys] -> merge / radius cut / decode / matches_3d gather
            m.process_device(q_dev.data_ptr(), nqt, matches.data_ptr(), counts.data_ptr(), pts3d.data_ptr(), sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    trace("warm-up enqueued")
    barrier()
    trace("warm-up done")

    # ---- device-resident timing (value) ----
    gpu_id = str(local_rank)
    if os.environ.get("CUDA_VISIBLE_DEVICES"):
        vis = os.environ["CUDA_VISIBLE_DEVICES"].split(",")
        if local_rank < len(vis):
            gpu_id = vis[local_rank].strip()
    sampler = ClockSampler(gpu_id)
    launches0 = lib.tod_kernel_launch_count()
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    k1_ms, xch_ms = [], []
    barrier()
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()
        ev[s][0].record(stream)
        step()
        ev[s][1].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    # CUDA events recorded by the library around every K1 launch of the timed region, on the launch stream; read after
    # the loop (reading them inside it would park the host on each step and let the ranks drift apart)
    k1_ms = [x for x in m.k1_ms_history(min(args.steps, 64)) if x >= 0]
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.tod_kernel_launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    frames_total = args.steps * args.frames
    value = frames_total / (dev_ms * 1e-3)
    trace("device-timed region done: %.1f frames/s" % value)

    # ---- end-to-end timing through host (pinned) buffers ----
    q_host = torch.from_numpy(queries).pin_memory()
    m_host = torch.empty((nqt, k, 4), dtype=torch.int32).pin_memory()
    c_host = torch.empty((nqt,), dtype=torch.int32).pin_memory()
    p_host = torch.empty((nqt, k, 3), dtype=torch.float32).pin_memory()
    m_np = m_host.numpy().view(capi.MATCH_DTYPE).reshape(nqt, k)
    out = {"matches": m_np, "counts": c_host.numpy(), "matches_3d": p_host.numpy()}

    if world > 1:
        # sharded: a streaming caller overlaps its copies with the compute — two buffer sets, copy-in / compute /
        # copy-out streams chained by events around tod_matcher_knn_device.  Every step still moves its own inputs from
        # pinned host memory and its own results back.
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        qd = [torch.empty_like(q_dev) for _ in range(2)]
        mt = [torch.empty_like(matches) for _ in range(2)]
        ct = [torch.empty_like(counts) for _ in range(2)]
        pt = [torch.empty_like(pts3d) for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        q_lo = (args.frames * rank // world) * args.keypoints
        q_hi = (args.frames * (rank + 1) // world) * args.keypoints
        own = slice(q_lo, q_hi)
        last = [0]

        def e2e_run(n_steps):
            for i in range(n_steps):
                b = i & 1
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_done[b])              # the compute that last read qd[b] has finished
                    qd[b].copy_(q_host, non_blocking=True)
                    ev_in[b].record(s_in)
                stream.wait_event(ev_in[b])
                stream.wait_event(ev_out[b])                 # the copy-out that last read mt[b] ... has finished
                m.process_device(qd[b].data_ptr(), nqt, mt[b].data_ptr(), ct[b].data_ptr(), pt[b].data_ptr(), sptr)
                ev_done[b].record(stream)
                with torch.cuda.stream(s_out):
                    # every rank holds the complete result in HBM; its host reads back the frames it owns (the guess
                    # generator of a frame runs on one rank) — together the ranks deliver every frame's matches once
                    s_out.wait_event(ev_done[b])
                    m_host[own].copy_(mt[b][own], non_blocking=True)
                    c_host[own].copy_(ct[b][own], non_blocking=True)
                    p_host[own].copy_(pt[b][own], non_blocking=True)
                    ev_out[b].record(s_out)
            torch.cuda.synchronize()
            if n_steps:                                   # outside the timed region's per-step work: nothing
                last[0] = (n_steps - 1) & 1
    else:
        def e2e_run(n_steps):
            for _ in range(n_steps):
                m.process(q_host.numpy(), out=out)      # tod_matcher_knn: H2D + K1 + merge + D2H, synchronous

    e2e_run(3)
    trace("e2e warm-up enqueued")
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = frames_total / e2e_s
    trace("e2e done: %.1f frames/s" % e2e_value)
    h2d = nqt * 32 * world                      # the queries are replicated: every rank uploads all of them
    d2h = nqt * k * 16 + nqt * 4 + nqt * k * 12   # every frame's result is read back once (by the rank that owns it)
    comm_mode = m.comm_mode
    if world > 1:
        # for the checks below: the complete result of the last timed step, read back after the timed region
        b = last[0]
        m_host.copy_(mt[b])
        c_host.copy_(ct[b])
        p_host.copy_(pt[b])
        torch.cuda.synchronize()

    # ---- the sharded reference-facing call itself (tod_matcher_knn, host buffers) must agree with the streamed result
    if world > 1:
        ref_m, ref_c, ref_p = m_np.copy(), c_host.numpy().copy(), p_host.numpy().copy()
        m.process(q_host.numpy(), out=out)
        assert (m_np == ref_m).all() and (c_host.numpy() == ref_c).all() and (p_host.numpy() == ref_p).all()

    if world > 1 and m.exchange_error != 0:
        raise SystemExit("peer exchange timed out on rank %d" % rank)
    shard_rows, kern = m.shard_rows, m.last_kernel
    per_rank = None
    if world > 1:
        # a short profiling loop OUTSIDE the timed regions: the exchange bracketed by events (off in the timed loops)
        m.set_stage_timing(True)
        for _ in range(5):
            m.process_device(q_dev.data_ptr(), nqt, matches.data_ptr(), counts.data_ptr(), pts3d.data_ptr(), sptr)
            xch_ms.append(m.last_exchange_ms)
        m.set_stage_timing(False)
        xch_ms = xch_ms[1:]   # per-rank device times of K1 and of the exchange (the all-gather also waits for the slowest rank)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"rank": rank, "k1_ms": float(np.mean(k1_ms)),
                                          "exchange_ms": float(np.mean(xch_ms)) if xch_ms else None,
                                          "shard_rows": shard_rows})
    barrier()
    m.close()                                            # every rank tears its handle (and communicator) down together
    trace("handle closed")
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- parity of exactly what was timed: bit-exact against the exact-Hamming reference on a fixed sample ----
    parity = parity_sample(args, m_np, c_host.numpy(), p_host.numpy(), descs, points, queries)
    if parity["mismatches"] != 0:
        print(json.dumps({"error": "parity check failed", "parity": parity}), file=sys.stderr)
        raise SystemExit(3)
    assert (c_host.numpy() == k).all() or args.radius > 0

    # ---- roofline of the dominant kernel (K1) ----
    cmp_per_launch = float(nqt) * shard_rows
    k1_avg_ms = float(np.mean(k1_ms))
    achieved_gcmp = cmp_per_launch / (k1_avg_ms * 1e-3) / 1e9
    peaks = load_int_peaks()
    hbm_peak, hbm_src = load_measured_peaks()
    alg_bytes = 32.0 * shard_rows + 32.0 * nqt + nqt * k * 4.0
    if kern == "mma":
        # tensor form: 256 int8 MACs = 512 tensor ops per compare.  peak = the MEASURED dense int8 tcgen05 rate of this
        # pool's B200 (bare MMA loop, tools/mma_peak -> profiles/int_peaks.json; MEASURED_PEAKS.json only holds bf16),
        # falling back to the nominal 4.5 POPS; the nominal figure is reported beside it.
        meas = peaks.get("i8_mma_measured_by_operand_values", {}).get("A +-1, B 0/1 (current encoding)", {}).get(
            "sustained_tops") or peaks.get("i8_mma_measured", {}).get("burst_tops")
        peak_tops = float(meas) if meas else 4500.0
        peak_src = ("measured: bare tcgen05.mma kind::i8 loop with this kernel's operand values, tools/mma_peak "
                    "(profiles/int_peaks.json)") if meas else "nominal 4.5 POPS dense int8 (B200_PROFILING.md table)"
        achieved = achieved_gcmp * 512.0 / 1000.0
        roofline = {"kernel": "k1_mma", "bound": "tensor", "achieved": achieved, "peak": peak_tops, "unit": "TFLOP/s",
                    "unit_note": "int8 tensor operations (TOP/s): 2 per MAC, 512 per 256-bit compare",
                    "frac": achieved / peak_tops, "peak_source": peak_src, "peak_nominal": 4500.0,
                    "frac_of_nominal": achieved / 4500.0, "achieved_gcmp_per_s": achieved_gcmp,
                    "peak_gcmp_per_s": peak_tops * 1000.0 / 512.0}
    else:
        peak = float(peaks["xor_popc_gcmp"])
        roofline = {"kernel": "k1_popc", "bound": "int-popc", "achieved": achieved_gcmp, "peak": peak,
                    "unit": "Gcmp/s", "frac": achieved_gcmp / peak, "peak_source": peaks.get("source", "")}
    roofline.update({"traffic": None, "k1_ms_per_launch": k1_avg_ms, "k1_share_of_step": k1_avg_ms * args.steps / dev_ms,
                     "cmp_per_launch": cmp_per_launch,
                     "hbm_view": {"algorithmic_bytes_per_launch": alg_bytes,
                                  "achieved_gbs": alg_bytes / (k1_avg_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                  "peak_source": hbm_src}})
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        try:
            roofline["traffic"] = json.load(open(tp)).get(kern)
        except Exception:
            pass

    cpu = lsh = None
    if world == 1 and not args.no_cpu_baseline:
        # the exact matcher north_star names, ALL keypoints of a frame per repetition, on the box's host cores
        nsamp = args.keypoints if args.cpu_sample_queries <= 0 else min(args.cpu_sample_queries, args.keypoints)
        times, info = [], None
        t_start = time.perf_counter()
        while len(times) < 8 and (time.perf_counter() - t_start) < 20.0:
            f = len(times) % args.frames
            dt, kind, cores, what, _ = cpu_reference_knn(descs, queries[f * args.keypoints:][:nsamp], args.k)
            times.append(dt)
            info = (kind, cores, what)
        timed = times[1:] if len(times) > 1 else times      # first repetition warms the thread pool / page cache
        cpu_value = (len(timed) * nsamp / float(args.keypoints)) / float(sum(timed))
        cpu = {"value": cpu_value, "unit": UNIT, "cores": info[1], "kind": info[0],
               "sample": "%d x (%d of the %d keypoints of a frame vs the full %d-descriptor DB); %s" % (
                   len(timed), nsamp, args.keypoints, args.objects * args.rows, info[2])}
        if not args.no_lsh:
            k5 = cpu_reference_knn(descs, queries[:args.keypoints], 5)[4]
            lsh = lsh_reference(descs, queries[:args.keypoints], k5)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(args, world),
            "gcmp_per_s": value * args.keypoints * args.objects * args.rows / 1e9,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "scope": ("DescriptorMatcher.process through the C-ABI (tod_matcher_knn) with pinned host buffers"
                              if world == 1 else
                              "tod_matcher_knn_device per step (K1 + in-library exchange + merge), pinned host buffers: "
                              "every rank uploads the whole query batch and reads back the results of the frames it "
                              "owns; copies of step i+1 / i-1 overlapped with the compute of step i")},
            "parity": parity,
            "collective": (None if world == 1 else
                           {"where": ("inside libtod_b200.so: the top-k reduction kernel stores every rank's packed keys "
                                      "straight into the other GPUs' exchange buffers over NVLink (CUDA-IPC-mapped peer "
                                      "memory) and raises a flag there; the merge kernel starts when all flags have "
                                      "arrived — no NCCL kernel on the data path" if comm_mode == 2 else
                                      "inside libtod_b200.so (ncclAllGather of packed top-k keys on the handle's "
                                      "stream)") + "; no torch.distributed collective in the timed region",
                            "comm_mode": comm_mode, "device_timed_exchange": args.exchange, "per_rank": per_rank}),
            "gpu_launches": int(launches), "wall_s_timed_region": wall, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "cpu_baseline_lsh": lsh}
    if world == 1 and not args.no_pipeline:
        torch.cuda.synchronize()
        e2e_pipe, stages = pipeline_leg(args, descs, points, hbm_peak, hbm_src)
        line["e2e_pipeline"] = e2e_pipe
        line["stages"] = stages
        if args.c5_objects > 0:
            line["stages_c5"] = c5_leg(args, hbm_peak, hbm_src)
        line["feature_stage"] = feature_leg(args)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
