#!/usr/bin/env python
"""bench.py — headline benchmark of the TOD detection hot path on B200.

Metric (BASELINE.json): frames/sec at 2k keypoints x 1M-descriptor DB (100 objects x 10k ORB descriptors), exact
Hamming k-NN k=2, on 1/2/4/8 B200; Hamming Gcmp/s vs the matching roofline.

A "step" = one batch of `--frames` synthetic frames (2000 query descriptors each) through the hot path:
  value : frames/s with the queries already resident in HBM (K1 k-NN -> [NCCL all-gather of packed top-k keys when
          the DB is sharded] -> merge/radius/decode/3-D gather), timed with CUDA events, max over ranks;
  e2e   : the same metric through the reference-facing call with HOST buffers (pinned): H2D of the descriptors,
          DescriptorMatcher.process, D2H of matches / counts / matches_3d, wall clock around synchronous calls.
          (The geometry half is measured by tools/bench_geometry.py and, together with the matcher on the C4 stream
          configuration, by tools/bench_pipeline.py.)
N > 1 shards the DB rows over the ranks (strong scaling: the 1M-descriptor DB is fixed).

`--impl reference` times the reference's own CPU implementation of the path on the host cores: OpenCV's
cv::BFMatcher(NORM_HAMMING) (the exact matcher north_star names; through cv2), on a bounded sample of each frame.
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
    ap.add_argument("--cpu-sample-queries", type=int, default=250)
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
    """The reference's matcher on the host: cv2.BFMatcher(NORM_HAMMING) if importable, else the oracle C port."""
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
        return dt, "reference", cv2.getNumThreads(), "cv2 %s BFMatcher(NORM_HAMMING).knnMatch" % cv2.__version__
    except ImportError:
        from oracle import hamming_knn as hk
        t0 = time.perf_counter()
        hk.knn_c(queries, descs, k)
        dt = time.perf_counter() - t0
        return dt, "port", os.cpu_count(), "oracle/hamming_knn.c (OpenMP)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    descs, points, queries = workload(args)
    nsamp = min(args.cpu_sample_queries, args.keypoints)
    times = []
    info = None
    for s in range(args.warmup + args.steps):
        f = s % args.frames
        q = queries[f * args.keypoints: f * args.keypoints + nsamp]
        dt, kind, cores, what = cpu_reference_knn(descs, q, args.k)
        info = (kind, cores, what)
        if s >= args.warmup:
            times.append(dt)
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
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from tod_b200 import DescriptorMatcher, capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = capi.load()

    descs, points, queries = workload(args)
    kernel = {"auto": capi.TOD_KERNEL_AUTO, "popc": capi.TOD_KERNEL_POPC, "mma": capi.TOD_KERNEL_MMA}[args.kernel]
    m = DescriptorMatcher(k=args.k, radius=args.radius, device=local_rank, shard_rank=rank, shard_count=world,
                          kernel=kernel)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("object_%03d" % i, d, p)
    m.train()

    k = args.k
    nqt = queries.shape[0]
    stream = torch.cuda.Stream(device=dev)     # explicit, non-legacy: K1 / NCCL / merge are all ordered on it
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    q_dev = torch.from_numpy(queries).to(dev)
    keys = torch.empty((nqt, k), dtype=torch.int32, device=dev)
    keys_all = torch.empty((world, nqt, k), dtype=torch.int32, device=dev) if world > 1 else keys
    matches = torch.empty((nqt, k, 4), dtype=torch.int32, device=dev)
    counts = torch.empty((nqt,), dtype=torch.int32, device=dev)
    pts3d = torch.empty((nqt, k, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        m.knn_keys_device(q_dev.data_ptr(), nqt, keys.data_ptr(), sptr)
        if world > 1:
            dist.all_gather_into_tensor(keys_all.view(-1), keys.view(-1))
        m.merge_device(keys_all.data_ptr(), world, nqt, matches.data_ptr(), counts.data_ptr(), pts3d.data_ptr(), sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

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
    k1_ms = []
    barrier()
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()
        ev[s][0].record(stream)
        step()
        ev[s][1].record(stream)
        k1_ms.append(m.last_k1_ms)   # CUDA events recorded by the library around the K1 launch, on the launch stream
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.tod_kernel_launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    frames_total = args.steps * args.frames
    value = frames_total / (dev_ms * 1e-3)

    # ---- end-to-end timing through host (pinned) buffers ----
    q_host = torch.from_numpy(queries).pin_memory()
    m_host = torch.empty((nqt, k, 4), dtype=torch.int32).pin_memory()
    c_host = torch.empty((nqt,), dtype=torch.int32).pin_memory()
    p_host = torch.empty((nqt, k, 3), dtype=torch.float32).pin_memory()
    m_np = m_host.numpy().view(capi.MATCH_DTYPE).reshape(nqt, k)
    out = {"matches": m_np, "counts": c_host.numpy(), "matches_3d": p_host.numpy()}

    if world > 1:
        # sharded path: the public API is the device-stage calls, the host<->device copies are the caller's — and a
        # streaming caller overlaps them with compute: two buffer sets, copy-in / compute / copy-out streams chained by
        # events.  Every step still moves its own inputs from pinned host memory and its own results back.
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        qd = [torch.empty_like(q_dev) for _ in range(2)]
        ky = [torch.empty_like(keys) for _ in range(2)]
        ka = [torch.empty_like(keys_all) for _ in range(2)]
        mt = [torch.empty_like(matches) for _ in range(2)]
        ct = [torch.empty_like(counts) for _ in range(2)]
        pt = [torch.empty_like(pts3d) for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]

        def e2e_run(n_steps):
            for i in range(n_steps):
                b = i & 1
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_done[b])              # the compute that last read qd[b] has finished
                    qd[b].copy_(q_host, non_blocking=True)
                    ev_in[b].record(s_in)
                stream.wait_event(ev_in[b])
                stream.wait_event(ev_out[b])                 # the copy-out that last read mt[b] ... has finished
                m.knn_keys_device(qd[b].data_ptr(), nqt, ky[b].data_ptr(), sptr)
                dist.all_gather_into_tensor(ka[b].view(-1), ky[b].view(-1))
                m.merge_device(ka[b].data_ptr(), world, nqt, mt[b].data_ptr(), ct[b].data_ptr(), pt[b].data_ptr(), sptr)
                ev_done[b].record(stream)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[b])
                    m_host.copy_(mt[b], non_blocking=True)
                    c_host.copy_(ct[b], non_blocking=True)
                    p_host.copy_(pt[b], non_blocking=True)
                    ev_out[b].record(s_out)
            torch.cuda.synchronize()
    else:
        def e2e_run(n_steps):
            for _ in range(n_steps):
                m.process(q_host.numpy(), out=out)      # tod_matcher_knn: H2D + K1 + merge + D2H, synchronous

    e2e_run(3)
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
    h2d = nqt * 32
    d2h = nqt * k * 16 + nqt * 4 + nqt * k * 12

    # ---- sanity: the timed path produced real matches (planted queries are found) ----
    res = m_np
    if not os.environ.get("TOD_K1_DEBUG_MODE"):
        assert (c_host.numpy() == k).all() or args.radius > 0
        assert (res["distance"][:, 0] <= res["distance"][:, -1]).all()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (K1) ----
    shard_rows = m.shard_rows
    cmp_per_launch = float(nqt) * shard_rows
    k1_avg_ms = float(np.mean(k1_ms))
    achieved_gcmp = cmp_per_launch / (k1_avg_ms * 1e-3) / 1e9
    peaks = load_int_peaks()
    hbm_peak, hbm_src = load_measured_peaks()
    alg_bytes = 32.0 * shard_rows + 32.0 * nqt + nqt * k * 4.0
    kern = m.last_kernel
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

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        nsamp = min(args.cpu_sample_queries, args.keypoints)
        times, info = [], None
        t_start = time.perf_counter()
        while len(times) < 7 and (time.perf_counter() - t_start) < 15.0:
            f = len(times) % args.frames
            dt, kind, cores, what = cpu_reference_knn(descs, queries[f * args.keypoints:][:nsamp], args.k)
            times.append(dt)
            info = (kind, cores, what)
        timed = times[1:] if len(times) > 1 else times      # first repetition warms the thread pool / page cache
        cpu_value = (len(timed) * nsamp / float(args.keypoints)) / float(sum(timed))
        cpu = {"value": cpu_value, "unit": UNIT, "cores": info[1], "kind": info[0],
               "sample": "%d x (%d of the %d keypoints of a frame vs the full %d-descriptor DB); %s" % (
                   len(timed), nsamp, args.keypoints, args.objects * args.rows, info[2])}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(args, world),
            "gcmp_per_s": value * args.keypoints * args.objects * args.rows / 1e9,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "scope": ("DescriptorMatcher.process through the C-ABI with pinned host buffers" if world == 1 else
                              "device-stage C-ABI calls + NCCL all-gather, pinned host buffers, copies of step i+1 / "
                              "i-1 overlapped with the compute of step i (two buffer sets)")},
            "gpu_launches": int(launches), "wall_s_timed_region": wall, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
