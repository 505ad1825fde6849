#!/usr/bin/env python
"""Where the K1 tensor-core kernel spends its cycles: runs the C3 workload through an INSTRUMENTED build of the library
(TOD_B200_VARIANT=stats TOD_B200_DEFINES=-DTOD_K1_STATS=1 python -m tod_b200._build; TOD_B200_LIB=<that .so>) and prints
the counters of tod_debug_k1_stats.  Diagnostic tool, not part of the product path."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tod_b200 import DescriptorMatcher, capi, synth  # noqa: E402


def main():
    frames_list = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,8,64").split(",")]
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    shards = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # time rank 0 of `shards` (the per-GPU work at N GPUs)
    lib = capi.load()
    descs, points = synth.make_db(100, 10000, seed=synth.BASE_SEED + 2)
    m = DescriptorMatcher(k=k, radius=0, kernel=capi.TOD_KERNEL_MMA, shard_rank=0, shard_count=shards)
    for i, (d, p) in enumerate(zip(descs, points)):
        m.add_object("o%d" % i, d, p)
    m.train()
    dev = torch.device("cuda", 0)
    names = ["slow_calls", "slow_cycles", "mma_wait_acc_empty", "mma_wait_b_full", "epi_wait_acc_full", "cta_cycles",
             "epi_groups", "epi_busy_cycles", "ctas", "tma_wait_b_empty", "slow_ld_cycles", "slow_mask_cycles",
             "slow_loop_cycles", "slow_events", "epi_hold_cycles"]
    for F in frames_list:
        q = np.concatenate([synth.make_queries(descs, 2000, seed=synth.BASE_SEED + 102 + f)[0] for f in range(F)])
        q_dev = torch.from_numpy(q).to(dev)
        keys = torch.empty((q.shape[0], k), dtype=torch.int32, device=dev)
        out = (ctypes.c_ulonglong * 16)()
        tile = (ctypes.c_ulonglong * 256)()
        for it in range(3):
            m.knn_keys_device(q_dev.data_ptr(), q.shape[0], keys.data_ptr(), None)
            torch.cuda.synchronize()
            if it == 1:
                lib.tod_debug_k1_stats(out, 1)      # drop the warm-up counts
                lib.tod_debug_k1_tile_stats(tile, 1)
        ms = m.last_k1_ms
        lib.tod_debug_k1_stats(out, 1)
        lib.tod_debug_k1_tile_stats(tile, 1)
        s = dict(zip(names, [int(x) for x in out]))
        ctas = max(s["ctas"], 1)
        per_cta = s["cta_cycles"] / ctas
        d = {"frames": F, "k": k, "shards": shards, "k1_ms": ms,
             "gcmp_s": q.shape[0] * float(m.shard_rows) / (ms * 1e-3) / 1e9, "ctas": ctas, "cycles_per_cta": per_cta,
             "slow_call_rate": s["slow_calls"] / max(s["epi_groups"], 1),
             "cycles_per_slow_call": s["slow_cycles"] / max(s["slow_calls"], 1),
             "mma_wait_acc_empty_frac": s["mma_wait_acc_empty"] / max(s["cta_cycles"], 1),
             "mma_wait_b_full_frac": s["mma_wait_b_full"] / max(s["cta_cycles"], 1),
             "tma_wait_b_empty_frac": s["tma_wait_b_empty"] / max(s["cta_cycles"], 1),
             "epi_wait_acc_full_frac_per_warp": s["epi_wait_acc_full"] / 8.0 / max(s["cta_cycles"], 1),
             "epi_busy_frac_per_warp": s["epi_busy_cycles"] / 8.0 / max(s["cta_cycles"], 1),
             "epi_slow_frac_per_warp": s["slow_cycles"] / 8.0 / max(s["cta_cycles"], 1),
             "slow_ld_cyc": s["slow_ld_cycles"] / max(s["slow_calls"], 1),
             "slow_mask_cyc": s["slow_mask_cycles"] / max(s["slow_calls"], 1),
             "slow_loop_cyc": s["slow_loop_cycles"] / max(s["slow_calls"], 1),
             "slow_events_per_call": s["slow_events"] / max(s["slow_calls"], 1),
             "epi_hold_cycles_per_tile": s["epi_hold_cycles"] * 8.0 / max(s["epi_groups"], 1),
             "epi_busy_cycles_per_group": s["epi_busy_cycles"] / max(s["epi_groups"], 1), "raw": s}
        # per tile index of a CTA's sweep (last bucket = tiles >= 63), per CTA: where the start-up goes
        tl = [int(x) for x in tile]
        d["by_tile"] = {"mma_wait_acc_empty_cycles_per_cta": [round(x / ctas * 2, 0) for x in tl[0:64]],
                        "slow_calls_per_cta": [round(x / ctas, 2) for x in tl[64:128]],
                        "slow_cycles_per_call": [round(tl[128 + i] / max(tl[64 + i], 1), 0) for i in range(64)],
                        "mma_wait_b_full_cycles_per_cta": [round(x / ctas * 2, 0) for x in tl[192:256]]}
        print(json.dumps(d))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
