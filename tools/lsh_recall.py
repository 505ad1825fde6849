#!/usr/bin/env python
"""north_star: "the reference LSH matcher is reported separately as recall".  The reference's DescriptorMatcher builds
cv::FlannBasedMatcher(LshIndexParams(n_tables, key_size, multi_probe_level)) (src/detection/DescriptorMatcher.cpp:175-181,
parameters from conf/detection.ork:32-39: 10 tables, key size 16, multi-probe 1), runs knnMatch(k = 5) and cuts at
radius 35 (:211-220).  This tool runs exactly that through cv2 on the host (CPU only) next to the exact search
(cv2.BFMatcher(NORM_HAMMING), which is what libtod_b200 reproduces bit for bit) and reports, on the BASELINE C1 / C2
shapes: recall of the exact within-radius matches, fraction of queries whose best match agrees, and the CPU times.
usage: python tools/lsh_recall.py [out.json]"""
import json
import os
import sys
import time

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tod_b200 import synth  # noqa: E402


def run(name, n_obj, rows, nq, radius=35, k=5, flip_p=0.04):
    descs, _ = synth.make_db(n_obj, rows, seed=synth.BASE_SEED + 1)
    q, src_obj, src_row = synth.make_queries(descs, nq, seed=synth.BASE_SEED + 101, flip_p=flip_p)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    bf.add([np.ascontiguousarray(d) for d in descs])
    t0 = time.perf_counter()
    exact = bf.knnMatch(q, k)
    t_bf = time.perf_counter() - t0
    lsh = cv2.FlannBasedMatcher(dict(algorithm=6, table_number=10, key_size=16, multi_probe_level=1), dict())
    lsh.add([np.ascontiguousarray(d) for d in descs])
    t0 = time.perf_counter()
    lsh.train()
    t_train = time.perf_counter() - t0
    t0 = time.perf_counter()
    approx = lsh.knnMatch(q, k)
    t_lsh = time.perf_counter() - t0

    def cut(lst):  # DescriptorMatcher.cpp:212-220
        out = []
        for m in lst[:5]:
            if m.distance > radius:
                break
            out.append((m.imgIdx, m.trainIdx))
        return out
    tot = found = best_same = with_match = 0
    for e, a in zip(exact, approx):
        es, as_ = cut(e), cut(a)
        tot += len(es)
        found += len(set(es) & set(as_))
        if es:
            with_match += 1
            best_same += bool(as_) and as_[0] == es[0]
    return {"case": name, "bit_flip_probability": flip_p, "db_descriptors": n_obj * rows, "queries": nq, "k": k, "radius": radius,
            "exact_matches_within_radius": tot, "lsh_recall_of_exact_matches": found / max(tot, 1),
            "queries_with_an_exact_match": with_match, "lsh_best_match_agrees": best_same / max(with_match, 1),
            "cpu_ms": {"BFMatcher_exact_%d_threads" % cv2.getNumThreads(): 1e3 * t_bf, "LSH_train": 1e3 * t_train,
                       "LSH_knnMatch": 1e3 * t_lsh}}


def main():
    res = {"lsh_params": {"n_tables": 10, "key_size": 16, "multi_probe_level": 1, "source": "conf/detection.ork:32-39"},
           "opencv": cv2.__version__,
           "cases": [run("C1: 1 object x 5000 descriptors, 5000 keypoints", 1, 5000, 5000),
                     run("C2: 10 objects x 5000 descriptors, 1000 keypoints", 10, 5000, 1000),
                     run("C2 with noisier descriptors (10% of the bits flipped, mean distance 26)", 10, 5000, 1000,
                         flip_p=0.10),
                     run("C2 with 12% of the bits flipped (mean distance 31, half of the matches beyond the radius)",
                         10, 5000, 1000, flip_p=0.12)]}
    s = json.dumps(res, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(s)
    print(s)


if __name__ == "__main__":
    main()
