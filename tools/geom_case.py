#!/usr/bin/env python
"""Small, fixed geometry case for ncu: C5-shaped input (20 objects x 2000 correspondences, 90 % outliers, 4096
iterations) through GuessGenerator.process — launches K2 (k2_adjacency_kernel), the per-round degree-mask kernel, K3
(k3_score_kernel) and K4 (k4_gate_kernel) — plus one dense C4-shaped frame (4 objects x 700 inliers).
usage: python tools/geom_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tod_b200 import GuessGenerator, synth  # noqa: E402


def main():
    g = synth.make_guess_inputs(20, 2000, 0.1, seed=synth.BASE_SEED + 5, k=1, height=960, width=1280)
    gg = GuessGenerator(min_inliers=15, n_ransac_iterations=4096, sensor_error=0.01, seed=11)
    res = gg.process(g["keypoints_xy"], g["cloud"], g["matches"], g["counts"], g["points3d"], g["spans"],
                     max_poses=640)
    st = gg.last_stats()
    print("C5-shaped: poses", len(res["pose_results"]), "k2_ms", st["k2_ms"], "k3_ms", st["k3_ms"], "hyp",
          st["n_hypotheses"], "gate", st["gate_shape"])
    d = synth.make_guess_inputs(4, 700, 0.97, seed=5, k=1, height=960, width=1280)
    res = gg.process(d["keypoints_xy"], d["cloud"], d["matches"], d["counts"], d["points3d"], d["spans"])
    print("C4-shaped: poses", len(res["pose_results"]), gg.last_stats()["host_ms"])
    gg.close()


if __name__ == "__main__":
    main()
