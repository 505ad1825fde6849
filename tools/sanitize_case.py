#!/usr/bin/env python
"""Smallest end-to-end case for `compute-sanitizer --tool memcheck` (one tool per gpurun call, B200_PROFILING.md):
K1 (popc or mma, argv[1]), merge/finalize, K2, K3 and the guess generator on tiny inputs, checked against the oracle."""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import geometry as og  # noqa: E402
from oracle import hamming_knn as hk  # noqa: E402
from tod_b200 import DescriptorMatcher, GuessGenerator, capi, fill_adjacency, score_hypotheses, synth  # noqa: E402

kernel = {"popc": capi.TOD_KERNEL_POPC, "mma": capi.TOD_KERNEL_MMA}[sys.argv[1] if len(sys.argv) > 1 else "popc"]
descs, points = synth.make_db(3, [300, 517, 64], seed=1)
q, _, _ = synth.make_queries(descs, 300, seed=2)
m = DescriptorMatcher(k=5, radius=35, kernel=kernel)
for i, (d, p) in enumerate(zip(descs, points)):
    m.add_object("o%d" % i, d, p)
m.train()
out = m.process(q)
em, ec = hk.knn_c(q, descs, 5, 35)
assert (out["counts"] == ec).all() and (out["matches"]["trainIdx"] == em["trainIdx"]).all()
print("K1", m.last_kernel, "ok")
n = 100
qq, tt, px, _, _ = synth.make_cluster(n, 0.6, seed=3)
P, S, _ = fill_adjacency([0, n], qq, tt, px, [0.25], 0.01)
eP, eS = og.fill_adjacency_dense(qq, tt, px, 0.25, 0.01)
assert (P.reshape(n, -1) == og.pack_bits(eP)).all() and (S.reshape(n, -1) == og.pack_bits(eS)).all()
print("K2 ok")
tri = np.array([[a, b, c] for a in range(0, 30, 3) for b in range(1, 30, 7) for c in range(2, 30, 11)], np.uint32)
V = np.packbits(np.concatenate([np.ones(n, bool), np.zeros(og.row_words(n) * 32 - n, bool)]), bitorder="little").view("<u4")
counts, R, T = score_hypotheses(qq, tt, P.reshape(n, -1), V, tri)
exp = np.array([(eP[a] & eP[b] & eP[c]).sum() + 3 for a, b, c in tri])
assert (counts == exp).all()
print("K3 ok")
gi = synth.make_guess_inputs(3, 60, 0.7, seed=4)
gg = GuessGenerator(min_inliers=8, n_ransac_iterations=100, seed=1)
res = gg.process(gi["keypoints_xy"], gi["cloud"], gi["matches"], gi["counts"], gi["points3d"], gi["spans"])
print("guess ok", len(res["pose_results"]), "poses")
m.close()
