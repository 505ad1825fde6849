#!/usr/bin/env python
"""Generate tests/golden/geom_*.npz from the REFERENCE'S OWN geometry code: src/common/adjacency_ransac.cpp,
maximum_clique.cpp and the sac*.h / ransac.h headers of wg-perception/tod 0.5.6, compiled unmodified from
/root/reference into oracle/_ref/libtod_ref.so (oracle/build_ref.py).  Run once in the build container:

    python tests/golden/make_geometry_golden.py

The .npz files are committed; tests need neither /root/reference nor oracle/_ref at run time.

  geom_adjacency_*.npz : one cluster — query/train points, pixels, span, sensor_error  ->  the neighbour lists of
                         FillAdjacency (adjacency_ransac.cpp:127-172) as dense n x n bool matrices (physical, sample)
  geom_guess_*.npz     : GuessGenerator inputs injected at the cell boundary (keypoints, cloud, matches, counts,
                         matches_3d, spans; parameters; sampler seed)  ->  the poses and inlier keypoint sets of the
                         reference's ClusterPerObject + FillAdjacency + Ransac loop (GuessGenerator.cpp:127-250),
                         with the harness's rand() driven by the shared seeded stream (tod_rng_*)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
from tod_b200 import synth  # noqa: E402


def adjacency_case(name, n, frac, seed, span, err, nan_rows=()):
    q, t, px, _, _ = synth.make_cluster(n, frac, seed=seed, span=span)
    for r, c in nan_rows:
        q[r, c] = np.nan
    ar = ref.RefAdjacencyRansac()
    for i in range(n):
        ar.add_points(t[i], q[i], i)
    ar.fill_adjacency(px, span, err)
    np.savez_compressed(os.path.join(HERE, name), query=q, train=t, pixels=px, span=np.float32(span),
                        sensor_error=np.float32(err), physical=np.packbits(ar.dense("physical"), axis=1),
                        sample=np.packbits(ar.dense("sample"), axis=1), n=n)
    print(name, n, int(ar.dense("physical").sum()), int(ar.dense("sample").sum()))


def guess_case(name, n_obj, n_per, frac, iters, min_inliers, seed, rng_seed):
    gi = synth.make_guess_inputs(n_obj, n_per, frac, seed=seed)
    exp = ref.process(gi["keypoints_xy"], gi["cloud"], gi["matches"], gi["counts"], gi["points3d"], gi["spans"],
                      min_inliers, iters, 0.01, seed=rng_seed)
    # the cloud is sparse: store only the finite pixels
    ys, xs = np.nonzero(~np.isnan(gi["cloud"][:, :, 0]))
    inl = np.concatenate([np.array(e[3], np.int32) for e in exp]) if exp else np.zeros(0, np.int32)
    np.savez_compressed(
        os.path.join(HERE, name), keypoints_xy=gi["keypoints_xy"], cloud_shape=np.array(gi["cloud"].shape[:2]),
        cloud_y=ys.astype(np.int32), cloud_x=xs.astype(np.int32), cloud_v=gi["cloud"][ys, xs],
        matches=gi["matches"], counts=gi["counts"], points3d=gi["points3d"], spans=gi["spans"],
        min_inliers=min_inliers, n_ransac_iterations=iters, sensor_error=np.float32(0.01), seed=rng_seed,
        pose_object=np.array([e[0] for e in exp], np.int32), pose_R=np.array([e[1] for e in exp], np.float32),
        pose_T=np.array([e[2] for e in exp], np.float32), pose_n_inliers=np.array([len(e[3]) for e in exp], np.int32),
        inliers=inl)
    print(name, "poses:", len(exp), "inliers:", [len(e[3]) for e in exp])


def main():
    assert ref.available(), "oracle/_ref/libtod_ref.so must be buildable (needs /root/reference)"
    adjacency_case("geom_adjacency_n257_half.npz", 257, 0.5, 11, 0.25, 0.01)
    adjacency_case("geom_adjacency_n700_outliers.npz", 700, 0.2, 12, 0.2, 0.005)
    adjacency_case("geom_adjacency_n64_nan.npz", 64, 0.8, 13, 0.25, 0.01, nan_rows=((5, 1), (40, 2)))
    guess_case("geom_guess_6x150.npz", 6, 150, 0.5, 300, 8, 1150, 11)
    guess_case("geom_guess_4x400_outliers.npz", 4, 400, 0.15, 800, 8, 1400, 23)


if __name__ == "__main__":
    main()
