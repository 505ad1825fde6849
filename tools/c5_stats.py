import sys, json, time
sys.path.insert(0, '.')
import numpy as np
from tod_b200 import GuessGenerator, synth
g = synth.make_guess_inputs(100, 2000, 0.1, seed=synth.BASE_SEED + 5, k=1, height=960, width=1280)
gg = GuessGenerator(min_inliers=15, n_ransac_iterations=4096, sensor_error=0.01, seed=11)
for _ in range(2):
    t0=time.perf_counter()
    res = gg.process(g["keypoints_xy"], g["cloud"], g["matches"], g["counts"], g["points3d"], g["spans"], max_poses=3200)
    print(time.perf_counter()-t0)
print(json.dumps(gg.last_stats()))
