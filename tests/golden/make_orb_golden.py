#!/usr/bin/env python
"""Golden vectors for the feature stage, generated with the cv2 of this image:
cv2.ORB_create(5000, 1.2, 3) — the configuration of conf/detection.ork:23-31 — on deterministic synthetic frames
(tod_b200.synth.make_textured_image).  Stored per frame: the keypoints cv2 selected (x, y, octave, angle), and their
descriptors.  The frames themselves are regenerated from their seeds.   usage: python tests/golden/make_orb_golden.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tod_b200 import synth  # noqa: E402

CASES = [("orb_640x480", 480, 640, 11, 1500), ("orb_333x517_ragged", 333, 517, 12, 900)]


def main():
    for name, h, w, seed, keep in CASES:
        img = synth.make_textured_image(h, w, seed=seed)
        orb = cv2.ORB_create(5000, 1.2, 3)
        kp, des = orb.detectAndCompute(img, None)
        idx = np.arange(len(kp))                                               # every keypoint cv2 selected
        out = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(out, height=h, width=w, seed=seed, cv2_version=cv2.__version__,
                            x=np.array([kp[i].pt[0] for i in idx], np.float32),
                            y=np.array([kp[i].pt[1] for i in idx], np.float32),
                            octave=np.array([kp[i].octave for i in idx], np.int32),
                            angle=np.array([kp[i].angle for i in idx], np.float32),
                            response=np.array([kp[i].response for i in idx], np.float32),
                            size=np.array([kp[i].size for i in idx], np.float32),
                            descriptors=des[idx], n_detected=len(kp))
        print(out, len(idx), "of", len(kp), "keypoints; octaves", np.bincount([kp[i].octave for i in idx]))


if __name__ == "__main__":
    main()
