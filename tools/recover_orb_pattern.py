#!/usr/bin/env python
"""Recover ORB's rBRIEF test pattern (OpenCV's bit_pattern_31_, 256 point pairs) from the cv2 binary by black-box
probing — OpenCV's sources are neither vendored in the reference nor present in this image.

cv2.ORB.compute on one keypoint with angle 0 compares, for bit i, the smoothed image at two fixed offsets (x_a, y_a) and
(x_b, y_b).  A vertical step edge at column t, smoothed by ORB's 7 x 7 Gaussian, is strictly increasing on [t-4, t+3]
and flat elsewhere, so bit i fires exactly for t in [x_a - 2, x_b + 3] when x_a < x_b (inverted polarity when
x_a > x_b); horizontal edges give the y coordinates; pairs with equal x (or y) are located with a quadrant image.
Writes oracle/orb_pattern.npy and tod_b200/csrc/orb_pattern.h is generated from it; tests/test_orb_oracle.py checks
that descriptors computed with the table equal cv2.ORB's bit for bit.   usage: python tools/recover_orb_pattern.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = W = 129
CX = CY = 64


def main():
    orb = cv2.ORB_create(nfeatures=10, scaleFactor=1.2, nlevels=1, edgeThreshold=31, patchSize=31)
    kp = [cv2.KeyPoint(float(CX), float(CY), 31.0, 0.0, 1.0, 0, -1)]

    def bits(img):
        return np.unpackbits(orb.compute(img, kp)[1][0], bitorder="little")

    def sweep(axis, inverted):
        first, last = np.full(256, 10 ** 6), np.full(256, -10 ** 6)
        for t in range(CX - 24, CX + 25):
            img = np.zeros((H, W), np.uint8)
            if axis == 0:
                img[:, t:] = 255
            else:
                img[t:, :] = 255
            on = np.nonzero(bits(255 - img if inverted else img))[0]
            first[on] = np.minimum(first[on], t)
            last[on] = np.maximum(last[on], t)
        return first, last

    pat = np.full((256, 4), -99, np.int64)          # x_a, y_a, x_b, y_b
    for axis in (0, 1):
        fn, ln = sweep(axis, False)
        fi, li = sweep(axis, True)
        for i in range(256):
            if fn[i] < 10 ** 6:                      # a < b
                pat[i, axis], pat[i, 2 + axis] = fn[i] + 2 - CX, ln[i] - 3 - CX
            elif fi[i] < 10 ** 6:                    # a > b
                pat[i, 2 + axis], pat[i, axis] = fi[i] + 2 - CX, li[i] - 3 - CX
    for i in range(256):                             # equal coordinates: quadrant images
        xa, ya, xb, yb = pat[i]
        if xa == -99:
            hi, last = max(ya, yb), None
            for t in range(CX - 24, CX + 25):
                img = np.zeros((H, W), np.uint8)
                img[CY + hi:, t:] = 255
                if bits(img if ya < yb else 255 - img)[i]:
                    last = t
            pat[i, 0] = pat[i, 2] = last - 3 - CX
        if ya == -99:
            hi, last = max(pat[i, 0], pat[i, 2]), None
            for t in range(CY - 24, CY + 25):
                img = np.zeros((H, W), np.uint8)
                img[t:, CX + hi:] = 255
                if bits(img if pat[i, 0] < pat[i, 2] else 255 - img)[i]:
                    last = t
            pat[i, 1] = pat[i, 3] = last - 3 - CY
    assert (pat != -99).all() and pat.min() >= -15 and pat.max() <= 15
    out = os.path.join(ROOT, "oracle", "orb_pattern.npy")
    if os.path.exists(out):
        same = (np.load(out) == pat).all()
        print("recovered pattern %s the committed oracle/orb_pattern.npy" % ("EQUALS" if same else "DIFFERS FROM"))
        return 0 if same else 1
    np.save(out, pat.astype(np.int8))
    print("wrote", out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
