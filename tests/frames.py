"""Bundled-sequence fixture frames (tests/golden/frames), decoded exactly as grid_demo1.py:80-84."""
import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRAMES_DIR = os.path.join(GOLDEN_DIR, "frames")


def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as fh:
        return json.load(fh)


def golden_long():
    """Reference state after 50 and 100 bundled frames (tests/golden/make_golden_long.py)."""
    with open(os.path.join(GOLDEN_DIR, "golden_long.json")) as fh:
        return json.load(fh)


def intrinsics():
    return np.loadtxt(os.path.join(FRAMES_DIR, "camera-intrinsics.txt"), delimiter=" ")


def load_frame(i, check=True):
    import cv2
    color = cv2.cvtColor(cv2.imread(os.path.join(FRAMES_DIR, "frame-%06d.color.jpg" % i)), cv2.COLOR_BGR2RGB)
    depth = cv2.imread(os.path.join(FRAMES_DIR, "frame-%06d.depth.png" % i), -1).astype(float)
    depth /= 1000.
    depth[depth == 65.535] = 0
    pose = np.loadtxt(os.path.join(FRAMES_DIR, "frame-%06d.pose.txt" % i))
    if check:
        g = golden()["inputs"].get(str(i)) or golden_long()["inputs"][str(i)]
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        # a different JPEG/PNG decoder would silently invalidate every digest downstream
        assert sha(color) == g["color"] and sha(depth) == g["depth"] and sha(pose) == g["pose"], \
            "decoded fixture frame %d differs from the one the goldens were made with" % i
    return color, depth, pose


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
