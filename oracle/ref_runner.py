"""Import the UNMODIFIED reference (/root/reference) as the ground-truth oracle.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does not
exist on the GPU box, so nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` imports this.
It exists to (1) validate oracle/tsdf_oracle.c and (2) generate tests/golden/ fixtures
(tests/golden/make_golden.py).

The reference does ``from skimage import measure`` at import time (grid_fusion.py:6,
hash_fusion.py:4); scikit-image is absent here, so an empty stub module is installed in
sys.modules first.  pycuda is absent too, which selects the numba CPU path
(grid_fusion.py:8-16) -- the path BASELINE.json names as the parity target.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TSDF_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "grid_fusion.py"))


def load():
    """Returns (grid_fusion, hash_fusion) modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        sk.measure = types.ModuleType("skimage.measure")
        sys.modules["skimage"] = sk
        sys.modules["skimage.measure"] = sk.measure
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import grid_fusion  # noqa: E402
    import hash_fusion  # noqa: E402
    return grid_fusion, hash_fusion


def load_bundled_frame(i, data_dir=None):
    """Frame i of the bundled sequence, preprocessed exactly as grid_demo1.py:80-84."""
    import cv2
    import numpy as np
    d = data_dir or os.path.join(REFERENCE_ROOT, "data")
    color = cv2.cvtColor(cv2.imread(os.path.join(d, "frame-%06d.color.jpg" % i)), cv2.COLOR_BGR2RGB)
    depth = cv2.imread(os.path.join(d, "frame-%06d.depth.png" % i), -1).astype(float)
    depth /= 1000.
    depth[depth == 65.535] = 0
    pose = np.loadtxt(os.path.join(d, "frame-%06d.pose.txt" % i))
    return color, depth, pose
