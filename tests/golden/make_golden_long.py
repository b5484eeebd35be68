"""Long-history golden: the UNMODIFIED reference over the first 100 bundled frames (config 1, 2 cm).

    python tests/golden/make_golden_long.py          # ~7 min on 8 cores; needs /root/reference

Writes frames 10..99 next to the short fixture's frames (tests/golden/frames) and golden_long.json:
decoded-input digests plus sha256 / counts of the reference TSDFVolume after 50 and 100 frames -- the rows of
SURVEY.md Appendix B (2e9bb509beb8d38d / 72591f9faf816955 / db891032ce9791f5 and 30144e7bf800e1cc /
e9d842e8c8beb495 / fe43593d4f315064).  Refuses to write anything the C oracle disagrees with.
"""
import hashlib
import json
import os
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle, ref_runner  # noqa: E402
import scenes  # noqa: E402

N_FRAMES = 100
CHECKPOINTS = (50, 100)
APPENDIX_B = {50: ("2e9bb509beb8d38d", "72591f9faf816955", "db891032ce9791f5", 467564, 17060168),
              100: ("30144e7bf800e1cc", "e9d842e8c8beb495", "fe43593d4f315064", 567594, 29899191)}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    grid_fusion, _ = ref_runner.load()
    src = os.path.join(ref_runner.REFERENCE_ROOT, "data")
    fdir = os.path.join(HERE, "frames")
    K = np.loadtxt(os.path.join(src, "camera-intrinsics.txt"), delimiter=" ")
    gold = {"generated_with": "reference grid_fusion.py TSDFVolume, numba CPU path, frames 0..%d in order" % (N_FRAMES - 1),
            "inputs": {}, "dense": {}}
    ref = grid_fusion.TSDFVolume(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    orc = oracle.OracleTSDFVolume(np.array(scenes.CONFIG1_BOUNDS), scenes.CONFIG1_VOXEL)
    t0 = time.time()
    for n, i in enumerate(range(N_FRAMES), 1):
        for ext in ("color.jpg", "depth.png", "pose.txt"):
            name = "frame-%06d.%s" % (i, ext)
            if not os.path.exists(os.path.join(fdir, name)):
                shutil.copyfile(os.path.join(src, name), os.path.join(fdir, name))
        c, d, p = ref_runner.load_bundled_frame(i)
        gold["inputs"][str(i)] = {"color": sha(c), "depth": sha(d), "pose": sha(p)}
        ref.integrate(c, d, K, p, obs_weight=1.)
        orc.integrate(c, d, K, p, obs_weight=1.)
        if n in CHECKPOINTS:
            rec = {"tsdf": sha(ref._tsdf_vol_cpu), "weight": sha(ref._weight_vol_cpu), "color": sha(ref._color_vol_cpu),
                   "n_w_pos": int((ref._weight_vol_cpu > 0).sum()),
                   "sum_w": int(ref._weight_vol_cpu.astype(np.float64).sum()),
                   "max_w": float(ref._weight_vol_cpu.max())}
            assert (rec["tsdf"], rec["weight"], rec["color"]) == (sha(orc._tsdf_vol_cpu), sha(orc._weight_vol_cpu),
                                                                   sha(orc._color_vol_cpu)), "oracle != reference at %d" % n
            b = APPENDIX_B[n]
            assert (rec["tsdf"][:16], rec["weight"][:16], rec["color"][:16], rec["n_w_pos"], rec["sum_w"]) == b, \
                "reference run differs from SURVEY Appendix B at %d frames: %r" % (n, rec)
            gold["dense"][str(n)] = rec
            print("after", n, rec, "%.0fs" % (time.time() - t0), flush=True)
    with open(os.path.join(HERE, "golden_long.json"), "w") as fh:
        json.dump(gold, fh, indent=1)
    print("wrote golden_long.json")


if __name__ == "__main__":
    main()
