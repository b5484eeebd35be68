"""Recipe for oracle/_ref/: the UNMODIFIED reference, byte-compiled where it lies.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference (DiWu9/Union-Thesis-SLAM) is pure Python +
numba, so "compiling it from its own sources" means ``py_compile``: every module of the hot path
is compiled from /root/reference into a sourceless ``.pyc`` under ``oracle/_ref/`` -- no reference
source text enters this repository, and ``oracle/_ref/`` is git-ignored (it is NOT gpurun-ignored,
so it travels to the GPU box like our own built ``.so`` files).  numba JIT-compiles from bytecode,
so the sourceless modules run the reference's numba CPU path unchanged (grid_fusion.py:170-212,
214-314; hash_fusion.py:103-145).

Used by: ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg (timing the
reference's own ``TSDFVolume.integrate`` / ``HashTable.integrate`` on the GPU box's host cores),
and by tests that cross-check the C oracle against the reference.  Never by the product package.

    python -m oracle.build_ref        # needs /root/reference; no-op with a note when it is absent
"""
import importlib.util
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("TSDF_REFERENCE_ROOT", "/root/reference")
MODULES = ["grid_fusion.py", "hash_fusion.py", "data_structures/bucket.py", "data_structures/hash_entry.py",
           "data_structures/voxel.py"]


def source_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "grid_fusion.py"))


def built():
    return all(os.path.isfile(os.path.join(REF_OUT, m + "c")) for m in MODULES)


def build(force=False):
    """Byte-compiles the reference's hot-path modules into oracle/_ref/.  Returns the directory,
    or None when /root/reference is absent (GPU box: the prebuilt files are used as they are)."""
    if not source_available():
        return REF_OUT if built() else None
    for rel in MODULES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = os.path.join(REF_OUT, rel + "c")
        if not force and os.path.isfile(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path shown in tracebacks; sourceless import needs the .pyc beside where the .py would be
        py_compile.compile(src, cfile=dst, dfile="reference/" + rel, doraise=True, optimize=0)
    with open(os.path.join(REF_OUT, "BUILD_INFO.txt"), "w") as fh:
        fh.write("byte-compiled from %s by oracle/build_ref.py with Python %s (magic %s)\n" % (
            REFERENCE_ROOT, sys.version.split()[0], importlib.util.MAGIC_NUMBER.hex()))
    return REF_OUT


def load():
    """Imports (grid_fusion, hash_fusion) of the reference from oracle/_ref/.  scikit-image and pycuda
    are absent from the image: an empty ``skimage.measure`` stub satisfies the import-time dependency
    (grid_fusion.py:6, hash_fusion.py:4) and the missing pycuda selects the numba CPU path
    (grid_fusion.py:8-16) -- the path BASELINE.json names as the parity and timing target."""
    if not built():
        raise RuntimeError("oracle/_ref is not built (run `python -m oracle.build_ref` where /root/reference exists)")
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        sk.measure = types.ModuleType("skimage.measure")
        sys.modules["skimage"] = sk
        sys.modules["skimage.measure"] = sk.measure
    if REF_OUT not in sys.path:
        sys.path.insert(0, REF_OUT)
    import grid_fusion  # noqa: E402  (the reference's module name)
    import hash_fusion  # noqa: E402
    if not grid_fusion.__file__.startswith(REF_OUT):
        raise RuntimeError("grid_fusion resolved to %s, not to oracle/_ref" % grid_fusion.__file__)
    return grid_fusion, hash_fusion


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else "reference tree not found at %s and oracle/_ref not prebuilt" % REFERENCE_ROOT)
