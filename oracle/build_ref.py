"""Recipe for oracle/_ref/: the UNMODIFIED reference, byte-compiled where it lies.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference (DiWu9/Union-Thesis-SLAM) is pure Python +
numba, so "compiling it from its own sources" means ``py_compile``: every module of the hot path
is compiled from /root/reference into sourceless bytecode (``<module>.refbin``: the ``.pyc`` image under a
suffix that file-sync tools do not drop as a cache file) under ``oracle/_ref/`` -- no reference
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
import marshal
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


SUFFIX = ".refbin"


def _out(rel):
    return os.path.join(REF_OUT, rel[:-3] + SUFFIX)


def built():
    return all(os.path.isfile(_out(m)) for m in MODULES)


def build(force=False):
    """Byte-compiles the reference's hot-path modules into oracle/_ref/.  Returns the directory,
    or None when /root/reference is absent (GPU box: the prebuilt files are used as they are)."""
    if not source_available():
        return REF_OUT if built() else None
    for rel in MODULES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = _out(rel)
        if not force and os.path.isfile(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path shown in tracebacks
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
    magic = importlib.util.MAGIC_NUMBER

    def load_module(name, rel, package=None):
        if name in sys.modules and getattr(sys.modules[name], "__file__", "").startswith(REF_OUT):
            return sys.modules[name]
        path = _out(rel)
        with open(path, "rb") as fh:
            data = fh.read()
        if data[:4] != magic:
            raise RuntimeError("%s was byte-compiled by another Python (magic %s, this interpreter %s): rebuild "
                               "oracle/_ref" % (path, data[:4].hex(), magic.hex()))
        mod = types.ModuleType(name)
        mod.__file__ = path
        mod.__package__ = package or ""
        sys.modules[name] = mod
        exec(marshal.loads(data[16:]), mod.__dict__)  # the module body, exactly what `import` would run
        return mod

    pkg = sys.modules.get("data_structures")
    if pkg is None or not getattr(pkg, "__file__", "").startswith(REF_OUT):
        pkg = types.ModuleType("data_structures")
        pkg.__file__ = os.path.join(REF_OUT, "data_structures")
        pkg.__path__ = []
        sys.modules["data_structures"] = pkg
    for sub in ("voxel", "hash_entry", "bucket"):  # dependency order of the reference's own imports
        setattr(pkg, sub, load_module("data_structures." + sub, "data_structures/%s.py" % sub, "data_structures"))
    grid_fusion = load_module("grid_fusion", "grid_fusion.py")
    hash_fusion = load_module("hash_fusion", "hash_fusion.py")
    return grid_fusion, hash_fusion


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else "reference tree not found at %s and oracle/_ref not prebuilt" % REFERENCE_ROOT)
