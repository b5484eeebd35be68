"""oracle/_ref (the unmodified reference byte-compiled by oracle/build_ref.py) against the C oracle on CPU.

The recipe only runs where /root/reference exists; on a box that only has the prebuilt oracle/_ref the test uses it as it
is, and where neither exists there is nothing to check (the bench then times the C port and says so)."""
import contextlib
import io

import numpy as np
import pytest

import scenes
from oracle import build_ref, oracle


def _load():
    if build_ref.source_available():
        build_ref.build()
    if not build_ref.built():
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref on this machine")
    with contextlib.redirect_stdout(io.StringIO()):
        return build_ref.load()


def test_byte_compiled_reference_dense_equals_c_oracle():
    gf, _ = _load()
    case = scenes.small_case(501, dims=(33, 29, 37), voxel_size=0.02, nframes=3, obs_weights=(1.0, 0.5, 2.0))
    with contextlib.redirect_stdout(io.StringIO()):
        ref = gf.TSDFVolume(case["vol_bnds"].copy(), 0.02, use_gpu=False)
    orc = oracle.OracleTSDFVolume(case["vol_bnds"].copy(), 0.02)
    assert ref.gpu_mode == 0 and gf.__file__.endswith(".refbin")   # the numba CPU path of the sourceless module
    for c, d, p, obs in case["frames"]:
        ref.integrate(c, d, case["K"], p, obs_weight=obs)
        orc.integrate(c, d, case["K"], p, obs_weight=obs)
    t, col = ref.get_volume()
    assert np.array_equal(ref._weight_vol_cpu, orc._weight_vol_cpu)
    assert np.array_equal(t, orc._tsdf_vol_cpu) and np.array_equal(col, orc._color_vol_cpu)


def test_byte_compiled_reference_hash_equals_c_oracle():
    _, hf = _load()
    case = scenes.small_case(502, dims=(18, 16, 20), voxel_size=0.03, nframes=2)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = hf.HashTable(case["vol_bnds"].copy(), 0.03, map_size=4096)
        orh = oracle.OracleHashVox(case["vol_bnds"].copy(), 0.03)
        for c, d, p, _ in case["frames"]:
            ref.integrate(c, d, case["K"], p)
            orh.integrate(c, d, case["K"], p)
        t, col = ref.get_volume()
    assert ref.count_num_hash_entries() == int(orh.entries.sum())
    ot, oc = orh.get_volume()
    assert np.array_equal(t, ot) and np.array_equal(col, oc)
