"""CPU: the C-ABI shared library loads and exports every symbol include/highway_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "highway_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gym_highway_b200 import build, _lib
    build.build()
    names = _declared()
    assert "hw_sa_step" in names and "hw_sa_reset" in names and len(names) >= 6
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_lib.exported_symbols()) == names
    assert _lib.load().hw_abi_version() == _lib.ABI_VERSION


def test_bad_arguments_are_rejected_without_a_gpu():
    from gym_highway_b200 import _lib
    lib = _lib.load()
    st = _lib.HwSaState(0, 0, 0, 0)
    assert lib.hw_sa_reset(ctypes.byref(st), None, None, 16, None) == -1
    st = _lib.HwSaState(16, 16, 16, 16)
    assert lib.hw_sa_reset(ctypes.byref(st), None, None, 0, None) == -1
    st = _lib.HwSaState(8, 16, 16, 16)
    assert lib.hw_sa_reset(ctypes.byref(st), None, None, 4, None) == -2
    assert b"aligned" in lib.hw_error_string(-2)


def test_consumer_entry_points_reject_bad_arguments_without_a_gpu():
    from gym_highway_b200 import _lib
    lib = _lib.load()
    ring = _lib.HwRing(0, 0, 0, 0, 0, 0)
    blk = _lib.HwTransitionBlock(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    assert lib.hw_ring_store(ctypes.byref(ring), ctypes.byref(blk), None, 8, 5, 34, 2, 100, 1.0, None) == -1   # null ring arrays
    assert lib.hw_ring_store(None, ctypes.byref(blk), None, 8, 5, 34, 2, 100, 1.0, None) == -1
    ring = _lib.HwRing(16, 16, 16, 16, 16, 16)
    assert lib.hw_ring_store(ctypes.byref(ring), ctypes.byref(blk), None, 8, 5, 34, 2, 100, 1.0, None) == -1   # null block arrays
    assert lib.hw_ring_store(ctypes.byref(ring), ctypes.byref(blk), None, 0, 5, 34, 2, 100, 1.0, None) == -1   # no transitions
    assert lib.hw_obs_moments(None, 8, 170, 8.0, None, None, None) == -1
    # workspace of the two-stage column sums: one {sum, sum of squares} row of doubles per 64-row slab
    assert lib.hw_obs_moments_workspace(1000, 170) == 16 * 2 * 170 * 8 and lib.hw_obs_moments_workspace(0, 170) == 0
    assert lib.hw_gae(None, None, None, None, 0.99, 0.95, None, None, 5, 8, None) == -1
    assert lib.hw_nstep_returns(None, None, None, 0.99, None, 5, 8, None) == -1
    st = _lib.HwMaState(16, 16, 16, 9, 16)                                     # more cars than the extension's eight
    assert lib.hw_ma_reset(ctypes.byref(st), None, None, 4, None) == -1


def test_sim_params_follow_the_reference_clock():
    from gym_highway_b200 import _lib
    p = _lib.sim_params(False)
    assert (p.dt, p.ticks_per_step, p.timeout_tick) == (1.0 / 36.0, 9, 2160)
    p = _lib.sim_params(True)
    assert (p.dt, p.ticks_per_step) == (1.0 / 60.0, 15) and p.timeout_tick in (3600, 3601)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from gym_highway_b200 import _lib
    import pytest
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.HighwayLibraryError):
        _lib.load()


def test_binding_constants_mirror_the_header():
    """every HW_FLAG_* / dimension the Python binding uses has the header's value"""
    from gym_highway_b200 import _lib
    text = open(os.path.join(ROOT, "include", "highway_b200.h")).read()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+(HW_[A-Z0-9_]+)\s+(\(?-?[0-9xXa-fA-F]+\)?)\s", text)
            if re.fullmatch(r"-?(0[xX][0-9a-fA-F]+|\d+)", m.group(2))}
    for py, c in (("FLAG_CONTINUOUS", "HW_FLAG_CONTINUOUS"), ("FLAG_NO_INF_OBS", "HW_FLAG_NO_INF_OBS"), ("FLAG_NO_AUTORESET", "HW_FLAG_NO_AUTORESET"),
                  ("FLAG_SHARED_REWARD", "HW_FLAG_SHARED_REWARD"), ("FLAG_EXACT_STEERING", "HW_FLAG_EXACT_STEERING"),
                  ("FLAG_MA_THREAD_PER_ENV", "HW_FLAG_MA_THREAD_PER_ENV"), ("FLAG_MA_LANE_PER_CAR", "HW_FLAG_MA_LANE_PER_CAR"),
                  ("FLAG_COMPACT_OBS", "HW_FLAG_COMPACT_OBS"), ("SA_OBS_DIM", "HW_SA_OBS_DIM"), ("SA_OBS_DIM_COMPACT", "HW_SA_OBS_DIM_COMPACT"),
                  ("ABI_VERSION", "HW_ABI_VERSION")):
        assert getattr(_lib, py) == defs[c], (py, c)
    assert len(_lib.SA_COMPACT_COLUMNS) == _lib.SA_OBS_DIM_COMPACT
    assert sorted(set(range(_lib.SA_OBS_DIM)) - set(_lib.SA_COMPACT_COLUMNS)) == [11, 13, 15, 17]


def test_expand_obs_restores_the_reference_layout():
    import numpy as np
    from gym_highway_b200 import _lib
    from gym_highway_b200.vec_env import HighwayVecEnv
    full = np.random.RandomState(0).rand(9, _lib.SA_OBS_DIM).astype(np.float32)
    full[:, 11:18:2] = 0.5
    assert np.array_equal(HighwayVecEnv.expand_obs(full[:, list(_lib.SA_COMPACT_COLUMNS)]), full)
