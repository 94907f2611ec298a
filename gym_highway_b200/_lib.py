"""ctypes binding of the C ABI declared in include/highway_b200.h.

There is deliberately no fallback: if libhighway_b200.so is missing or does not export the ABI this
module raises, and every env in this package with it.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhighway_b200.so")

ABI_VERSION = 3

FLAG_CONTINUOUS = 1
FLAG_NO_INF_OBS = 2
FLAG_NO_AUTORESET = 4
FLAG_SHARED_REWARD = 16
FLAG_EXACT_STEERING = 32
FLAG_MA_THREAD_PER_ENV = 64
FLAG_MA_LANE_PER_CAR = 128
FLAG_COMPACT_OBS = 256
MA_MAX_AGENTS = 5
MA_MAX_AGENTS_EXT = 8
MA_MAX_SLOTS = 16

SA_OBS_DIM = 18
SA_OBS_DIM_COMPACT = 14
SA_COMPACT_COLUMNS = (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16)  # the four dropped columns (11, 13, 15, 17) are the constant 0.5


class HwSimParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("ticks_per_step", C.c_int32), ("timeout_tick", C.c_int32)]


class HwSaState(C.Structure):
    _fields_ = [("car_a", C.c_void_p), ("car_b", C.c_void_p), ("obst", C.c_void_p), ("stat", C.c_void_p)]


class HwSaOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("rew", C.c_void_p), ("done", C.c_void_p), ("term", C.c_void_p),
                ("stats", C.c_void_p)]


class HwMaState(C.Structure):
    _fields_ = [("cars", C.c_void_p), ("obst", C.c_void_p), ("head", C.c_void_p),
                ("num_agents", C.c_int32), ("num_slots", C.c_int32)]


class HwMaOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("rew", C.c_void_p), ("done", C.c_void_p), ("term", C.c_void_p),
                ("stats", C.c_void_p)]


class HwRing(C.Structure):
    _fields_ = [("obs0", C.c_void_p), ("obs1", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p),
                ("terminals1", C.c_void_p), ("state", C.c_void_p)]


class HwTransitionBlock(C.Structure):
    _fields_ = [("obs0", C.c_void_p), ("stride_obs0", C.c_int64), ("obs1", C.c_void_p), ("stride_obs1", C.c_int64),
                ("actions", C.c_void_p), ("stride_actions", C.c_int64), ("rewards", C.c_void_p), ("stride_rewards", C.c_int64),
                ("done", C.c_void_p), ("stride_done", C.c_int64), ("done_is_float", C.c_int32)]


class HwEpisodeTracker(C.Structure):
    _fields_ = [("ep_reward", C.c_void_p), ("ep_step", C.c_void_p), ("fin_reward", C.c_void_p), ("fin_step", C.c_void_p),
                ("epoch_reward", C.c_void_p), ("epoch_counts", C.c_void_p)]


class HighwayLibraryError(RuntimeError):
    pass


_EXPORTS = {
    # name: (restype, argtypes)
    "hw_abi_version": (C.c_int, []),
    "hw_error_string": (C.c_char_p, [C.c_int]),
    "hw_sa_reset": (C.c_int, [C.POINTER(HwSaState), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hw_sa_step": (C.c_int, [C.POINTER(HwSaState), C.c_void_p, C.POINTER(HwSaOut), C.POINTER(HwSimParams),
                             C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_void_p]),
    "hw_sa_observe": (C.c_int, [C.POINTER(HwSaState), C.c_void_p, C.c_int64, C.c_void_p]),
    "hw_sa_rollout_random": (C.c_int, [C.POINTER(HwSaState), C.POINTER(HwSaOut), C.POINTER(HwSimParams),
                                       C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "hw_ma_reset": (C.c_int, [C.POINTER(HwMaState), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hw_ma_step": (C.c_int, [C.POINTER(HwMaState), C.c_void_p, C.POINTER(HwMaOut), C.POINTER(HwSimParams),
                             C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_void_p]),
    "hw_ma_observe": (C.c_int, [C.POINTER(HwMaState), C.c_void_p, C.c_int64, C.c_void_p]),
    "hw_ma_step_host": (C.c_int, [C.POINTER(HwMaState), C.c_void_p, C.c_void_p, C.POINTER(HwMaOut),
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(HwSimParams),
                                  C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_void_p]),
    "hw_policy_sample_gaussian2": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hw_ring_store": (C.c_int, [C.POINTER(HwRing), C.POINTER(HwTransitionBlock), C.POINTER(HwEpisodeTracker), C.c_int64, C.c_int,
                                C.c_int, C.c_int, C.c_int64, C.c_float, C.c_void_p]),
    "hw_obs_moments_workspace": (C.c_int64, [C.c_int64, C.c_int]),
    "hw_obs_moments": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hw_nstep_returns": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "hw_gae": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                         C.c_int64, C.c_void_p]),
    "hw_cast_pad_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "hw_policy_sample_discrete": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hw_sa_step_host": (C.c_int, [C.POINTER(HwSaState), C.c_void_p, C.c_void_p, C.POINTER(HwSaOut),
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(HwSimParams),
                                  C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_EXPORTS)


def load():
    """dlopen the library and bind every declared entry point (raises if anything is missing)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HighwayLibraryError(
            "%s is missing: build it with `python -m gym_highway_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _EXPORTS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise HighwayLibraryError("%s does not export %s" % (LIB_PATH, name))
        fn.restype = res
        fn.argtypes = args
    if lib.hw_abi_version() != ABI_VERSION:
        raise HighwayLibraryError("ABI version mismatch: library %d, binding %d" % (lib.hw_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def launch_ctx(device_index, name):
    """context of one library call: the CUDA device of the environment made current (only when it is not - the usual case costs a
    comparison instead of torch.cuda.device's enter / exit) and, with HW_NVTX=1, an NVTX range around the launch"""
    import torch
    if torch.cuda.current_device() != device_index:
        return _LaunchCtx(device_index, name)
    return nvtx_range(name)


class _LaunchCtx(object):
    def __init__(self, device_index, name):
        import torch
        self._dev, self._rng = torch.cuda.device(device_index), nvtx_range(name)

    def __enter__(self):
        self._dev.__enter__(); self._rng.__enter__()

    def __exit__(self, *exc):
        self._rng.__exit__(*exc); self._dev.__exit__(*exc)
        return False


def raw_stream(device_index):
    """cudaStream_t of torch's current stream on `device_index` as an int (the C call torch.cuda.current_stream() wraps: no Stream object)"""
    import torch
    return torch._C._cuda_getCurrentRawStream(device_index)


def check(rc, what):
    if rc != 0:
        msg = load().hw_error_string(rc)
        raise HighwayLibraryError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def sim_params(real_time=False):
    """dt, ticks per env.step and the time-out tick exactly as the reference derives them:
    ticks = 60 if real_time else 36 (multi_lane_sim.py:358), dt = 1/ticks (:888),
    int(ticks/4) ticks per step (highway_env.py:101), done when the accumulated run_time >= 60 (:910-913)."""
    ticks = 60.0 if real_time else 36.0
    dt = 1.0 / ticks
    run_time, t = 0.0, 0
    while not run_time >= 60:
        run_time += dt
        t += 1
    return HwSimParams(dt, int(ticks / 4), t)


class _NoRange(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_RANGE = _NoRange()


def nvtx_range(name):
    """`with nvtx_range("hw_sa_step"):` - an NVTX range around a launch when HW_NVTX=1 (timeline tools: nsys, ncu --nvtx), a no-op
    otherwise (the default: a range costs a microsecond of host time per call)"""
    if os.environ.get("HW_NVTX", "0") != "1":
        return _NO_RANGE
    import torch
    return torch.cuda.nvtx.range(name)
