"""Packed <-> canonical views of the environment state.

Packed = what the kernels read (fp32 structure of arrays, include/highway_b200.h); canonical = one
plain numpy array per quantity (float64 / int), convenient for checkpoints, debugging and for
comparing against an external implementation.
"""
import numpy as np

BIT_TICK_MASK = 0xFFFF
BIT_LANE_SHIFT = 16
BIT_LEFT, BIT_RIGHT, BIT_DO_ACC, BIT_DO_MAINT = 1 << 18, 1 << 19, 1 << 20, 1 << 21


def unpack_sa(car_a, car_b, obst, stat):
    """packed numpy arrays -> canonical dict (car f64[n,7], lane, flags, tick, obst f64[n,3,4], ...)"""
    car_a = np.asarray(car_a, np.float32)
    car_b = np.ascontiguousarray(car_b, np.float32)
    bits = car_b[:, 3].copy().view(np.uint32)
    n = car_a.shape[0]
    car = np.zeros((n, 7), np.float64)
    car[:, 0:4] = car_a            # px, py, raw_x, vx
    car[:, 4:7] = car_b[:, 0:3]    # acc, steer, cruise
    flags = (((bits & BIT_LEFT) != 0) * 1 | ((bits & BIT_RIGHT) != 0) * 2 | ((bits & BIT_DO_ACC) != 0) * 4
             | ((bits & BIT_DO_MAINT) != 0) * 8)
    stat = np.asarray(stat).astype(np.int32).view(np.uint32)
    return dict(car=car,
                lane=((bits >> BIT_LANE_SHIFT) & 3).astype(np.int32),
                flags=flags.astype(np.int32),
                tick=(bits & BIT_TICK_MASK).astype(np.int32),
                obst=np.ascontiguousarray(np.transpose(np.asarray(obst, np.float32), (1, 0, 2))).astype(np.float64),
                ncoll=stat[:, 0].astype(np.int32),
                spawn_ctr=stat[:, 1].astype(np.uint32),
                ep_ret_milli=stat[:, 2].view(np.int32).astype(np.int64),
                ep_len=stat[:, 3].astype(np.int32))


def pack_sa(c):
    """canonical dict -> packed numpy arrays (car_a f32[n,4], car_b f32[n,4], obst f32[3,n,4], stat i32[n,4])"""
    car = np.asarray(c["car"], np.float64)
    n = car.shape[0]
    f = np.asarray(c["flags"]).astype(np.uint32)
    bits = ((np.asarray(c["tick"]).astype(np.uint32) & BIT_TICK_MASK)
            | (np.asarray(c["lane"]).astype(np.uint32) << BIT_LANE_SHIFT)
            | np.where(f & 1, BIT_LEFT, 0).astype(np.uint32) | np.where(f & 2, BIT_RIGHT, 0).astype(np.uint32)
            | np.where(f & 4, BIT_DO_ACC, 0).astype(np.uint32) | np.where(f & 8, BIT_DO_MAINT, 0).astype(np.uint32))
    car_a = car[:, 0:4].astype(np.float32)
    car_b = np.zeros((n, 4), np.float32)
    car_b[:, 0:3] = car[:, 4:7]
    car_b[:, 3] = bits.view(np.float32)
    obst = np.ascontiguousarray(np.transpose(np.asarray(c["obst"], np.float64), (1, 0, 2))).astype(np.float32)
    stat = np.zeros((n, 4), np.uint32)
    stat[:, 0] = np.asarray(c["ncoll"]).astype(np.uint32)
    stat[:, 1] = np.asarray(c["spawn_ctr"]).astype(np.uint32)
    stat[:, 2] = np.asarray(c["ep_ret_milli"]).astype(np.int32).view(np.uint32)
    stat[:, 3] = np.asarray(c["ep_len"]).astype(np.uint32)
    return dict(car_a=car_a, car_b=car_b, obst=obst, stat=stat.view(np.int32))
