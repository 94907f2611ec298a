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


# ----------------------------------------------------------------------------- multi-agent
def unpack_ma(cars, obst, head, A, K):
    """packed numpy arrays (cars f32[A*9, n], obst f32[K, n, 4], head i32/u32[10, n]) -> canonical dict"""
    cars = np.ascontiguousarray(cars, np.float32).reshape(A, 9, -1)
    n = cars.shape[2]
    head = np.ascontiguousarray(head).view(np.uint32).reshape(10, n)
    bits = np.ascontiguousarray(cars[:, 8, :]).view(np.uint32)              # [A, n]
    car = np.transpose(cars[:, 0:8, :], (2, 0, 1)).astype(np.float64)       # [n, A, 8]
    cflags = (((bits & BIT_LEFT) != 0) * 1 | ((bits & BIT_RIGHT) != 0) * 2 | ((bits & BIT_DO_ACC) != 0) * 4
              | ((bits & BIT_DO_MAINT) != 0) * 8)
    nobs = ((head[0] >> 20) & 31).astype(np.int32)
    k = np.arange(K)[None, :]
    live = k < nobs[:, None]
    obst_c = np.transpose(np.asarray(obst, np.float32).reshape(K, n, 4), (1, 0, 2)).astype(np.float64)
    obst_c = np.where(live[:, :, None], obst_c, 0.0)
    olane = np.where(live, (head[2][:, None] >> (2 * k)) & 3, 0).astype(np.int32)
    ofresh = np.where(live, (head[3][:, None] >> k) & 1, 0).astype(np.int32)
    newest = np.stack([(head[1] >> (5 * l)) & 31 for l in range(3)], axis=1).astype(np.int32)
    newest = np.where(newest == 31, -1, newest)
    ep_ret = (head[8].astype(np.uint64) | (head[9].astype(np.uint64) << np.uint64(32))).view(np.float64)
    return dict(car=car, lane=((bits >> BIT_LANE_SHIFT) & 3).T.astype(np.int32), cflags=cflags.T.astype(np.int32),
                tick=(head[0] & 0xFFFF).astype(np.int32), leader=((head[0] >> 16) & 7).astype(np.int32), nobs=nobs,
                obst=obst_c, olane=olane, ofresh=ofresh, newest=newest.astype(np.int32),
                nc_obs=head[4].astype(np.int32), nc_agent=head[5].astype(np.int32), spawn_ctr=head[6].astype(np.uint32),
                ep_len=head[7].astype(np.int32), ep_ret=ep_ret.copy(), overflow=((head[3] >> 16) & 0xFF).astype(np.int32),
                done_mask=((head[3] >> 24) & 0xFF).astype(np.int32))


def pack_ma(c, A, K):
    """canonical dict -> packed numpy arrays (cars f32[A*9, n], obst f32[K, n, 4], head i32[10, n])"""
    car = np.asarray(c["car"], np.float64)
    n = car.shape[0]
    f = np.asarray(c["cflags"]).astype(np.uint32).T                         # [A, n]
    bits = ((np.asarray(c["lane"]).astype(np.uint32).T << BIT_LANE_SHIFT)
            | np.where(f & 1, BIT_LEFT, 0).astype(np.uint32) | np.where(f & 2, BIT_RIGHT, 0).astype(np.uint32)
            | np.where(f & 4, BIT_DO_ACC, 0).astype(np.uint32) | np.where(f & 8, BIT_DO_MAINT, 0).astype(np.uint32))
    cars = np.zeros((A, 9, n), np.float32)
    cars[:, 0:8, :] = np.transpose(car, (1, 2, 0))
    cars[:, 8, :] = bits.view(np.float32)
    obst = np.ascontiguousarray(np.transpose(np.asarray(c["obst"], np.float64), (1, 0, 2))).astype(np.float32)
    head = np.zeros((10, n), np.uint32)
    nobs = np.asarray(c["nobs"]).astype(np.uint32)
    head[0] = (np.asarray(c["tick"]).astype(np.uint32) & 0xFFFF) | (np.asarray(c["leader"]).astype(np.uint32) << 16) | (nobs << 20)
    newest = np.asarray(c["newest"]).astype(np.int64)
    newest = np.where(newest < 0, 31, newest).astype(np.uint32)
    head[1] = newest[:, 0] | (newest[:, 1] << 5) | (newest[:, 2] << 10)
    k = np.arange(K, dtype=np.uint32)[None, :]
    live = k < nobs[:, None]
    head[2] = np.sum(np.where(live, np.asarray(c["olane"]).astype(np.uint32) << (2 * k), 0), axis=1, dtype=np.uint64).astype(np.uint32)
    head[3] = (np.sum(np.where(live, (np.asarray(c["ofresh"]).astype(np.uint32) & 1) << k, 0), axis=1, dtype=np.uint64).astype(np.uint32)
               | ((np.asarray(c.get("overflow", np.zeros(n))).astype(np.uint32) & 0xFF) << 16)
               | ((np.asarray(c.get("done_mask", np.zeros(n))).astype(np.uint32) & 0xFF) << 24))
    head[4] = np.asarray(c["nc_obs"]).astype(np.uint32)
    head[5] = np.asarray(c["nc_agent"]).astype(np.uint32)
    head[6] = np.asarray(c["spawn_ctr"]).astype(np.uint32)
    head[7] = np.asarray(c.get("ep_len", np.zeros(n))).astype(np.uint32)
    er = np.ascontiguousarray(c.get("ep_ret", np.zeros(n)), np.float64).view(np.uint64)
    head[8] = (er & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    head[9] = (er >> np.uint64(32)).astype(np.uint32)
    return dict(cars=cars.reshape(A * 9, n), obst=obst, head=head.view(np.int32))
