"""Drive the UNMODIFIED reference (/root/reference) under oracle/shim and move state in and out of it.

TEST INFRASTRUCTURE ONLY, and only usable in the build container (the GPU box has no
/root/reference): oracle/gen_golden.py uses it to write tests/golden/*.npz, and
tests/test_reference_live.py (skipped when the reference is absent) replays random trajectories.

What is patched, and why:
  * `highway_core.Action = actions.Action`   - the discrete multi-agent env raises NameError otherwise
    (highway_core.py:119-130 uses a name that is never imported).
  * `random.uniform`                          - replaced while stepping by a provider that returns
    a + (b-a)*u with u taken from the Philox stream keyed by (seed, env_id, spawn_ctr), so the
    reference, the C oracle and the CUDA kernel consume identical uniforms.
Nothing else in the reference is touched; state is written through ordinary attribute assignment.
"""
import os
import random as _random
import sys

import numpy as np

from . import oracle as _orc

REFERENCE_ROOT = os.environ.get("HW_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")

TRAIN_KW = {'manual': False, 'inf_obs': True, 'save': False, 'render': False, 'real_time': False}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_highway"))


_mods = None


def load():
    """import the reference env modules (idempotent)"""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
        for p in (REFERENCE_ROOT, _SHIM):
            if p not in sys.path:
                sys.path.insert(0, p)
        import models.config  # noqa: F401  (simple_base imports it; needs no TensorFlow)
        import gym_highway.envs as sa
        import gym_highway.multiagent_envs as ma
        from gym_highway.multiagent_envs import highway_core, actions
        highway_core.Action = actions.Action
        _mods = dict(sa=sa, ma=ma, core=highway_core, actions=actions)
    return _mods


class PhiloxUniform(object):
    """stand-in for random.uniform: spawn k of env `env_id` gets the two uniforms of Philox counter k"""

    def __init__(self, seed, env_id, spawn_ctr=0):
        self.seed, self.env_id, self.spawn_ctr = int(seed), int(env_id), int(spawn_ctr)
        self._pending = None
        self.drawn = []

    def __call__(self, a, b):
        if self._pending is None:
            assert (a, b) == (70, 100), (a, b)
            u1, u2 = _orc.spawn_uniforms(self.seed, self.env_id, self.spawn_ctr)
            self._pending = u2
            self.drawn += [u1, u2]
            return a + (b - a) * u1
        assert (a, b) == (5, 7), (a, b)
        u2, self._pending = self._pending, None
        self.spawn_ctr += 1
        return a + (b - a) * u2


class _patched_uniform(object):
    def __init__(self, provider):
        self.provider = provider

    def __enter__(self):
        self._old = _random.uniform
        _random.uniform = self.provider

    def __exit__(self, *a):
        _random.uniform = self._old


def _run_time_after(ticks, dt):
    t = 0.0
    for _ in range(int(ticks)):
        t += dt
    return t


def _set_rect(sprite):
    sprite.rect.x = sprite.position.x * 32 - sprite.rect.width / 2
    sprite.rect.y = sprite.position.y * 32 - sprite.rect.height / 2


class RefSingleAgent(object):
    """one reference HighwayEnv / HighwayEnvContinuous with state injection"""

    def __init__(self, continuous=False, real_time=False, inf_obs=True):
        m = load()
        kw = dict(TRAIN_KW, real_time=real_time, inf_obs=inf_obs)
        self.env = (m['sa'].HighwayEnvContinuous if continuous else m['sa'].HighwayEnv)(**kw)
        self.continuous = continuous
        self.sim = self.env.env
        self.dt = 1.0 / self.sim.ticks

    def reset(self):
        return self.env.reset()

    def inject(self, st, i):
        """load row i of an oracle.SaState into a freshly reset reference env"""
        self.env.reset()
        sim = self.sim
        car = sim.reference_car
        px, py, rx, vx, acc, steer, cruise = (float(v) for v in st.car[i])
        car.position.x, car.position.y = px, py
        car.raw_position.x, car.raw_position.y = rx, py
        car.velocity.x, car.velocity.y = vx, 0.0
        car.acceleration, car.steering, car.cruise_vel = acc, steer, cruise
        car.lane_id = int(st.lane[i])
        f = int(st.flags[i])
        car.left_mode, car.right_mode = bool(f & 1), bool(f & 2)
        car.do_accelerate, car.do_maintain = bool(f & 4), bool(f & 8)
        tick = int(st.tick[i])
        if tick > 0:
            _set_rect(car)
        for k in range(3):
            o = sim.lane_max_obs[k]
            opx, orx, ovx, oiv = (float(v) for v in st.obst[i, k])
            o.position.x, o.raw_position.x = opx, orx
            o.velocity.x, o.init_velocity.x = ovx, oiv
            if tick > 0:
                _set_rect(o)
        sim.run_time = _run_time_after(tick, self.dt)
        sim.collision_count_lock = (tick == 0)
        sim.num_obs_collisions = int(st.ncoll[i])
        self._tick = tick

    def extract(self, st, i):
        sim = self.sim
        car = sim.reference_car
        st.car[i] = (car.position.x, car.position.y, car.raw_position.x, car.velocity.x,
                     car.acceleration, car.steering, car.cruise_vel)
        assert car.raw_position.y == car.position.y and car.velocity.y == 0.0
        st.lane[i] = car.lane_id
        st.flags[i] = (1 * car.left_mode) | (2 * car.right_mode) | (4 * car.do_accelerate) | (8 * car.do_maintain)
        for k in range(3):
            o = sim.lane_max_obs[k]
            assert o.lane_id == k + 1 and o.position.y == _orc.LANE_C[k]
            st.obst[i, k] = (o.position.x, o.raw_position.x, o.velocity.x, o.init_velocity.x)
        st.ncoll[i] = sim.num_obs_collisions
        # tick from the reference's own accumulated clock
        t, rt = 0, 0.0
        while rt < sim.run_time:
            rt += self.dt
            t += 1
        assert rt == sim.run_time
        st.tick[i] = t

    def step(self, action, provider):
        """action: int (discrete) or 2 python floats; returns (obs f32[18], reward, done, info)"""
        if self.continuous:
            action = [float(np.float32(action[0])), float(np.float32(action[1]))]
        else:
            action = int(action)
        with _patched_uniform(provider):
            return self.env.step(action)


class RefMultiAgent(object):
    """one reference MultiAgentEnv / MultiAgentEnvContinuous with state injection (A <= 5 cars)"""

    def __init__(self, num_agents, continuous=False, real_time=False, shared_reward=False, inf_obs=True):
        m = load()
        kw = dict(TRAIN_KW, real_time=real_time, inf_obs=inf_obs)
        cls = m['ma'].MultiAgentEnvContinuous if continuous else m['ma'].MultiAgentEnv
        self.env = cls(world_config=kw, num_agents=num_agents, shared_reward=shared_reward)
        self.world = self.env.world
        self.A = num_agents
        self.continuous = continuous
        self.dt = 1.0 / self.world.ticks
        from gym_highway.multiagent_envs.agent import Obstacle
        from gym_highway.multiagent_envs import highway_constants
        import pygame
        self._Obstacle, self._K, self._pygame = Obstacle, highway_constants, pygame

    def reset(self):
        return self.env.reset()

    def inject(self, st, i):
        """load row i of an oracle MaState into a freshly reset reference world"""
        self.env.reset()
        w, A = self.world, self.A
        cars = list(w.agents)
        assert [c.id for c in cars] == list(range(A))
        tick = int(st.tick[i])
        for a, car in enumerate(cars):
            px, py, rx, vx, vy, acc, steer, cruise = (float(v) for v in st.car[i, a])
            car.position.x, car.position.y = px, py
            car.raw_position.x, car.raw_position.y = rx, py
            car.velocity.x, car.velocity.y = vx, vy
            car.acceleration, car.steering, car.cruise_vel = acc, steer, cruise
            car.lane_id = int(st.lane[i, a])
            f = int(st.cflags[i, a])
            car.left_mode, car.right_mode = bool(f & 1), bool(f & 2)
            car.do_accelerate, car.do_maintain = bool(f & 4), bool(f & 8)
            if tick > 0:
                _set_rect(car)
        obstacles = []
        for k in range(int(st.nobs[i])):
            opx, orx, ovx, oiv = (float(v) for v in st.obst[i, k])
            lane = int(st.olane[i, k])
            o = self._Obstacle(id=A + lane - 1, x=opx, y=_orc.LANE_C[lane - 1], raw_x=orx, vel_x=ovx, vel_y=0.0,
                               lane_id=lane, color=self._K.YELLOW)
            o.init_velocity.x = oiv
            if not int(st.ofresh[i, k]):
                _set_rect(o)
            obstacles.append(o)
        G = self._pygame.sprite.Group
        w.scripted_agents = G(*obstacles)
        w.all_obstacles = G(*(cars + obstacles))
        ghost = self._Obstacle(id=A, x=1.0e9, y=0.0, vel_x=0.0, lane_id=1)  # a retired lane_max_obs entry: never respawns
        w.lane_max_obs = [obstacles[int(s)] if int(s) >= 0 else ghost for s in st.newest[i]]
        w.reference_car = cars[int(st.leader[i])]
        w.run_time = _run_time_after(tick, self.dt)
        w.collision_count_lock = (tick == 0)
        w.num_obs_collisions = int(st.nc_obs[i])
        w.num_agent_collisions = int(st.nc_agent[i])
        w.is_done = [False] * A

    def extract(self, st, i):
        w, A = self.world, self.A
        cars = list(w.agents)
        for a, car in enumerate(cars):
            assert car.raw_position.y == car.position.y
            st.car[i, a] = (car.position.x, car.position.y, car.raw_position.x, car.velocity.x, car.velocity.y,
                            car.acceleration, car.steering, car.cruise_vel)
            st.lane[i, a] = car.lane_id
            st.cflags[i, a] = (1 * car.left_mode) | (2 * car.right_mode) | (4 * car.do_accelerate) | (8 * car.do_maintain)
        obstacles = list(w.scripted_agents)
        assert len(obstacles) <= st.K, "K too small for this trajectory: %d live obstacles" % len(obstacles)
        st.nobs[i] = len(obstacles)
        st.obst[i] = 0.0; st.olane[i] = 0; st.ofresh[i] = 0
        for k, o in enumerate(obstacles):
            assert o.position.y == _orc.LANE_C[o.lane_id - 1] and o.id == A + o.lane_id - 1 and o.velocity.y == 0.0
            st.obst[i, k] = (o.position.x, o.raw_position.x, o.velocity.x, o.init_velocity.x)
            st.olane[i, k] = o.lane_id
            st.ofresh[i, k] = int(o.rect.y == 0)
        for lane in range(3):
            o = w.lane_max_obs[lane]
            st.newest[i, lane] = obstacles.index(o) if o in obstacles else -1
        st.leader[i] = w.reference_car.id
        st.nc_obs[i] = w.num_obs_collisions
        st.nc_agent[i] = w.num_agent_collisions
        t, rt = 0, 0.0
        while rt < w.run_time:
            rt += self.dt
            t += 1
        assert rt == w.run_time
        st.tick[i] = t

    def step(self, actions, provider):
        """actions: int[A] (discrete) or float[A][2]; returns (obs_n list of f64 arrays, reward_n, done_n, info)"""
        if self.continuous:
            act = [[float(np.float32(a[0])), float(np.float32(a[1]))] for a in actions]
        else:
            act = [np.eye(4)[int(a)] for a in actions]
        with _patched_uniform(provider):
            ob, r, d, info = self.env.step(act)
        return ob, list(r), list(d), info
