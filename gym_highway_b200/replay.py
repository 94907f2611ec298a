"""Device-resident transition store and observation statistics for the MADDPG / DDPG consumers.

Mirrors maddpg/trainer/ma_memory.py:4-87 (RingBuffer / Memory: append one transition per env, uniform sampling
`randint(nb_entries - 2)`) and baselines/common/mpi_running_mean_std.py:41-48 (sum / sum of squares / count
accumulated over all ranks) — with the `(nenvs, A, ...)` blocks the multi-agent VecEnv produces appended in one
copy, the ring living on the GPU, and the MPI Allreduce replaced by one torch.distributed all_reduce.
"""
import torch
import torch.distributed as dist


class RingMemory(object):
    """one ring buffer per agent, stored as [limit, A, ...] tensors; append takes whole env batches.

    On a CUDA device the append is `hw_ring_store` (csrc/hw_replay.cu): one launch writes the five arrays of the block into
    the ring rows n successive `Memory.append` calls would use and runs the rollout loop's episode bookkeeping, a one-thread
    launch advances start / length, which live on the device (`state`) so that a captured CUDA graph keeps appending; the
    host keeps a mirror of the two words (same arithmetic, no synchronisation).  CPU tensors (host-side tests, offline
    tools) take the indexed torch path below."""

    def __init__(self, limit, num_agents, obs_dim, act_dim, device="cpu", reward_scale=1.0):
        self.limit, self.A = int(limit), int(num_agents)
        self.obs_dim, self.act_dim, self.reward_scale = int(obs_dim), int(act_dim), float(reward_scale)
        self.device = torch.device(device)
        f = dict(dtype=torch.float32, device=self.device)
        self.obs0 = torch.zeros(self.limit, self.A, obs_dim, **f)
        self.obs1 = torch.zeros(self.limit, self.A, obs_dim, **f)
        self.actions = torch.zeros(self.limit, self.A, act_dim, **f)
        self.rewards = torch.zeros(self.limit, self.A, 1, **f)
        self.terminals1 = torch.zeros(self.limit, self.A, 1, **f)
        self.start = 0
        self.length = 0
        self.state = None
        if self.device.type == "cuda":
            from . import _lib
            self._libmod, self._lib = _lib, _lib.load()   # raises when the CUDA library is missing: no fallback on a GPU
            self.state = torch.zeros(3, dtype=torch.int64, device=self.device)
            self._ring_c = _lib.HwRing(self.obs0.data_ptr(), self.obs1.data_ptr(), self.actions.data_ptr(), self.rewards.data_ptr(),
                                       self.terminals1.data_ptr(), self.state.data_ptr())

    @property
    def nb_entries(self):
        return self.length

    def _advance(self, n):
        overflow = max(0, self.length + n - self.limit)
        self.start = (self.start + overflow) % self.limit
        self.length = min(self.limit, self.length + n)

    def append(self, obs0, actions, rewards, obs1, terminals1, tracker=None):
        """obs0/obs1 [N, A, obs_dim], actions [N, A, act_dim], rewards/terminals1 [N, A]: N transitions at once,
        in env order — exactly what N successive Memory.append calls would leave behind"""
        n = obs0.shape[0]
        if self.device.type == "cuda":
            import ctypes as C
            A, D, K = self.A, self.obs_dim, self.act_dim
            for x in (obs0, obs1, actions, rewards):
                if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
                    raise ValueError("RingMemory.append on a CUDA ring takes contiguous float32 CUDA tensors")
            if not (terminals1.is_cuda and terminals1.is_contiguous() and terminals1.dtype in (torch.uint8, torch.bool, torch.float32)):
                raise ValueError("terminals1 must be a contiguous uint8 / bool / float32 CUDA tensor")
            blk = self._libmod.HwTransitionBlock(obs0.data_ptr(), A * D, obs1.data_ptr(), A * D, actions.data_ptr(), A * K,
                                                 rewards.data_ptr(), A, terminals1.data_ptr(), A, int(terminals1.dtype == torch.float32))
            with torch.cuda.device(self.device):
                rc = self._lib.hw_ring_store(C.byref(self._ring_c), C.byref(blk), C.byref(tracker) if tracker is not None else None,
                                             n, A, D, K, self.limit, self.reward_scale,
                                             C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            self._libmod.check(rc, "hw_ring_store")
            self._advance(n)
            return
        m = n
        if n > self.limit:
            obs0, actions, rewards, obs1, terminals1 = (x[-self.limit:] for x in (obs0, actions, rewards, obs1, terminals1))
            m = self.limit
        pos = (self.start + self.length + (n - m)) % self.limit
        idx = (pos + torch.arange(m, device=self.obs0.device)) % self.limit
        self.obs0[idx] = obs0.to(self.obs0.dtype)
        self.obs1[idx] = obs1.to(self.obs1.dtype)
        self.actions[idx] = actions.to(self.actions.dtype).reshape(m, self.A, -1)
        self.rewards[idx] = (rewards.to(self.rewards.dtype) * self.reward_scale).reshape(m, self.A, 1)
        self.terminals1[idx] = terminals1.to(self.terminals1.dtype).reshape(m, self.A, 1)
        self._advance(n)

    def get(self, i):
        """i-th oldest entry (RingBuffer.__getitem__)"""
        if i < 0 or i >= self.length:
            raise KeyError()
        j = (self.start + i) % self.limit
        return dict(obs0=self.obs0[j], obs1=self.obs1[j], actions=self.actions[j], rewards=self.rewards[j], terminals1=self.terminals1[j])

    def sample(self, batch_size, index=None, generator=None):
        if index is None:
            index = torch.randint(0, max(self.nb_entries - 2, 1), (batch_size,), device=self.obs0.device, generator=generator)
        j = (self.start + index) % self.limit
        return dict(obs0=self.obs0[j], obs1=self.obs1[j], rewards=self.rewards[j], actions=self.actions[j], terminals1=self.terminals1[j])


class RunningMeanStd(object):
    """sum / sumsq / count in float64, all-reduced over ranks on update (mpi_running_mean_std.py)"""

    def __init__(self, shape=(), epsilon=1e-2, device="cpu"):
        self._sum = torch.zeros(shape, dtype=torch.float64, device=device)
        self._sumsq = torch.full(shape, epsilon, dtype=torch.float64, device=device)
        self._count = torch.tensor(epsilon, dtype=torch.float64, device=device)
        self.shape = tuple(shape)

    def update(self, x, group=None, repeat=1):
        """x: [..., *shape] rows; `repeat`: the reference's MADDPG loop feeds every env row once per agent into the one obs_rms
        all agents share (ma_ddpg_learner.py:398-404), which scales sum, sum of squares and count alike"""
        x = x.reshape(-1, *self.shape)
        rows = x.shape[0]
        if x.is_cuda:
            import ctypes as C
            from . import _lib
            lib = _lib.load()
            x = x.contiguous().to(torch.float32)
            cols = max(1, int(self._sum.numel()))
            ws = torch.empty(max(1, lib.hw_obs_moments_workspace(rows, cols) // 8), dtype=torch.float64, device=x.device)
            vec = torch.empty(2 * cols + 1, dtype=torch.float64, device=x.device)
            with torch.cuda.device(x.device):
                rc = lib.hw_obs_moments(C.c_void_p(x.data_ptr()), rows, cols, float(rows), C.c_void_p(ws.data_ptr()), C.c_void_p(vec.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
            _lib.check(rc, "hw_obs_moments")
        else:
            x = x.to(torch.float64)
            vec = torch.cat([x.sum(0).reshape(-1), (x * x).sum(0).reshape(-1), torch.tensor([float(rows)], dtype=torch.float64, device=x.device)])
        if repeat != 1:
            vec = vec * float(repeat)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
        n = self._sum.numel()
        self._sum += vec[0:n].reshape(self._sum.shape)
        self._sumsq += vec[n:2 * n].reshape(self._sumsq.shape)
        self._count += vec[2 * n]

    @property
    def mean(self):
        return (self._sum / self._count).to(torch.float32)

    @property
    def std(self):
        m = self._sum / self._count
        return torch.sqrt(torch.clamp(self._sumsq / self._count - m * m, min=1e-2)).to(torch.float32)


class MADeviceCollector(object):
    """The rollout half of MADDPG's training loop (maddpg/trainer/ma_ddpg.py:193-265) on the device: for nb_rollout_steps,
    actions for every (env, agent) from `policy(obs [N, A, D]) -> [N, A, act_dim]` (any torch callable: the stacked per-agent
    actors), one `HighwayMAVecEnv` step writing into the collector's own buffers (`step_into`), and ONE `hw_ring_store` that
    appends the (obs, action, reward * reward_scale, obs', done) block to every agent's ring, adds the rewards to the
    per-env episode sums and closes the episodes that ended (episode_reward / episode_step / epoch sums, :228-262).
    `obs_rms` (a `RunningMeanStd` of shape (A, D)) is updated per step with the block's float64 moments
    (`hw_obs_moments`) and ONE all_reduce over the ranks, like mpi_running_mean_std.py:41-48.
    `central=True` all_gathers the blocks over the ranks first (one NCCL all_gather of a packed [N, W] buffer per step), so
    that every rank's ring is the central replay holding all ranks' transitions in global env order."""

    def __init__(self, env, policy, memory, obs_rms=None, central=False, group=None):
        from . import _lib
        self._libmod = _lib
        self.env, self.policy, self.memory, self.obs_rms, self.central, self.group = env, policy, memory, obs_rms, bool(central), group
        self.A, self.D = env.num_agents, env.obs_dim
        self.act_dim = memory.act_dim
        n, A, dev = env.num_envs, self.A, env.device
        self.world = dist.get_world_size(group) if (self.central and dist.is_available() and dist.is_initialized()) else 1
        self.obs = [torch.empty(n, A, self.D, dtype=torch.float32, device=dev) for _ in range(2)]
        self.rew = torch.empty(n, A, dtype=torch.float32, device=dev)
        self.done = torch.empty(n, A, dtype=torch.uint8, device=dev)
        self.cur = 0
        self.obs[0].copy_(env.reset())
        m = n * self.world                       # envs whose transitions this rank's ring receives per step
        self.ep_reward = torch.zeros(m, A, dtype=torch.float32, device=dev)
        self.ep_step = torch.zeros(m, dtype=torch.int64, device=dev)
        self.fin_reward = torch.zeros(m, A, dtype=torch.float32, device=dev)
        self.fin_step = torch.zeros(m, dtype=torch.int64, device=dev)
        self.epoch_reward = torch.zeros(A, dtype=torch.float64, device=dev)
        self.epoch_counts = torch.zeros(2, dtype=torch.int64, device=dev)
        self._tracker = _lib.HwEpisodeTracker(self.ep_reward.data_ptr(), self.ep_step.data_ptr(), self.fin_reward.data_ptr(),
                                              self.fin_step.data_ptr(), self.epoch_reward.data_ptr(), self.epoch_counts.data_ptr())
        self.t = 0

    @torch.no_grad()
    def step(self):
        """one rollout step; returns the action block that was executed"""
        env, A = self.env, self.A
        obs0, obs1 = self.obs[self.cur], self.obs[self.cur ^ 1]
        actions = self.policy(obs0).to(torch.float32).reshape(env.num_envs, A, self.act_dim).contiguous()
        if env.continuous:
            env_act = actions
        else:
            env_act = actions.argmax(dim=-1).to(torch.int32)       # np.argmax(action_n[id]), multiagent_envs/highway_env.py:78
        env.step_into(env_act.contiguous(), obs1, self.rew, self.done)
        if self.obs_rms is not None:
            self.obs_rms.update(obs0, group=self.group, repeat=A)
        if self.world > 1:
            W = A * (2 * self.D + self.act_dim + 2)
            packed = torch.cat([obs0.reshape(-1, A * self.D), obs1.reshape(-1, A * self.D), actions.reshape(-1, A * self.act_dim),
                                self.rew, self.done.to(torch.float32)], dim=1)
            allp = torch.empty(self.world * env.num_envs, W, dtype=torch.float32, device=env.device)
            dist.all_gather_into_tensor(allp, packed, group=self.group)
            self._store_packed(allp, W)
        else:
            self.memory.append(obs0, actions, self.rew, obs1, self.done, tracker=self._tracker)
        self.cur ^= 1
        self.t += 1
        return actions

    def _store_packed(self, allp, W):
        import ctypes as C
        A, D, K, mem = self.A, self.D, self.act_dim, self.memory
        base, fs = allp.data_ptr(), 4
        o1, ac, rw, dn = A * D, 2 * A * D, 2 * A * D + A * K, 2 * A * D + A * K + A
        blk = self._libmod.HwTransitionBlock(base, W, base + o1 * fs, W, base + ac * fs, W, base + rw * fs, W, base + dn * fs, W, 1)
        with torch.cuda.device(mem.device):
            rc = mem._lib.hw_ring_store(C.byref(mem._ring_c), C.byref(blk), C.byref(self._tracker), allp.shape[0], A, D, K, mem.limit,
                                        mem.reward_scale, C.c_void_p(torch.cuda.current_stream(mem.device).cuda_stream))
        self._libmod.check(rc, "hw_ring_store")
        mem._advance(allp.shape[0])

    def collect(self, nb_rollout_steps):
        for _ in range(nb_rollout_steps):
            self.step()

    @torch.no_grad()
    def collect_graphed(self, nb_rollout_steps):
        """`collect(nb_rollout_steps)` replayed from ONE CUDA graph (captured on the first call after an eager warm-up run): eagerly a
        step is a dozen small launches and the host is the bound.  The ring's start / length live on the device (`RingMemory.state`),
        so every replay appends where the last one stopped; the host mirror is advanced by the same arithmetic.  The policy must be
        capturable (device tensors only, no host-side state that changes per step); nb_rollout_steps must be even (the two
        observation buffers alternate, and a replay has to start on the one the capture started on)."""
        T = int(nb_rollout_steps)
        if T % 2:
            raise ValueError("collect_graphed needs an even nb_rollout_steps")
        if getattr(self, "_graph", None) is None or self._graph_T != T:
            self.collect(T)  # warm-up: allocator pools, cuBLAS workspaces, NCCL channels
            torch.cuda.synchronize(self.env.device)
            mirror = (self.memory.start, self.memory.length, self.t)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self.collect(T)
            self.memory.start, self.memory.length, self.t = mirror  # the capture recorded launches, it did not run them
            self._graph_T = T
        self._graph.replay()
        self.memory._advance(T * self.env.num_envs * self.world)
        self.t += T

    def finished_episodes(self):
        """(episode_reward [k, A], episode_step [k], env index [k]) of the episodes that ended on the LAST step (host copy of
        those rows only) - what the reference appends to its history deques"""
        idx = torch.nonzero(self.fin_step).squeeze(1)
        return self.fin_reward[idx].cpu(), self.fin_step[idx].cpu(), idx.cpu()

    def epoch_statistics(self, reset=False):
        """epoch_episode_rewards [A], episodes, sum of episode steps since the last reset (device accumulators)"""
        out = dict(epoch_episode_rewards=self.epoch_reward.cpu().clone(), episodes=int(self.epoch_counts[0]), episode_steps=int(self.epoch_counts[1]))
        if reset:
            self.epoch_reward.zero_(); self.epoch_counts.zero_()
        return out
