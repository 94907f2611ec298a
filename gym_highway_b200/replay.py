"""Device-resident transition store and observation statistics for the MADDPG / DDPG consumers.

Mirrors maddpg/trainer/ma_memory.py:4-87 (RingBuffer / Memory: append one transition per env, uniform sampling
`randint(nb_entries - 2)`) and baselines/common/mpi_running_mean_std.py:41-48 (sum / sum of squares / count
accumulated over all ranks) — with the `(nenvs, A, ...)` blocks the multi-agent VecEnv produces appended in one
copy, the ring living on the GPU, and the MPI Allreduce replaced by one torch.distributed all_reduce.
"""
import torch
import torch.distributed as dist


class RingMemory(object):
    """one ring buffer per agent, stored as [limit, A, ...] tensors; append takes whole env batches"""

    def __init__(self, limit, num_agents, obs_dim, act_dim, device="cpu"):
        self.limit, self.A = int(limit), int(num_agents)
        f = dict(dtype=torch.float32, device=device)
        self.obs0 = torch.zeros(self.limit, self.A, obs_dim, **f)
        self.obs1 = torch.zeros(self.limit, self.A, obs_dim, **f)
        self.actions = torch.zeros(self.limit, self.A, act_dim, **f)
        self.rewards = torch.zeros(self.limit, self.A, 1, **f)
        self.terminals1 = torch.zeros(self.limit, self.A, 1, **f)
        self.start = 0
        self.length = 0

    @property
    def nb_entries(self):
        return self.length

    def append(self, obs0, actions, rewards, obs1, terminals1):
        """obs0/obs1 [N, A, obs_dim], actions [N, A, act_dim], rewards/terminals1 [N, A]: N transitions at once,
        in env order — exactly what N successive Memory.append calls would leave behind"""
        n = obs0.shape[0]
        if n > self.limit:
            obs0, actions, rewards, obs1, terminals1 = (x[-self.limit:] for x in (obs0, actions, rewards, obs1, terminals1))
            n = self.limit
        pos = (self.start + self.length) % self.limit
        idx = (pos + torch.arange(n, device=self.obs0.device)) % self.limit
        self.obs0[idx] = obs0.to(self.obs0.dtype)
        self.obs1[idx] = obs1.to(self.obs1.dtype)
        self.actions[idx] = actions.to(self.actions.dtype).reshape(n, self.A, -1)
        self.rewards[idx] = rewards.to(self.rewards.dtype).reshape(n, self.A, 1)
        self.terminals1[idx] = terminals1.to(self.terminals1.dtype).reshape(n, self.A, 1)
        overflow = max(0, self.length + n - self.limit)
        self.start = (self.start + overflow) % self.limit
        self.length = min(self.limit, self.length + n)

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

    def update(self, x, group=None):
        x = x.reshape(-1, *self.shape).to(torch.float64)
        vec = torch.cat([x.sum(0).reshape(-1), (x * x).sum(0).reshape(-1), torch.tensor([float(x.shape[0])], dtype=torch.float64, device=x.device)])
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
