"""On-device rollout collection for the baselines ppo2 / a2c runners.

Mirrors baselines/ppo2/runner.py:20-66 (per-step buffers, GAE(gamma, lam), sf01 flatten, epinfos) and
baselines/a2c/runner.py:21-72 (n-step discounted returns with bootstrap) with every buffer a CUDA tensor
of shape [T, N, ...]: the observations never leave the GPU that produced them.  The policy is a torch MLP
with the reference's shape (`mlp(num_layers=3, num_hidden=256, activation=relu)`, models/config.py:35-38,
baselines/common/models.py:31) and separate policy / value heads; library GEMMs are fine here, the learner
is outside the hot path.
"""
import torch
import torch.nn as nn


class MlpPolicy(nn.Module):
    """18 -> 256 -> 256 -> 256 -> (4 logits | 2 means, 1 value)"""

    def __init__(self, obs_dim=18, n_actions=4, hidden=256, layers=3, continuous=False, bf16_inference=False):
        super().__init__()
        self.bf16_inference = bf16_inference  # rollout-time forward in bf16 autocast (tensor cores); heads read out in fp32
        mods, d = [], obs_dim
        for _ in range(layers):
            mods += [nn.Linear(d, hidden), nn.ReLU()]
            d = hidden
        self.body = nn.Sequential(*mods)
        self.continuous = continuous
        self.pi = nn.Linear(d, 2 if continuous else n_actions)
        self.vf = nn.Linear(d, 1)
        if continuous:
            self.logstd = nn.Parameter(torch.zeros(2))

    def forward(self, obs):
        if self.bf16_inference and obs.is_cuda and not torch.is_grad_enabled():
            return self._forward_bf16(obs)
        h = self.body(obs)
        return self.pi(h), self.vf(h).squeeze(-1)

    def _alloc_bf16(self):
        """persistent bf16 weight buffers, allocated ONCE per device: [in, out] layouts, the input width (18) and the merged
        head width (n_out + 1) zero-padded to multiples of 8 so that cuBLASLt picks its aligned tensor-core kernels (the
        unaligned fallbacks are 3x slower at 10^5 rows).  The tensors are never rebound afterwards - a CUDA graph that
        captured a rollout holds their addresses."""
        lin = [m for m in self.body if isinstance(m, nn.Linear)]
        dev = lin[0].weight.device
        body = []
        for k, m in enumerate(lin):
            kin = m.in_features + ((-m.in_features) % 8 if k == 0 else 0)
            body.append((torch.zeros(kin, m.out_features, dtype=torch.bfloat16, device=dev),
                         torch.zeros(m.out_features, dtype=torch.bfloat16, device=dev)))
        self._n_head = self.pi.out_features + 1
        hp = self._n_head + (-self._n_head) % 8
        self._bf16_body = body
        self._bf16_head_w = torch.zeros(self.pi.in_features, hp, dtype=torch.bfloat16, device=dev)
        self._bf16_head_b = torch.zeros(hp, dtype=torch.bfloat16, device=dev)
        self._bf16_dev = dev
        self._bf16_ver = None

    @torch.no_grad()
    def refresh_bf16_(self):
        """re-cast the fp32 parameters into the persistent bf16 buffers IN PLACE (a handful of small copy kernels).  The
        rollout runners call this at the top of every run(), so the casts are part of a captured CUDA graph and a replay
        after optimizer.step() samples from the updated policy."""
        lin = [m for m in self.body if isinstance(m, nn.Linear)]
        if getattr(self, "_bf16_body", None) is None or self._bf16_dev != lin[0].weight.device:
            self._alloc_bf16()
        for (w, b), m in zip(self._bf16_body, lin):
            w[:m.in_features].copy_(m.weight.t())
            b.copy_(m.bias)
        n = self.pi.out_features
        self._bf16_head_w[:, :n].copy_(self.pi.weight.t())
        self._bf16_head_w[:, n:n + 1].copy_(self.vf.weight.t())
        self._bf16_head_b[:n].copy_(self.pi.bias)
        self._bf16_head_b[n:n + 1].copy_(self.vf.bias)
        self._bf16_ver = tuple(p._version for p in self.parameters())

    def _bf16_weights(self):
        """the persistent bf16 copies, refreshed in place when a parameter has been updated since the last refresh (eager
        use; under graph capture / replay the runners refresh explicitly, the version check is host-side state)"""
        if getattr(self, "_bf16_body", None) is None or self._bf16_ver != tuple(p._version for p in self.parameters()):
            self.refresh_bf16_()
        return self._bf16_body, self._bf16_head_w, self._bf16_head_b

    def _forward_bf16(self, obs):
        """rollout-time forward: one cuBLASLt GEMM per layer with the bias + ReLU epilogue fused (no separate
        elementwise passes over the [N, 256] activations, which at 10^5 rows cost more than the GEMMs), heads merged"""
        body, hw, hb = self._bf16_weights()
        k0 = body[0][0].shape[0]
        if obs.shape[1] == k0:
            h = obs.to(torch.bfloat16)
        else:
            h = obs.new_zeros((obs.shape[0], k0), dtype=torch.bfloat16)
            h[:, :obs.shape[1]] = obs
        for w, b in body:
            h = torch._addmm_activation(b, h, w, use_gelu=False)
        out = torch.addmm(hb, h, hw).float()
        return out[:, :self._n_head - 1], out[:, self._n_head - 1]

    @torch.no_grad()
    def head_bf16(self, obs):
        """[N, 8k] bfloat16: the logits (or means), then the value, then zero padding - the merged head GEMM's output as it
        is, for consumers that read it in place (hw_policy_sample_discrete)"""
        body, hw, hb = self._bf16_weights()
        k0 = body[0][0].shape[0]
        if obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous():
            # one launch for cast + pad (zero-fill + strided copy are two, and the strided one is slow)
            import ctypes as C
            from . import _lib
            h = torch.empty((obs.shape[0], k0), dtype=torch.bfloat16, device=obs.device)
            with torch.cuda.device(obs.device):
                rc = _lib.load().hw_cast_pad_bf16(C.c_void_p(obs.data_ptr()), obs.shape[1], C.c_void_p(h.data_ptr()), k0, obs.shape[0],
                                                  C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream))
            _lib.check(rc, "hw_cast_pad_bf16")
        elif obs.shape[1] == k0:
            h = obs.to(torch.bfloat16)
        else:
            h = obs.new_zeros((obs.shape[0], k0), dtype=torch.bfloat16)
            h[:, :obs.shape[1]] = obs
        for w, b in body:
            h = torch._addmm_activation(b, h, w, use_gelu=False)
        return torch.addmm(hb, h, hw)

    @torch.no_grad()
    def step(self, obs, generator=None):
        """actions, values, neglogpacs — the triple `model.step` returns in ppo2 (runner.py:33)"""
        out, v = self.forward(obs)
        if self.continuous:
            std = self.logstd.exp()
            a = out + std * torch.randn(out.shape, device=out.device, generator=generator)
            neglogp = (0.5 * (((a - out) / std) ** 2).sum(-1) + 0.5 * 1.8378770664093453 * a.shape[-1] + self.logstd.sum())
            return a, v, neglogp  # the raw sample: what the learner stores and what neglogp belongs to (clamp only the env's copy)
        logp = torch.log_softmax(out, dim=-1)
        # categorical sample by inverse CDF on one uniform per env (cheaper than torch.multinomial at 10^5..10^6 rows)
        # (running sums over the handful of actions are spelled out: torch.cumsum over a last dimension of 4 costs 0.7 ms at 131k rows)
        u = torch.rand(out.shape[0], device=out.device, generator=generator)
        p = logp.exp()
        c = p[:, 0]
        a = (c < u).to(torch.int64)
        for k in range(1, out.shape[-1] - 1):
            c = c + p[:, k]
            a += c < u
        return a.to(torch.int32), v, -logp.gather(-1, a.unsqueeze(-1)).squeeze(-1)

    @torch.no_grad()
    def value(self, obs):
        return self.forward(obs)[1]


def gae(rewards, values, dones, last_values, last_dones, gamma, lam):
    """Generalised advantage estimation exactly as ppo2's runner computes it (runner.py:53-65), precision by precision:
    `gamma * nextvalues` is a float32 product, `1.0 - dones` makes everything after it float64, lastgaelam is carried in
    float64, the stored advantage is its float32 rounding and returns = advantages + values is a float32 sum.
    rewards, values: [T, N]; dones[t] = the `self.dones` the runner stored at step t, i.e. whether the episode
    ended on the transition BEFORE step t; last_dones = dones after the final step.  Returns (returns, advantages)."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    lastgaelam = torch.zeros_like(last_values, dtype=torch.float64)
    gl = float(gamma) * float(lam)
    for t in reversed(range(T)):
        if t == T - 1:
            nextnonterminal = 1.0 - last_dones.to(torch.float64)
            nextvalues = last_values
        else:
            nextnonterminal = 1.0 - dones[t + 1].to(torch.float64)
            nextvalues = values[t + 1]
        gv = nextvalues * float(gamma)  # a float32 product: the scalar adopts the tensor's type
        delta = rewards[t].to(torch.float64) + gv.to(torch.float64) * nextnonterminal - values[t].to(torch.float64)
        lastgaelam = delta + gl * nextnonterminal * lastgaelam
        adv[t] = lastgaelam.to(torch.float32)
    return adv + values, adv


def discount_with_dones(rewards, dones, gamma, bootstrap=None):
    """n-step discounted returns of a2c's runner (a2c/utils.py:147-153 discount_with_dones, a2c/runner.py:55-66): the
    recursion runs over Python floats there, i.e. in float64 on the float32 rewards / bootstrap values, and the result is
    stored back as float32.  rewards, dones: [T, N] with dones[t] = episode ended ON step t; bootstrap: value of the state
    after the last step (used where the last step did not end the episode)."""
    T = rewards.shape[0]
    out = torch.zeros_like(rewards)
    nd = 1.0 - dones.to(torch.float64)
    r = torch.zeros_like(rewards[0], dtype=torch.float64) if bootstrap is None else bootstrap.to(torch.float64) * nd[-1]
    for t in reversed(range(T)):
        r = rewards[t].to(torch.float64) + (gamma * r) * nd[t]
        out[t] = r.to(torch.float32)
    return out


def sf01(x):
    """swap and then flatten axes 0 and 1 (ppo2/runner.py:69-74): [T, N, ...] -> [N*T, ...], environment major"""
    s = x.shape
    return x.transpose(0, 1).reshape(s[0] * s[1], *s[2:])


class Runner(object):
    """ppo2's Runner over a GPU VecEnv: `run()` returns (obs, returns, masks, actions, values, neglogpacs, states, epinfos)"""

    def __init__(self, env, model, nsteps, gamma=0.99, lam=0.95, generator=None):
        self.env, self.model, self.nsteps, self.gamma, self.lam, self.generator = env, model, nsteps, gamma, lam, generator
        self.obs = env.reset().clone()
        self.dones = torch.zeros(env.num_envs, dtype=torch.float32, device=self.obs.device)
        n, dev = env.num_envs, self.obs.device
        cont = getattr(env, "continuous", False)
        self.mb_obs = torch.empty((nsteps, n) + tuple(self.obs.shape[1:]), dtype=torch.float32, device=dev)
        self.mb_rewards = torch.empty(nsteps, n, dtype=torch.float32, device=dev)
        self.mb_actions = torch.empty((nsteps, n, 2) if cont else (nsteps, n), dtype=torch.float32 if cont else torch.int32, device=dev)
        self.mb_values = torch.empty(nsteps, n, dtype=torch.float32, device=dev)
        self.mb_neglogpacs = torch.empty(nsteps, n, dtype=torch.float32, device=dev)
        self.mb_dones = torch.empty(nsteps, n, dtype=torch.float32, device=dev)

    @torch.no_grad()
    def run(self, want_epinfos=True):
        epinfos = []
        if getattr(self.model, "bf16_inference", False):
            self.model.refresh_bf16_()  # in place, inside a captured graph too: replays follow optimizer.step()
        cont = getattr(self.env, "continuous", False)
        for t in range(self.nsteps):
            actions, values, neglogpacs = self.model.step(self.obs, generator=self.generator)
            self.mb_obs[t].copy_(self.obs)
            self.mb_actions[t].copy_(actions)   # the raw sample (what neglogpacs belongs to) ...
            self.mb_values[t].copy_(values)
            self.mb_neglogpacs[t].copy_(neglogpacs)
            self.mb_dones[t].copy_(self.dones)
            obs, rewards, dones, infos = self.env.step(actions.clamp(-1, 1) if cont else actions)  # ... the env acts on it clamped to its Box
            self.obs = obs
            self.dones = dones.to(torch.float32)
            self.mb_rewards[t].copy_(rewards)
            if want_epinfos:
                epinfos.extend(infos.episodes())
        last_values = self.model.value(self.obs)
        returns, _ = gae(self.mb_rewards, self.mb_values, self.mb_dones, last_values, self.dones, self.gamma, self.lam)
        return (sf01(self.mb_obs), sf01(returns), sf01(self.mb_dones), sf01(self.mb_actions), sf01(self.mb_values),
                sf01(self.mb_neglogpacs), None, epinfos)

    @torch.no_grad()
    def run_graphed(self):
        """`run(want_epinfos=False)` replayed from ONE CUDA graph holding the whole rollout (nsteps x (policy forward,
        sampling, buffer writes, env step kernel) + GAE + sf01).  Eagerly a step is ~45 small launches and the host
        cannot issue them as fast as the GPU retires them at 10^5 envs; the graph removes that bound.  The first call
        captures (after one eager rollout as warm-up); the returned tensors live in the graph's memory pool and are
        overwritten by the next call.  Episode statistics come from `env.episode_statistics()` (device counters);
        per-episode infos are not available on this path.  nsteps must be even (the env alternates two output
        buffers, and a replay has to start on the one the capture started on)."""
        if self.nsteps % 2:
            raise ValueError("run_graphed needs an even nsteps")
        if getattr(self, "_graph", None) is None:
            if self.generator is not None:
                raise ValueError("run_graphed samples from the default CUDA generator (it is the one graph capture tracks); construct the Runner with generator=None")
            self.run(want_epinfos=False)  # warm-up: cuBLAS workspaces, allocator pools
            # the env hands out its own output buffer: keep private copies so that replays start from well-defined tensors
            self._g_obs = self.obs.clone()
            self._g_dones = self.dones.clone()
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self.obs, self.dones = self._g_obs, self._g_dones
                out = self.run(want_epinfos=False)
                self._g_obs.copy_(self.obs)
                self._g_dones.copy_(self.dones)
                self._g_out = out[:7]
            self.obs, self.dones = self._g_obs, self._g_dones
        self._graph.replay()
        return self._g_out + ([],)


class A2CRunner(object):
    """a2c's Runner over a GPU VecEnv (baselines/a2c/runner.py:21-72): `run()` returns
    (obs, states, rewards, masks, actions, values) flattened environment-major, where `rewards` are the n-step discounted
    returns bootstrapped off the value function where the last step did not end the episode, and `masks[e, t]` says whether
    the episode of env e ended on the transition BEFORE step t."""

    def __init__(self, env, model, nsteps=5, gamma=0.99, generator=None):
        self.env, self.model, self.nsteps, self.gamma, self.generator = env, model, nsteps, gamma, generator
        self.obs = env.reset().clone()
        n, dev = env.num_envs, self.obs.device
        cont = getattr(env, "continuous", False)
        self.dones = torch.zeros(n, dtype=torch.float32, device=dev)
        self.mb_obs = torch.empty((nsteps, n) + tuple(self.obs.shape[1:]), dtype=torch.float32, device=dev)
        self.mb_rewards = torch.empty(nsteps, n, dtype=torch.float32, device=dev)
        self.mb_actions = torch.empty((nsteps, n, 2) if cont else (nsteps, n), dtype=torch.float32 if cont else torch.int32, device=dev)
        self.mb_values = torch.empty(nsteps, n, dtype=torch.float32, device=dev)
        self.mb_dones = torch.empty(nsteps + 1, n, dtype=torch.float32, device=dev)

    @torch.no_grad()
    def run(self):
        if getattr(self.model, "bf16_inference", False):
            self.model.refresh_bf16_()
        cont = getattr(self.env, "continuous", False)
        for t in range(self.nsteps):
            actions, values, _ = self.model.step(self.obs, generator=self.generator)
            self.mb_obs[t].copy_(self.obs)
            self.mb_actions[t].copy_(actions)
            self.mb_values[t].copy_(values)
            self.mb_dones[t].copy_(self.dones)
            obs, rewards, dones, _ = self.env.step(actions.clamp(-1, 1) if cont else actions)
            self.obs = obs
            self.dones = dones.to(torch.float32)
            self.mb_rewards[t].copy_(rewards)
        self.mb_dones[self.nsteps].copy_(self.dones)
        masks, dones = self.mb_dones[:-1], self.mb_dones[1:]
        rewards = self.mb_rewards
        if self.gamma > 0.0:
            rewards = discount_with_dones(self.mb_rewards, dones, self.gamma, bootstrap=self.model.value(self.obs))
        return sf01(self.mb_obs), None, sf01(rewards), sf01(masks), sf01(self.mb_actions), sf01(self.mb_values)


class _DeviceCollector(object):
    """what the on-device runners share: [T, rows] rollout buffers the kernels write in place, the fused sampling launch, and
    CUDA-graph capture of a whole rollout.  `rows` = num_envs for the single-agent env and num_envs * A for the multi-agent
    env (every car is one row of the shared policy's batch: observations [N, A, 4A+14] are [N * A, 4A+14] rows in memory,
    rewards / dones [N, A] are [N * A])."""

    def __init__(self, env, model, nsteps, seed):
        self.continuous = bool(getattr(env, "continuous", False))
        if self.continuous != bool(model.continuous):
            raise ValueError("policy and env disagree on the action space")
        from . import _lib
        self._libmod, self._lib = _lib, _lib.load()
        self.env, self.model, self.nsteps = env, model, nsteps
        self.seed = int(seed) & (2 ** 64 - 1)
        self.agents = int(getattr(env, "num_agents", 0))          # 0: single-agent env
        A = max(self.agents, 1)
        self.rows = env.num_envs * A
        self.row_id0 = env._env_id0 * A                           # global row id of this shard's first row (Philox key)
        self.obs_dim = int(env.obs_dim) if self.agents else 18
        if model.body[0].in_features != self.obs_dim:
            raise ValueError("the policy reads %d inputs, the env observes %d" % (model.body[0].in_features, self.obs_dim))
        n, dev = self.rows, env.device
        self.n_actions = model.pi.out_features
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=dev)   # steps sampled so far (device side: survives graph replay)
        self.mb_obs = torch.empty(nsteps + 1, n, self.obs_dim, dtype=torch.float32, device=dev)   # row t: what the policy saw at step t
        self.mb_done_u8 = torch.zeros(nsteps + 1, n, dtype=torch.uint8, device=dev)      # row t: episode ended BEFORE step t
        self.mb_rewards = torch.empty(nsteps, n, dtype=torch.float32, device=dev)
        self.mb_actions = (torch.empty(nsteps, n, 2, dtype=torch.float32, device=dev) if self.continuous
                           else torch.empty(nsteps, n, dtype=torch.int32, device=dev))
        self.env_actions = torch.empty(n, 2, dtype=torch.float32, device=dev) if self.continuous else None  # clamped to [-1, 1]
        self.mb_values = torch.empty(nsteps, n, dtype=torch.float32, device=dev)
        self.mb_neglogpacs = torch.empty(nsteps, n, dtype=torch.float32, device=dev)
        self.mb_obs[nsteps].copy_(env.reset().reshape(n, self.obs_dim))
        self._graph = None

    def _stream(self):
        import ctypes as C
        return C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)

    def _sample(self, head, t):
        import ctypes as C
        p = lambda x: C.c_void_p(x.data_ptr())
        bf16 = 1 if head.dtype == torch.bfloat16 else 0
        with torch.cuda.device(self.env.device):
            if self.continuous:
                rc = self._lib.hw_policy_sample_gaussian2(p(head), bf16, head.shape[1], p(self.model.logstd.detach()), self.seed,
                                                          self.row_id0, t, p(self.step_counter), p(self.mb_actions[t]),
                                                          p(self.env_actions), p(self.mb_neglogpacs[t]), p(self.mb_values[t]),
                                                          head.shape[0], self._stream())
            else:
                rc = self._lib.hw_policy_sample_discrete(p(head), bf16, head.shape[1], self.n_actions, self.seed, self.row_id0, t,
                                                         p(self.step_counter), p(self.mb_actions[t]), p(self.mb_neglogpacs[t]),
                                                         p(self.mb_values[t]), head.shape[0], self._stream())
        self._libmod.check(rc, "hw_policy_sample")
        return self.env_actions if self.continuous else self.mb_actions[t]

    def _collect(self):
        """T x (policy forward, one sampling launch, one env step writing obs / reward / done into row t + 1 / t / t + 1)"""
        T = self.nsteps
        self.model.refresh_bf16_()  # bf16 weight copies re-cast in place: part of the captured graph, so replays see optimizer steps
        # the last observation / done row of the previous rollout is the first of this one
        self.mb_obs[0].copy_(self.mb_obs[T])
        self.mb_done_u8[0].copy_(self.mb_done_u8[T])
        for t in range(T):
            act = self._sample(self.model.head_bf16(self.mb_obs[t]), t)
            self.env.step_into(act, self.mb_obs[t + 1], self.mb_rewards[t], self.mb_done_u8[t + 1])
        self.step_counter += T

    @torch.no_grad()
    def run_graphed(self):
        """`run()` replayed from one CUDA graph (captured on the first call, after an eager warm-up rollout); the returned
        tensors live in the graph's memory pool and are overwritten by the next call"""
        if self._graph is None:
            self.run()
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._g_out = self.run()
        self._graph.replay()
        return self._g_out


class DeviceRunner(_DeviceCollector):
    """ppo2's Runner (baselines/ppo2/runner.py:20-66) for a `HighwayVecEnv` or a `HighwayMAVecEnv` (discrete or continuous), with
    nothing between the policy's GEMMs and the env kernel but ONE launch: `hw_policy_sample_discrete` /
    `hw_policy_sample_gaussian2` turns the merged head's output into actions, values and neglogpacs written in place into the
    rollout's [T, rows] buffers, and the env step writes the next observation, the reward and the done flag straight into
    theirs (`step_into`) - no per-step copies, casts or elementwise torch kernels.  `run()` returns what
    `Runner.run(want_epinfos=False)` returns; `run_graphed()` replays the rollout from one CUDA graph.
    Actions are sampled from a Philox stream keyed by (seed, global row id, step counter), so a rollout is reproducible
    and independent of how the envs are sharded over GPUs.  For the multi-agent env every car is a row of the batch (the
    cars share the policy; a row's done flag is its own car's, and all cars of an env restart together when any of them
    finishes, dummy_vec_ma_env.py:74-77)."""

    def __init__(self, env, model, nsteps, gamma=0.99, lam=0.95, seed=0):
        _DeviceCollector.__init__(self, env, model, nsteps, seed)
        self.gamma, self.lam = gamma, lam

    @torch.no_grad()
    def run(self):
        T = self.nsteps
        self._collect()
        dones = self.mb_done_u8.to(torch.float32)
        last_values = self.model.value(self.mb_obs[T]).contiguous()
        # GAE in one launch (the torch loop is 8 small kernels per step: a fifth of the rollout at T = 240)
        import ctypes as C
        returns = torch.empty_like(self.mb_rewards)
        with torch.cuda.device(self.env.device):
            rc = self._lib.hw_gae(C.c_void_p(self.mb_rewards.data_ptr()), C.c_void_p(self.mb_values.data_ptr()),
                                  C.c_void_p(self.mb_done_u8.data_ptr()), C.c_void_p(last_values.data_ptr()), self.gamma, self.lam,
                                  C.c_void_p(returns.data_ptr()), None, T, self.rows, self._stream())
        self._libmod.check(rc, "hw_gae")
        return (sf01(self.mb_obs[:T]), sf01(returns), sf01(dones[:T]), sf01(self.mb_actions), sf01(self.mb_values),
                sf01(self.mb_neglogpacs), None, [])


class DeviceA2CRunner(_DeviceCollector):
    """a2c's Runner (baselines/a2c/runner.py:21-72) on the device path of `DeviceRunner`: fused sampling, `step_into`, and the
    n-step discounted returns with value bootstrap in one launch (`hw_nstep_returns`: discount_with_dones, a2c/utils.py:147-153,
    in the float64 of the reference's Python floats); `run_graphed()` replays the nsteps (5 by default) from one CUDA graph.
    `run()` returns (obs, states, rewards, masks, actions, values) flattened environment-major like the reference."""

    def __init__(self, env, model, nsteps=5, gamma=0.99, seed=0):
        _DeviceCollector.__init__(self, env, model, nsteps, seed)
        self.gamma = gamma

    @torch.no_grad()
    def run(self):
        T = self.nsteps
        self._collect()
        rewards = self.mb_rewards
        if self.gamma > 0.0:
            import ctypes as C
            last_values = self.model.value(self.mb_obs[T]).contiguous()
            rewards = torch.empty_like(self.mb_rewards)
            with torch.cuda.device(self.env.device):
                rc = self._lib.hw_nstep_returns(C.c_void_p(self.mb_rewards.data_ptr()), C.c_void_p(self.mb_done_u8.data_ptr()),
                                                C.c_void_p(last_values.data_ptr()), self.gamma, C.c_void_p(rewards.data_ptr()), T,
                                                self.rows, self._stream())
            self._libmod.check(rc, "hw_nstep_returns")
        masks = self.mb_done_u8[:T].to(torch.float32)
        return sf01(self.mb_obs[:T]), None, sf01(rewards), sf01(masks), sf01(self.mb_actions), sf01(self.mb_values)
