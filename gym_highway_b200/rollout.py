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
            with torch.autocast("cuda", dtype=torch.bfloat16):
                h = self.body(obs)
                return self.pi(h).float(), self.vf(h).float().squeeze(-1)
        h = self.body(obs)
        return self.pi(h), self.vf(h).squeeze(-1)

    @torch.no_grad()
    def step(self, obs, generator=None):
        """actions, values, neglogpacs — the triple `model.step` returns in ppo2 (runner.py:33)"""
        out, v = self.forward(obs)
        if self.continuous:
            std = self.logstd.exp()
            a = out + std * torch.randn(out.shape, device=out.device, generator=generator)
            neglogp = (0.5 * (((a - out) / std) ** 2).sum(-1) + 0.5 * 1.8378770664093453 * a.shape[-1] + self.logstd.sum())
            return a.clamp(-1, 1), v, neglogp
        logp = torch.log_softmax(out, dim=-1)
        # categorical sample by inverse CDF on one uniform per env (cheaper than torch.multinomial at 10^5..10^6 rows)
        u = torch.rand(out.shape[0], 1, device=out.device, generator=generator)
        a = (logp.exp().cumsum(-1) < u).sum(-1).clamp_(max=out.shape[-1] - 1)
        return a.to(torch.int32), v, -logp.gather(-1, a.unsqueeze(-1)).squeeze(-1)

    @torch.no_grad()
    def value(self, obs):
        return self.forward(obs)[1]


def gae(rewards, values, dones, last_values, last_dones, gamma, lam):
    """Generalised advantage estimation exactly as ppo2's runner computes it (runner.py:53-65).
    rewards, values: [T, N]; dones[t] = the `self.dones` the runner stored at step t, i.e. whether the episode
    ended on the transition BEFORE step t; last_dones = dones after the final step.  Returns (returns, advantages)."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    lastgaelam = torch.zeros_like(last_values)
    for t in reversed(range(T)):
        if t == T - 1:
            nextnonterminal = 1.0 - last_dones
            nextvalues = last_values
        else:
            nextnonterminal = 1.0 - dones[t + 1]
            nextvalues = values[t + 1]
        delta = rewards[t] + gamma * nextvalues * nextnonterminal - values[t]
        lastgaelam = delta + gamma * lam * nextnonterminal * lastgaelam
        adv[t] = lastgaelam
    return adv + values, adv


def discount_with_dones(rewards, dones, gamma, bootstrap=None):
    """n-step discounted returns of a2c's runner (a2c/utils.py discount_with_dones, a2c/runner.py:55-66).
    rewards, dones: [T, N] with dones[t] = episode ended ON step t; bootstrap: value of the state after the last step
    (used where the last step did not end the episode)."""
    T = rewards.shape[0]
    out = torch.zeros_like(rewards)
    r = torch.zeros_like(rewards[0]) if bootstrap is None else bootstrap * (1.0 - dones[-1])
    for t in reversed(range(T)):
        r = rewards[t] + gamma * r * (1.0 - dones[t])
        out[t] = r
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
        for t in range(self.nsteps):
            actions, values, neglogpacs = self.model.step(self.obs, generator=self.generator)
            self.mb_obs[t].copy_(self.obs)
            self.mb_actions[t].copy_(actions)
            self.mb_values[t].copy_(values)
            self.mb_neglogpacs[t].copy_(neglogpacs)
            self.mb_dones[t].copy_(self.dones)
            obs, rewards, dones, infos = self.env.step(actions)
            self.obs = obs
            self.dones = dones.to(torch.float32)
            self.mb_rewards[t].copy_(rewards)
            if want_epinfos:
                epinfos.extend(infos.episodes())
        last_values = self.model.value(self.obs)
        returns, _ = gae(self.mb_rewards, self.mb_values, self.mb_dones, last_values, self.dones, self.gamma, self.lam)
        return (sf01(self.mb_obs), sf01(returns), sf01(self.mb_dones), sf01(self.mb_actions), sf01(self.mb_values),
                sf01(self.mb_neglogpacs), None, epinfos)
