"""Episode statistics in the reference's Monitor format, fed from the kernels' terminal-info tensors.

Mirrors baselines/bench/monitor.py:58-124: `*.monitor.csv` = one `# {json header}` line, a csv header
`r,l,t`, then one row per finished episode with r = round(return, 6), l = length, t = seconds since start;
plus the deque(maxlen=100) means the trainers print (`eprewmean` / `eplenmean`, ppo2/ppo2.py:119,194-195).
The per-env accounting itself (sum of rewards, step count, reset on done) is fused into the step kernels;
this class only drains the few finished episodes of each step to the host.
"""
import csv
import json
import os
import time
from collections import deque

EXT = "monitor.csv"


class ResultsWriter(object):
    def __init__(self, filename=None, header='', extra_keys=()):
        self.extra_keys = tuple(extra_keys)
        self.f = None
        self.logger = None
        if filename is None:
            return
        if not filename.endswith(EXT):
            filename = os.path.join(filename, EXT) if os.path.isdir(filename) else filename + "." + EXT
        self.f = open(filename, "wt")
        if isinstance(header, dict):
            header = '# {} \n'.format(json.dumps(header))
        self.f.write(header)
        self.logger = csv.DictWriter(self.f, fieldnames=('r', 'l', 't') + self.extra_keys)
        self.logger.writeheader()
        self.f.flush()

    def write_row(self, epinfo):
        if self.logger:
            self.logger.writerow({k: epinfo[k] for k in ('r', 'l', 't') + self.extra_keys})
            self.f.flush()

    def close(self):
        if self.f is not None:
            self.f.close()
            self.f = None


class VecMonitor(object):
    """wraps a GPU VecEnv: same step()/reset() results, episodes logged like `Monitor` would"""

    def __init__(self, venv, filename=None, env_id=None, keep=100):
        self.venv = venv
        self.tstart = time.time()
        self.results_writer = ResultsWriter(filename, header={"t_start": self.tstart, "env_id": env_id})
        self.epinfobuf = deque(maxlen=keep)
        self.episode_rewards, self.episode_lengths, self.episode_times = [], [], []
        self.total_steps = 0

    def __getattr__(self, name):
        return getattr(self.venv, name)

    def reset(self):
        return self.venv.reset()

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        obs, rews, dones, infos = self.venv.step_wait()
        self.total_steps += self.venv.num_envs
        for ep in infos.episodes():
            ep = dict(ep, t=round(time.time() - self.tstart, 6))
            self.episode_rewards.append(ep['r']); self.episode_lengths.append(ep['l']); self.episode_times.append(ep['t'])
            self.epinfobuf.append(ep)
            self.results_writer.write_row(ep)
        return obs, rews, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def means(self):
        """(eprewmean, eplenmean) over the last `keep` episodes — safemean of ppo2.py:212-213"""
        if not self.epinfobuf:
            return float('nan'), float('nan')
        n = len(self.epinfobuf)
        return sum(e['r'] for e in self.epinfobuf) / n, sum(e['l'] for e in self.epinfobuf) / n

    def get_total_steps(self):
        return self.total_steps

    def get_episode_rewards(self):
        return self.episode_rewards

    def get_episode_lengths(self):
        return self.episode_lengths

    def get_episode_times(self):
        return self.episode_times

    def close(self):
        self.results_writer.close()
        self.venv.close()


def load_results(path):
    """rows of a monitor file as dicts (subset of baselines.bench.monitor.load_results, without pandas)"""
    with open(path, "rt") as fh:
        first = fh.readline()
        assert first[0] == '#'
        header = json.loads(first[1:])
        rows = [dict(r=float(r['r']), l=int(r['l']), t=float(r['t'])) for r in csv.DictReader(fh)]
    return header, rows
