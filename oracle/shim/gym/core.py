"""gym.core stand-in: the reference's bench/monitor.py imports `Wrapper` from here.  TEST INFRASTRUCTURE ONLY."""


class Env(object):
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    @property
    def unwrapped(self):
        return self

    def close(self):
        return None

    def seed(self, seed=None):
        return []


class Wrapper(Env):
    """forwards step / reset / close / seed and unknown attributes to the wrapped env (gym 0.10 semantics)"""

    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.reward_range = getattr(env, "reward_range", Env.reward_range)
        self.metadata = getattr(env, "metadata", {})

    @property
    def spec(self):
        return getattr(self.env, "spec", None)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)
