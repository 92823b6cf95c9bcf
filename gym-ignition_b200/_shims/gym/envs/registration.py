import importlib


class EnvSpec:
    def __init__(self, id, entry_point=None, max_episode_steps=None, kwargs=None, reward_threshold=None,
                 nondeterministic=False):
        self.id = id
        self.entry_point = entry_point
        self.max_episode_steps = max_episode_steps
        self.reward_threshold = reward_threshold
        self.nondeterministic = nondeterministic
        self._kwargs = {} if kwargs is None else kwargs

    def make(self, **kwargs):
        if self.entry_point is None:
            raise RuntimeError(f"Attempting to make deprecated env {self.id}")
        merged = dict(self._kwargs)
        merged.update(kwargs)
        if callable(self.entry_point):
            env = self.entry_point(**merged)
        else:
            mod_name, attr = self.entry_point.split(":")
            env = getattr(importlib.import_module(mod_name), attr)(**merged)
        env.unwrapped.spec = self
        return env


class EnvRegistry:
    def __init__(self):
        self.env_specs = {}

    def register(self, id, **kwargs):
        if id in self.env_specs:
            raise RuntimeError(f"Cannot re-register id: {id}")
        self.env_specs[id] = EnvSpec(id, **kwargs)

    def spec(self, id):
        if id not in self.env_specs:
            raise RuntimeError(f"No registered env with id: {id}")
        return self.env_specs[id]

    def make(self, id, **kwargs):
        spec = self.spec(id)
        env = spec.make(**kwargs)
        if env.spec.max_episode_steps is not None:
            from ..wrappers.time_limit import TimeLimit
            env = TimeLimit(env, max_episode_steps=env.spec.max_episode_steps)
        return env

    def all(self):
        return self.env_specs.values()


registry = EnvRegistry()


def register(id, **kwargs):
    return registry.register(id, **kwargs)


def make(id, **kwargs):
    return registry.make(id, **kwargs)


def spec(id):
    return registry.spec(id)
