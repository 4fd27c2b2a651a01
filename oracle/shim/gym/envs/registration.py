import importlib
from .. import error


class EnvSpec(object):
    def __init__(self, id, entry_point=None, kwargs=None, **_):
        self.id = id
        self.entry_point = entry_point
        self._kwargs = {} if kwargs is None else kwargs

    def make(self, **kwargs):
        kw = dict(self._kwargs)
        kw.update(kwargs)
        if callable(self.entry_point):
            cls = self.entry_point
        else:
            mod_name, attr = self.entry_point.split(':')
            cls = getattr(importlib.import_module(mod_name), attr)
        env = cls(**kw)
        env.unwrapped.spec = self
        return env


class EnvRegistry(object):
    def __init__(self):
        self.env_specs = {}

    def register(self, id, **kwargs):
        # classic gym raised on re-registration; the reference never re-registers
        self.env_specs[id] = EnvSpec(id, **kwargs)

    def spec(self, id):
        try:
            return self.env_specs[id]
        except KeyError:
            raise error.UnregisteredEnv('No registered env with id: {}'.format(id))

    def make(self, id, **kwargs):
        return self.spec(id).make(**kwargs)

    def all(self):
        return self.env_specs.values()


registry = EnvRegistry()


def register(id, **kwargs):
    return registry.register(id, **kwargs)


def make(id, **kwargs):
    return registry.make(id, **kwargs)


def spec(id):
    return registry.spec(id)
