registry = {}


class EnvSpec(object):
    def __init__(self, id, entry_point=None, kwargs=None, **_):
        self.id = id
        self.entry_point = entry_point
        self.kwargs = kwargs or {}


def register(id, entry_point=None, kwargs=None, **kw):
    registry[id] = EnvSpec(id, entry_point, kwargs)
