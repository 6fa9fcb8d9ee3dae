import importlib

registry = {}


def register(id, entry_point=None, kwargs=None, **_ignored):
    registry[id] = (entry_point, dict(kwargs or {}))


def make(id, **kwargs):
    entry_point, base_kwargs = registry[id]
    merged = dict(base_kwargs)
    merged.update(kwargs)
    mod_name, attr = entry_point.split(':')
    cls = getattr(importlib.import_module(mod_name), attr)
    return cls(**merged)
