"""Stand-in for `gymnasium.envs.registration` (reference __init__.py:1). Test infrastructure only."""
registry = {}


def register(id, entry_point=None, **kw):
    registry[id] = entry_point
