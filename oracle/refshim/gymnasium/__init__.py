"""Stand-in for `gymnasium` (reference BaseAviary.py:14,18). Test infrastructure only."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def reset(self, seed=None, options=None):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError
